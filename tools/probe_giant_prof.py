"""Tuning aid: phase cycle counts of the speculative-window kernel (build with EXTRA=-DMBSV_GIANT_PROF)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg5")
dp = m.DeviceProblem(p)
for X in (8, 64):
    o = dp.run(m.MODE_FAST, 100000, n_chains=X, want_state=False, per_chain=False)
    print('X', X, 'kernel %.1f ms' % o.kernel_ms, 'reassign', int(o.totals[4]), 'accepted', int(o.totals[1] + o.totals[3]), flush=True)
