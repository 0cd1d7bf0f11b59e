"""Tuning aid: one run of the giant component (cfg5) for ncu: 64 chains x 100k iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg5")
dp = m.DeviceProblem(p)
X = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
o = dp.run(m.MODE_FAST, T, n_chains=X, want_state=False, per_chain=False)
print('X', X, 'T', T, 'kernel %.1f ms' % o.kernel_ms, 'reassign', int(o.totals[4]), flush=True)
