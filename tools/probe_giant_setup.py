"""Tuning aid: set-up phases of the speculative-window kernel (profile build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg5")
dp = m.DeviceProblem(p)
for X in (8, 64):
    for T in (0, 0, 20000):
        o = dp.run(m.MODE_FAST, T, n_chains=X, want_state=False, per_chain=False)
        print('X', X, 'T', T, 'kernel %.2f ms' % o.kernel_ms, flush=True)
