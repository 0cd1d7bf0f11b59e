"""Tuning aid: giant single component (BASELINE configs[4]): setup vs per-iteration cost for 64 and 1024 chains."""
import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg5")
print(p.scalars)
dp = m.DeviceProblem(p)
for X in (8, 64, 1024):
    for T in (0, 20000, 100000):
        o = dp.run(m.MODE_FAST, T, n_chains=X, want_state=False, per_chain=False)
        print('X', X, 'T', T, 'kernel %.1f ms'%o.kernel_ms, 'reassign', int(o.totals[4]))
