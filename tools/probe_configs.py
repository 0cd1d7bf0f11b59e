"""Tuning aid: throughput of BASELINE configs[1] (cfg2: paired-end, 1 chain) and configs[2] (cfg3: long reads, 16 chains)
at their quoted iteration counts (GPU only; the CPU baseline lives in bench.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth

for cfg, T in (("cfg2", 100000), ("cfg3", 20000)):
    c = synth.CONFIGS[cfg]
    p = synth.generate(cfg)
    dp = m.DeviceProblem(p)
    for memo in (65536, 0):
        for it in range(2):
            g = dp.run(m.MODE_FAST, T, n_chains=c.n_chains, want_state=False, per_chain=False, memo_states=memo)
        print(cfg, "memo_states", memo, p.scalars["n_sub"], "subs x", c.n_chains, "chains x", T, "iters: kernel %.1f ms" % g.kernel_ms,
              "reassign/s %.3e" % (g.totals[4] / (g.kernel_ms / 1e3)), [(li["rows"], li["warps"], round(li["ms"], 1)) for li in dp.launch_info()])
