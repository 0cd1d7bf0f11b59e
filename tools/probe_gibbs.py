"""Tuning aid: throughput of MBSV_MODE_GIBBS next to MBSV_MODE_FAST on a reduced mixed config (no connected components)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
import dataclasses
# (Gibbs mode keeps a chain's state in shared memory: subproblems are capped at 2000 fragments here)
p = synth.generate(dataclasses.replace(synth.CONFIGS["cfg4"], sub_max=2000), n_frag=1_000_000, seed=2)
dp = m.DeviceProblem(p)
for mode, name in ((m.MODE_GIBBS, "gibbs"), (m.MODE_FAST, "fast ")):
    for it in range(2):
        g = dp.run(mode, 2000, n_chains=64, want_state=False, per_chain=False)
    print(name, "kernel %.1f ms" % g.kernel_ms, "updates/proposals per s %.3e" % (float(g.totals[0] + g.totals[2]) / (g.kernel_ms / 1e3)),
          "state changes per s %.3e" % (float(g.totals[1] + g.totals[3]) / (g.kernel_ms / 1e3)))
