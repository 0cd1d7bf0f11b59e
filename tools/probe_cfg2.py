import os, sys
sys.path.insert(0, "/root/repo")
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg2")
dp = m.DeviceProblem(p)
for ms in (65536,):
    for it in range(2):
        g = dp.run(m.MODE_FAST, 100000, n_chains=1, want_state=False, per_chain=False, memo_states=ms)
    print("memo_states", ms, "kernel %.1f ms" % g.kernel_ms, "reassign/s %.3e" % (g.totals[4] / (g.kernel_ms / 1e3)), [(li["rows"], li["warps"], round(li["ms"], 1)) for li in dp.launch_info()])
