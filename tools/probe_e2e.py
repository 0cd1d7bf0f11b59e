"""Tuning aid: where the end-to-end step (mbsv_problem_create + mbsv_run with host buffers + destroy) spends its time on
the 5M-fragment config.  MBSV_DEBUG_TIMING=1 adds the library's own laps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg4", skeleton=synth.draw_skeleton("cfg4"))
for it in range(3):
    t0 = time.perf_counter()
    dp = m.DeviceProblem(p)
    t1 = time.perf_counter()
    o = dp.run(m.MODE_FAST, 5000, n_chains=64, want_state=False, per_chain=False)
    t2 = time.perf_counter()
    dp.close()
    t3 = time.perf_counter()
    print("create %.1f ms  run %.1f ms (kernel %.1f)  destroy %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), o.kernel_ms, 1e3 * (t3 - t2)), flush=True)
