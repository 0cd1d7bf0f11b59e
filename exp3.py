import sys, time; sys.path.insert(0,'.')
import numpy as np
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
sk = synth.draw_skeleton("cfg4")
p = synth.generate("cfg4", skeleton=sk)
for it in range(3):
    t0=time.perf_counter(); d = p.desc(); t1=time.perf_counter()
    dp = m.DeviceProblem(p); t2=time.perf_counter()
    o = dp.run(m.MODE_FAST, 2000, n_chains=64, want_state=False, per_chain=False); t3=time.perf_counter()
    dp.close(); t4=time.perf_counter()
    print('desc %.3f create %.3f run %.3f (kernel %.3f) close %.3f'%(t1-t0, t2-t1, t3-t2, o.kernel_ms/1e3, t4-t3))
