import sys, time, os; sys.path.insert(0,'.')
import numpy as np
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg4", skeleton=synth.draw_skeleton("cfg4"))
dp = m.DeviceProblem(p)
for T in (0, 2000, 4000):
    o = dp.run(m.MODE_FAST, T, n_chains=64, want_state=False, per_chain=False)
    print('T', T, 'kernel %.1f ms'%o.kernel_ms, [ '%d:%.0f'%(li['rows'],li['ms']) for li in dp.launch_info()])
