import sys, time; sys.path.insert(0,'.')
import numpy as np
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg4", n_frag=1000000)
dp = m.DeviceProblem(p)
for memo in (256, 512, 1024, 2048):
    for it in range(2):
        o = dp.run(m.MODE_FAST, 2000, n_chains=64, want_state=False, per_chain=False, memo_states=memo)
    print(memo, 'kernel %.1f ms'%o.kernel_ms, 'reassign/s %.3e'%(o.totals[4]/(o.kernel_ms/1e3)), ['%d:%.0f'%(li['rows'],li['ms']) for li in dp.launch_info()])
