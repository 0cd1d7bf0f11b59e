import sys, time, json; sys.path.insert(0,'.')
import numpy as np
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
for nf, X, T in ((200000, 64, 500), (1000000, 64, 500)):
    p = synth.generate("cfg4", n_frag=nf)
    t=time.time(); dp = m.DeviceProblem(p); tc=time.time()-t
    for it in range(2):
        t=time.time(); o = dp.run(m.MODE_FAST, T, n_chains=X, want_state=False, per_chain=False); dt=time.time()-t
    print(nf, X, T, 'create %.2fs run %.2fs kernel %.1f ms'%(tc, dt, o.kernel_ms), 'reassign/s %.3e'%(o.totals[4]/(o.kernel_ms/1e3)), 'iters/s %.3e'%(p.scalars['n_sub']*X*T/(o.kernel_ms/1e3)), 'acc', o.totals[1]/o.totals[0])
    for li in dp.launch_info(): print('   ', li)
    dp.close()
