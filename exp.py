"""Kernel-tuning probe: 1M-fragment cfg4, 64 chains, 500 iterations; prints device time per size class."""
import sys, time; sys.path.insert(0,'.')
import numpy as np
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg4", n_frag=1000000)
dp = m.DeviceProblem(p)
best = None
for it in range(3):
    o = dp.run(m.MODE_FAST, 2000, n_chains=64, want_state=False, per_chain=False)
    if best is None or o.kernel_ms < best[0]: best = (o.kernel_ms, dp.launch_info())
print('kernel %.1f ms'%best[0], 'reassign/s %.3e'%(o.totals[4]/(best[0]/1e3)), ['%d:%.0f'%(li['rows'],li['ms']) for li in best[1]], 'chk', int(o.aln_count.sum()%1000003), int(o.cluster_hist.sum()%1000003))
o0 = dp.run(m.MODE_FAST, 2000, n_chains=64, want_state=False, per_chain=False, memo_states=0)
print('memo off: kernel %.1f ms'%o0.kernel_ms, ['%d:%.0f'%(li['rows'],li['ms']) for li in dp.launch_info()], 'same tallies', bool((o0.aln_count==o.aln_count).all() and (o0.cluster_hist==o.cluster_hist).all()))
import ctypes as C
st = np.zeros(7, np.int64); dp.lib.mbsv_problem_stats(dp.handle, st.ctypes.data_as(C.POINTER(C.c_int64)), 7); print('stats', st.tolist())
