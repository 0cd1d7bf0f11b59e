"""Tuning aid: device time of every launch (size class) of one sampler step on the 5M-fragment mixed config.\nMBSV_SERIAL_LAUNCHES=1 runs the classes back to back so that each time is the standalone cost."""
import sys, time, os; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
import multibreak_sv_b200 as m
from multibreak_sv_b200 import synth
p = synth.generate("cfg4", skeleton=synth.draw_skeleton("cfg4"))
dp = m.DeviceProblem(p)
for it in range(2):
    o = dp.run(m.MODE_FAST, int(os.environ.get('ITERS', 2000)), n_chains=64, want_state=False, per_chain=False, memo_states=int(os.environ.get('MEMO', 65536)))
print('kernel %.1f ms'%o.kernel_ms, 'reassign/s %.3e'%(o.totals[4]/(o.kernel_ms/1e3)))
for li in dp.launch_info(): print('  rows %4d warps %8d ms %7.1f'%(li['rows'], li['warps'], li['ms']))
