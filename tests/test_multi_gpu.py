"""The giant-component multi-GPU path inside the library: chains split over GPUs, tallies pooled with ONE ncclReduce
(mbsv_run_multi for a single host process, mbsv_comm_* + mbsv_reduce_tallies for one process per GPU)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_run_multi_pools_the_chains_of_all_devices(oracle, mbsv):
    from multibreak_sv_b200 import synth
    L = mbsv.load_library()
    n_dev = min(int(L.mbsv_device_count()), 2)
    p = synth.generate("cfg5", n_frag=6000, seed=41)
    op = oracle.Problem(p.scalars, p.arrays)
    probs = [mbsv.DeviceProblem(p, device=d) for d in range(n_dev)]
    X, T = 6, 4000
    g, reduce_ms = mbsv.run_multi(probs, mbsv.MODE_FAST, T, n_chains=X)
    o = oracle.run(op, "incremental", T, n_chains=X, nthreads=8, java_int_wrap=False)
    assert np.array_equal(o.aln_count, g.aln_count) and np.array_equal(o.frag_err_count, g.frag_err_count)
    assert np.array_equal(o.cluster_hist, g.cluster_hist)
    assert g.totals[:5].tolist() == o.counters.reshape(-1, 5).sum(0).tolist() and int(g.totals[5]) == 0
    assert reduce_ms >= 0.0
    # per-chain arrays are refused, not silently dropped
    r = mbsv._abi.Results()
    buf = np.zeros(X * p.scalars["n_frag"], np.int32)
    r.final_assign = buf.ctypes.data_as(mbsv._abi.I32P)
    rp = mbsv._abi.RunParams()
    rp.mode, rp.n_chains, rp.num_iters, rp.burnin, rp.perr, rp.seed = mbsv.MODE_FAST, X, 10, 10, 0.001, 123456
    handles = (C.c_void_p * n_dev)(*[q.handle for q in probs])
    assert L.mbsv_run_multi(handles, n_dev, C.byref(rp), C.byref(r), None) == -1


def test_comm_single_rank_reduce_is_identity(mbsv):
    import torch
    comm = mbsv.Comm(1, 0, 0, lambda ident: ident)
    t = torch.arange(100000, dtype=torch.int32, device="cuda:0")
    ms = comm.reduce_tallies(t.data_ptr(), t.numel(), 0)
    assert ms >= 0.0 and torch.equal(t.cpu(), torch.arange(100000, dtype=torch.int32))
    comm.close()


def test_split_chains_matches_the_host_helper(mbsv):
    from multibreak_sv_b200 import sharding
    for n, k in ((64, 8), (10, 4), (3, 2), (7, 7)):
        parts = [sharding.split_chains(n, k, r) for r in range(k)]
        assert sum(c for _, c in parts) == n and parts[0][0] == 0
        assert all(parts[r + 1][0] == parts[r][0] + parts[r][1] for r in range(k - 1))
