"""The speculative-window kernel (giant_kernel.cuh: one CTA per chain, windows of consecutive iterations evaluated at
once) against the CPU oracle: every output bit-exact, like the lane-per-chain kernels it replaces for few chains on a
large state (BASELINE configs[4], the giant connected component)."""
import os

import numpy as np
import pytest

from test_gpu_parity import _with_experiments, assert_same

pytestmark = pytest.mark.gpu


def used_giant(dp):
    """The speculative-window kernel is the only global-workspace launch with dynamic shared memory."""
    return any(li["rows"] == 0 and li["smem_bytes"] > 0 for li in dp.launch_info())


@pytest.fixture
def all_global(monkeypatch):
    """No shared-memory size classes: every subproblem of the batch goes to the global-workspace launch."""
    monkeypatch.setenv("MBSV_SMEM_CLASSES", "0")


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
@pytest.mark.parametrize("nfrag,chains,iters,perr,seed", [(3000, 3, 6000, 0.001, 1), (20000, 2, 30000, 0.001, 2),
                                                          (60, 4, 3000, 0.001, 3), (700, 5, 4000, 0.3, 4),
                                                          (2500, 1, 700, 0.6, 5), (40000, 2, 20000, 0.001, 6)])
def test_giant_component_matches_oracle(oracle, mbsv, all_global, monkeypatch, nfrag, chains, iters, perr, seed, cluster):
    """cluster = CTAs per chain (a thread-block cluster shares one chain's windows through distributed shared memory);
    forced here also where the planner would not pick it: on small states most windows are cut by stale reads."""
    from multibreak_sv_b200 import synth
    monkeypatch.setenv("MBSV_GIANT_CLUSTER", str(cluster))
    p = synth.generate("cfg5", n_frag=nfrag, seed=seed, shuffle_orders=True)
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", iters, n_chains=chains, perr=perr, nthreads=8, java_int_wrap=False)
    g = dp.run(mbsv.MODE_FAST, iters, n_chains=chains, perr=perr, memo_states=0)
    assert used_giant(dp)
    assert_same(o, g)
    assert int(g.counters[..., 2].sum()) > 0 or nfrag < 100          # AlterSV proposals went through the windows too


@pytest.mark.parametrize("cluster", [1, 4])
@pytest.mark.parametrize("iters", [0, 1, 2, 3, 5, 8, 13, 40, 157, 158, 159, 1000])
def test_giant_run_lengths(oracle, mbsv, all_global, monkeypatch, iters, cluster):
    """Runs that end inside the first window, exactly at a window's last iteration, and after several windows."""
    from multibreak_sv_b200 import synth
    monkeypatch.setenv("MBSV_GIANT_CLUSTER", str(cluster))
    p = synth.generate("cfg5", n_frag=1500, seed=11)
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", iters, n_chains=3, nthreads=8, java_int_wrap=False)
    g = dp.run(mbsv.MODE_FAST, iters, n_chains=3, memo_states=0)
    assert used_giant(dp)
    assert_same(o, g)


def test_giant_host_supplied_initial_state_and_int_wrap(oracle, mbsv, all_global):
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg5", n_frag=4000, seed=12)
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    rng = np.random.default_rng(3)
    nal = np.diff(p.arrays["frag_aln_ptr"])
    init = (rng.integers(0, 1 << 30, len(nal)) % (nal + 1) - 1).astype(np.int32)
    o = oracle.run(op, "incremental", 5000, n_chains=2, init_assign=init, nthreads=8, java_int_wrap=True)
    g = dp.run(mbsv.MODE_FAST, 5000, n_chains=2, init_assign=init, java_int_wrap=True, memo_states=0)
    assert used_giant(dp)
    assert_same(o, g)
    assert np.array_equal(g.initial_assign[0], init)


@pytest.mark.parametrize("n_exp,cluster", [(1, 1), (3, 1), (4, 1), (3, 2), (4, 8)])
def test_giant_experiment_counts(oracle, mbsv, all_global, monkeypatch, n_exp, cluster):
    from multibreak_sv_b200 import synth
    monkeypatch.setenv("MBSV_GIANT_CLUSTER", str(cluster))
    p = _with_experiments(mbsv, synth.generate("cfg5", n_frag=2500, seed=13), n_exp, 5)
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", 4000, n_chains=2, nthreads=8, java_int_wrap=False)
    g = dp.run(mbsv.MODE_FAST, 4000, n_chains=2, memo_states=0)
    assert used_giant(dp)
    assert_same(o, g)


def test_giant_next_to_small_subproblems(oracle, mbsv):
    """A mixed batch: the large subproblems of a few-chain batch take the speculative-window kernel, the small ones their
    shared-memory classes, in one mbsv_run."""
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=30000, seed=14, shuffle_orders=True)
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", 3000, n_chains=2, nthreads=8, java_int_wrap=False)
    g = dp.run(mbsv.MODE_FAST, 3000, n_chains=2)
    assert used_giant(dp)
    assert_same(o, g)


def test_giant_equals_lane_per_chain_kernel(mbsv, monkeypatch):
    """200k fragments: the speculative-window kernel and the one-chain-per-warp kernel it replaces agree on every output."""
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg5", n_frag=200_000, seed=15)
    dp = mbsv.DeviceProblem(p)
    a = dp.run(mbsv.MODE_FAST, 100_000, n_chains=4)
    assert used_giant(dp)
    monkeypatch.setenv("MBSV_GIANT_MAX_CHAINS", "0")
    b = dp.run(mbsv.MODE_FAST, 100_000, n_chains=4)
    assert not used_giant(dp)
    assert_same(a, b)
    assert a.kernel_ms < b.kernel_ms
