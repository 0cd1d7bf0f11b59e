"""Parity of the CUDA sampler (through the C ABI) against the CPU oracle.  Integer outputs — trajectory,
tallies, counters, RNG state, final state — must be bit-exact; log-probabilities must be bit-exact too because
both sides run the same individually-rounded fdlibm arithmetic."""
import hashlib
import os

import numpy as np
import pytest

from conftest import to_packed

pytestmark = pytest.mark.gpu

KAT6_SHA = "27aa9c30b1a4ac191a2942d3a8930d0fc723a978ee800960c2a7ed375e6bfcc6"


def assert_same(o, g, trace=False, logprob_exact=True):
    assert np.array_equal(o.error, g.error)
    assert np.array_equal(o.counters, g.counters)
    assert np.array_equal(o.rng_state, g.rng_state)
    assert np.array_equal(o.initial_assign, g.initial_assign)
    assert np.array_equal(o.final_assign, g.final_assign)
    assert np.array_equal(o.aln_count, g.aln_count)
    assert np.array_equal(o.frag_err_count, g.frag_err_count)
    assert np.array_equal(o.cluster_hist, g.cluster_hist)
    if logprob_exact:
        assert np.array_equal(o.initial_logprob, g.initial_logprob)
        assert np.array_equal(o.final_logprob, g.final_logprob)
    if trace:
        assert np.array_equal(o.trace_move_frag, g.trace_move_frag)
        assert np.array_equal(o.trace_move_assign, g.trace_move_assign)
        if logprob_exact:
            assert np.array_equal(o.trace_logprob, g.trace_logprob)


def test_example_replay_exact(oracle, mbsv, example_problem):
    """BASELINE config 1: the shipped example, seed 123456, 10,000 iterations, REPLAY_EXACT."""
    dp = mbsv.DeviceProblem(to_packed(mbsv, example_problem))
    g = dp.run(mbsv.MODE_REPLAY_EXACT, 10000, trace=True)
    o = oracle.run(example_problem, "faithful", 10000, trace=True)
    assert_same(o, g, trace=True)
    assert g.counters[0, 0].tolist() == [5006, 1174, 0, 0, 5006]
    assert int(g.rng_state[0, 0]) == 0x226A00D3B97D
    st = oracle.replay_states(example_problem, g.initial_assign[0], g.trace_move_frag[0, 0], g.trace_move_assign[0, 0])
    order = example_problem.names["sorted_frag"]
    h = hashlib.sha256()
    for row in st:
        h.update((",".join(str(row[i]) for i in order) + "\n").encode())
    assert h.hexdigest() == KAT6_SHA


def used_kernels(dp):
    """'memo' / 'smem' / 'global' per launch of the last run (mbsv_last_launch_info: negative rows = memoised kernel)."""
    return ["memo" if li["rows"] < 0 else "global" if li["rows"] == 0 else "smem" for li in dp.launch_info()]


def test_example_replay_exact_general_kernel(oracle, mbsv, example_problem):
    """BASELINE config 1 again with the memo disabled (memo_states = 0): the shipped example has 32,670 joint states, so
    with the default memo_states it runs in mbsv_memo_kernel; here mbsv_chain_kernel<REPLAY_EXACT, TRACE> and
    mbsv_chain_kernel<FAST> replay the same chain (MultiBreakSV.java:78,773-877): KAT-6 trajectory SHA, counters, LCG state."""
    dp = mbsv.DeviceProblem(to_packed(mbsv, example_problem))
    o = oracle.run(example_problem, "faithful", 10000, trace=True)
    for trace in (True, False):
        g = dp.run(mbsv.MODE_REPLAY_EXACT, 10000, trace=trace, memo_states=0)
        assert used_kernels(dp) == ["smem"]
        assert_same(o, g, trace=trace)
        assert g.counters[0, 0].tolist() == [5006, 1174, 0, 0, 5006]
        assert int(g.rng_state[0, 0]) == 0x226A00D3B97D
    g = dp.run(mbsv.MODE_REPLAY_EXACT, 10000, trace=True, memo_states=0)
    st = oracle.replay_states(example_problem, g.initial_assign[0], g.trace_move_frag[0, 0], g.trace_move_assign[0, 0])
    order = example_problem.names["sorted_frag"]
    h = hashlib.sha256()
    for row in st:
        h.update((",".join(str(row[i]) for i in order) + "\n").encode())
    assert h.hexdigest() == KAT6_SHA
    # FAST (Delta arithmetic, general kernel), with and without trace, one and many chains
    oi = oracle.run(example_problem, "incremental", 10000, trace=True, memo_states=0)
    for trace in (True, False):
        f = dp.run(mbsv.MODE_FAST, 10000, trace=trace, java_int_wrap=True, memo_states=0)
        assert used_kernels(dp) == ["smem"]
        assert_same(oi, f, trace=trace)
        assert f.counters[0, 0].tolist() == [5006, 1174, 0, 0, 5006] and int(f.rng_state[0, 0]) == 0x226A00D3B97D
    assert np.array_equal(oi.trace_move_frag, o.trace_move_frag)          # the Delta arithmetic takes the same decisions
    o64 = oracle.run(example_problem, "incremental", 3000, n_chains=64, memo_states=0)
    f64 = dp.run(mbsv.MODE_FAST, 3000, n_chains=64, java_int_wrap=True, memo_states=0)
    assert_same(o64, f64)


def test_second_shipped_data_set_general_kernel(oracle, mbsv):
    """Preprocessor-example (16 long reads, 10 clusters, AlterSV candidates) with the memo disabled: Naive and AlterSV
    proposals of real data through mbsv_chain_kernel in both modes."""
    from conftest import GOLDEN
    pre = os.path.join(GOLDEN, "preproc_infiles")
    prob = oracle.load_files(os.path.join(pre, "gasv.in.clusters"), os.path.join(pre, "assignments.txt"),
                             os.path.join(pre, "experiments.txt"))
    dp = mbsv.DeviceProblem(to_packed(mbsv, prob))
    for mode, om in ((mbsv.MODE_REPLAY_EXACT, "faithful"), (mbsv.MODE_FAST, "incremental")):
        o = oracle.run(prob, om, 10000, trace=True, memo_states=0)
        g = dp.run(mode, 10000, trace=True, java_int_wrap=True, memo_states=0)
        assert "memo" not in used_kernels(dp)
        assert_same(o, g, trace=True)
        assert int(g.counters[0, 0, 2]) > 100                             # AlterSV proposals
        g = dp.run(mode, 10000, java_int_wrap=True, memo_states=0)
        assert_same(o, g)


def test_example_fast_matches_incremental_and_replay(oracle, mbsv, example_problem):
    dp = mbsv.DeviceProblem(to_packed(mbsv, example_problem))
    g = dp.run(mbsv.MODE_FAST, 10000, trace=True, java_int_wrap=True)
    o = oracle.run(example_problem, "incremental", 10000, trace=True)
    assert_same(o, g, trace=True)
    r = dp.run(mbsv.MODE_REPLAY_EXACT, 10000, trace=True)
    assert np.array_equal(r.trace_move_frag, g.trace_move_frag)
    assert np.array_equal(r.aln_count, g.aln_count) and np.array_equal(r.cluster_hist, g.cluster_hist)


@pytest.mark.parametrize("iters", [1, 9, 10, 24, 25, 99, 1000])
def test_example_window_edges(oracle, mbsv, example_problem, iters):
    """burnin = N/10 (0 -> 10): windows longer than the run, empty windows, N < 25 (quirk Q10 does not apply)."""
    dp = mbsv.DeviceProblem(to_packed(mbsv, example_problem))
    for mode, om in ((mbsv.MODE_REPLAY_EXACT, "faithful"), (mbsv.MODE_FAST, "incremental")):
        g = dp.run(mode, iters, trace=True, java_int_wrap=True)
        o = oracle.run(example_problem, om, iters, trace=True)
        assert_same(o, g, trace=True)


def test_example_multichain_and_init(oracle, mbsv, example_problem):
    dp = mbsv.DeviceProblem(to_packed(mbsv, example_problem))
    g = dp.run(mbsv.MODE_FAST, 5000, n_chains=40, java_int_wrap=True)
    o = oracle.run(example_problem, "incremental", 5000, n_chains=40)
    assert_same(o, g)
    init = np.array([3, -1, 0, 2], np.int32)          # the --matchfile path: host-supplied initial state
    g = dp.run(mbsv.MODE_REPLAY_EXACT, 3000, init_assign=init, trace=True)
    o = oracle.run(example_problem, "faithful", 3000, init_assign=init, trace=True)
    assert_same(o, g, trace=True)
    assert np.array_equal(g.initial_assign[0], init)


@pytest.mark.parametrize("cfg,nfrag,chains,iters", [("cfg2", 3000, 1, 3000), ("cfg3", 1500, 4, 2000),
                                                    ("cfg4", 6000, 3, 2000), ("cfg4", 2500, 33, 1500)])
def test_synthetic_both_modes(oracle, mbsv, cfg, nfrag, chains, iters):
    """Synthetic batches of independent subproblems (all size classes incl. the global-workspace path)."""
    from multibreak_sv_b200 import synth
    p = synth.generate(cfg, n_frag=nfrag, shuffle_orders=True)
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", iters, n_chains=chains, nthreads=8, java_int_wrap=False)
    g = dp.run(mbsv.MODE_FAST, iters, n_chains=chains)
    assert_same(o, g)
    o = oracle.run(op, "faithful", iters, n_chains=chains, nthreads=8, trace=True)
    g = dp.run(mbsv.MODE_REPLAY_EXACT, iters, n_chains=chains, trace=True)
    assert_same(o, g, trace=True)
    assert g.launches >= 1


@pytest.mark.parametrize("cfg,nfrag,chains,iters", [("cfg4", 3000, 2, 2000), ("cfg3", 800, 33, 1200)])
def test_connected_components(oracle, mbsv, cfg, nfrag, chains, iters):
    """GASV connected components (localization -1): greedy set cover multisets on the device."""
    from multibreak_sv_b200 import synth
    p = synth.generate(cfg, n_frag=nfrag, shuffle_orders=True, conncomp_frac=0.5)
    assert int(p.arrays["cluster_kind"].sum()) > 10
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", iters, n_chains=chains, nthreads=8, java_int_wrap=False, trace=True)
    g = dp.run(mbsv.MODE_FAST, iters, n_chains=chains, trace=True)
    assert_same(o, g, trace=True)
    o = oracle.run(op, "faithful", iters, n_chains=chains, nthreads=8, trace=True)
    g = dp.run(mbsv.MODE_REPLAY_EXACT, iters, n_chains=chains, trace=True)
    assert_same(o, g, trace=True)
    g2 = dp.run(mbsv.MODE_FAST, iters, n_chains=chains)          # non-trace kernels (look-ahead lazy skip)
    assert np.array_equal(g2.cluster_hist, g.cluster_hist) and np.array_equal(g2.rng_state, g.rng_state)


def test_fast_trace_equals_replay_trace_synthetic(mbsv):
    """FAST and REPLAY_EXACT consume the same RNG stream and take the same decisions."""
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=4000, shuffle_orders=True, seed=99)
    dp = mbsv.DeviceProblem(p)
    a = dp.run(mbsv.MODE_FAST, 2000, n_chains=2, trace=True)
    b = dp.run(mbsv.MODE_REPLAY_EXACT, 2000, n_chains=2, trace=True, java_int_wrap=False)
    assert np.array_equal(a.trace_move_frag, b.trace_move_frag)
    assert np.array_equal(a.aln_count, b.aln_count)
    assert np.array_equal(a.rng_state, b.rng_state)


def test_device_math_matches_host(mbsv, oracle):
    """jm_log/jm_exp host build == oracle's restatement (bitwise) on a grid; the device build is exercised by the
    trajectory tests above (any ulp difference changes trace_logprob)."""
    L = mbsv.load_library()
    xs = np.concatenate([np.linspace(1e-6, 50, 4001), np.arange(1, 3000, dtype=float)])
    for x in xs:
        assert L.mbsv_log(float(x)) == oracle.jlog(float(x))
    for x in np.linspace(-60, 10, 5001):
        assert L.mbsv_exp(float(x)) == oracle.jexp(float(x))


def test_host_main_file_to_file(oracle, mbsv, example_problem, tmp_path):
    """MultiBreakSV.main on the shipped example through the native host mirror: input files -> GPU -> the
    reference's four output files, checked against the oracle's run of the same chain."""
    import os
    from conftest import EXAMPLE
    from multibreak_sv_b200 import host
    prefix = os.path.join(str(tmp_path), "out")
    sv = host.main(["--clusterfile", os.path.join(EXAMPLE, "gasv.clusters"), "--assignmentfile",
                    os.path.join(EXAMPLE, "assignments.txt"), "--experimentfile", os.path.join(EXAMPLE, "experiments.txt"),
                    "--numiterations", "10000", "--perr", "0.001", "--prefix", prefix])
    o = oracle.run(example_problem, "faithful", 10000, trace=True)
    # expected long-burn-in files: same writer fed with the oracle's tallies
    ref = host.MultiBreakSV(["--clusterfile", os.path.join(EXAMPLE, "gasv.clusters"), "--assignmentfile",
                             os.path.join(EXAMPLE, "assignments.txt"), "--experimentfile",
                             os.path.join(EXAMPLE, "experiments.txt"), "--numiterations", "10000", "--perr", "0.001",
                             "--prefix", prefix + "_ref"]).load()
    ref.setResults(o.aln_count, o.frag_err_count, o.cluster_hist)
    ref.writeFinalFiles()
    for ext in (".finalbreakpoints.longburnin", ".finalassignment.longburnin"):
        assert open(prefix + ext).read() == open(prefix + "_ref" + ext).read()
    # the C++ command-line driver (what replaces `java -jar MultiBreakSV.jar`) writes the same four files
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.abspath(host.__file__)), "mbsv_cli")
    r = subprocess.run([cli, "--clusterfile", os.path.join(EXAMPLE, "gasv.clusters"), "--assignmentfile",
                        os.path.join(EXAMPLE, "assignments.txt"), "--experimentfile", os.path.join(EXAMPLE, "experiments.txt"),
                        "--numiterations", "10000", "--perr", "0.001", "--prefix", prefix + "_cli"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "1174 of 5006(" + host.java_double_to_string(1174 / 5006) + ") Naive Moves" in r.stdout      # KAT-6, :842
    assert "0 of 0(NaN) AlterSV Moves" in r.stdout
    assert "After 10000 iterations, probability is " + host.java_double_to_string(float(o.final_logprob.ravel()[0])) in r.stdout
    for ext in (".finalbreakpoints.longburnin", ".finalassignment.longburnin", ".mcmc.txt", ".samplingprobs"):
        assert open(prefix + ext).read() == open(prefix + "_cli" + ext).read(), ext
    # mcmc.txt: Iteration, LogProb, NumBreakpoints (C*E every skip=10 iterations), MadeMove?
    lines = open(prefix + ".mcmc.txt").read().splitlines()
    assert lines[0] == "Iteration\tLogProb\tNumBreakpoints\tMadeMove?" and len(lines) == 10001
    for i in (0, 1, 2, 6, 21, 9999):
        it, lp, nb, mv = lines[1 + i].split("\t")
        assert int(it) == i and lp == host.java_double_to_string(o.trace_logprob[0, 0, i])
        assert int(nb) == (27 if i % 10 == 0 else 0) and int(mv) == int(o.trace_move_frag[0, 0, i] != -1)
    assert sum(int(l.split("\t")[3]) for l in lines[1:]) == 1174
    # .samplingprobs: distinct states, SampledProb = count / numiters
    st = oracle.replay_states(example_problem, o.initial_assign[0], o.trace_move_frag[0, 0], o.trace_move_assign[0, 0])
    order = example_problem.names["sorted_frag"]
    counts = {}
    for row in st:
        k = "\t".join(str(row[i]) for i in order)
        counts[k] = counts.get(k, 0) + 1
    sp = open(prefix + ".samplingprobs").read().splitlines()
    assert sp[0] == "SampledProb\tLogProb\tlongread_118669\tlongread_118671\tlongread_118673\tlongread_118674"
    got = {}
    for l in sp[1:]:
        f = l.split("\t")
        got["\t".join(f[2:])] = f[0]
    assert set(got) == set(counts)
    assert all(got[k] == host.java_double_to_string(counts[k] / 10000) for k in counts)
    # HashMap iteration order of the state strings
    assert ["\t".join(l.split("\t")[2:]) for l in sp[1:]] == oracle.hashmap_order(list(dict.fromkeys(
        "\t".join(str(row[i]) for i in order) for row in st)))[0]


def _exact_posterior(oracle, p):
    """Exact marginals of a tiny single-subproblem PackedProblem by enumeration: per fragment P(assignment), per
    cluster the distribution of the support value (sum over experiments of selected-ESP counts listed separately)."""
    op = oracle.Problem(p.scalars, p.arrays)
    lp = oracle.enumerate_logprobs(op, 0, 0.05)
    w = np.exp(lp - lp.max()); w /= w.sum()
    a = p.arrays
    nal = np.diff(a["frag_aln_ptr"])
    F, C, E = p.scalars["n_frag"], p.scalars["n_cluster"], p.scalars["n_exp"]
    frag = [np.zeros(n + 1) for n in nal]
    hist = np.zeros(p.scalars["n_hist"])
    for s in range(len(w)):
        t = s
        cnt = np.zeros((C, E), int)
        for f in range(F):
            d = t % (nal[f] + 1); t //= (nal[f] + 1)
            frag[f][d] += w[s]
            if d > 0:
                al = a["frag_aln_ptr"][f] + d - 1
                for j in range(a["aln_esp_ptr"][al], a["aln_esp_ptr"][al + 1]):
                    if a["aln_esp_cluster"][j] >= 0:
                        cnt[a["aln_esp_cluster"][j], a["aln_exp"][al]] += 1
        for c in range(C):
            for x in range(E):
                if cnt[c, x] > 0:
                    hist[a["cluster_hist_ptr"][c] + cnt[c, x]] += w[s]
    return frag, hist


def test_multichain_posterior_within_monte_carlo_error(oracle, mbsv):
    """Statistical tier: pooled multi-chain posteriors (fragment assignments and cluster-support histograms) agree
    with the exactly enumerated posterior within 4 Monte-Carlo standard errors estimated from independent batches.
    Fragments have <= 2 alignments so that the reference's biased 'other n-1' scan (quirk Q1) cannot bite."""
    from multibreak_sv_b200 import synth
    p = None
    for seed in range(300):
        q = synth.generate("cfg4", n_frag=6, seed=5000 + seed)
        nal = np.diff(q.arrays["frag_aln_ptr"])
        if q.scalars["n_sub"] == 1 and nal.max() == 2 and q.scalars["n_sv"] > 0:
            p = q
            break
    assert p is not None
    frag, hist = _exact_posterior(oracle, p)
    dp = mbsv.DeviceProblem(p)
    B, X, N = 12, 256, 6000
    burn = N - 2000                                   # window = the last 1999 iterations
    est_f, est_h = [], []
    fp = p.arrays["frag_aln_ptr"]
    for b in range(B):
        g = dp.run(mbsv.MODE_FAST, N, n_chains=X, perr=0.05, seed=1000 * (b + 1), burnin=N - burn, per_chain=False,
                   want_state=False)
        nwin = (N - max(N - (N - burn) + 1, 0)) * X
        est_f.append(np.concatenate([np.concatenate([[g.frag_err_count[f]], g.aln_count[fp[f]:fp[f + 1]]]) for f in range(6)]) / nwin)
        est_h.append(g.cluster_hist / nwin)
    for est, exact in ((np.array(est_f), np.concatenate(frag)), (np.array(est_h), hist)):
        mean, se = est.mean(0), est.std(0, ddof=1) / np.sqrt(B)
        z = np.abs(mean - exact) / np.maximum(se, 1e-4)
        assert z.max() < 4.5, (z.max(), mean[z.argmax()], exact[z.argmax()])


def test_full_size_invariants(mbsv):
    """BASELINE-size batch (mixed config, 64 chains): properties that do not need the oracle.
    (1) every fragment's tallies sum to window x chains; (2) every (cluster, experiment) contributes at most one bin
    per sample, so a cluster's histogram mass is <= E x window x chains; (3) proposals + lazy iterations = iterations,
    with the lazy share near 1/2; (4) a second run reproduces the first bit for bit (no race in the atomics' sums);
    (5) REPLAY_EXACT and FAST agree on every integer output of a sub-batch."""
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=2_000_000)
    dp = mbsv.DeviceProblem(p)
    X, N = 64, 400
    g = dp.run(mbsv.MODE_FAST, N, n_chains=X, want_state=False, per_chain=False)
    burnin = mbsv.reference_burnin(N)
    nwin = N - max(N - burnin + 1, 0)
    a = p.arrays
    per_frag = np.add.reduceat(g.aln_count.astype(np.int64), a["frag_aln_ptr"][:-1]) + g.frag_err_count
    assert (per_frag == nwin * X).all()
    hp = a["cluster_hist_ptr"]
    mass = np.add.reduceat(g.cluster_hist.astype(np.int64), hp[:-1])
    assert (g.cluster_hist[hp[:-1]] == 0).all()                         # bin 0 is derived by the host
    assert (mass <= p.scalars["n_exp"] * nwin * X).all() and mass.sum() > 0
    chains = p.scalars["n_sub"] * X
    props = int(g.totals[0] + g.totals[2])
    assert abs(props / (chains * N) - 0.5) < 0.002 and int(g.totals[5]) == 0
    assert int(g.totals[1]) <= int(g.totals[0]) and int(g.totals[4]) >= props - int(g.totals[2])
    g2 = dp.run(mbsv.MODE_FAST, N, n_chains=X, want_state=False, per_chain=False)
    assert np.array_equal(g.aln_count, g2.aln_count) and np.array_equal(g.cluster_hist, g2.cluster_hist)
    assert np.array_equal(g.totals, g2.totals)
    small = synth.generate("cfg4", n_frag=60_000, seed=4)
    ds = mbsv.DeviceProblem(small)
    f = ds.run(mbsv.MODE_FAST, N, n_chains=X)
    r = ds.run(mbsv.MODE_REPLAY_EXACT, N, n_chains=X, java_int_wrap=False)
    for k in ("aln_count", "frag_err_count", "cluster_hist", "final_assign", "counters", "rng_state"):
        assert np.array_equal(getattr(f, k), getattr(r, k)), k


@pytest.mark.parametrize("n_chains", [1, 3, 8, 16, 24, 48])
def test_memo_kernel_with_partial_warps(oracle, mbsv, n_chains):
    """Chain counts that are not a multiple of 32 (and small batches) run the memoised kernel with mixed warps: every lane
    its own subproblem, strides in the lane's column, tables in device memory."""
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg3", n_frag=1200, seed=21)
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", 1500, n_chains=n_chains, nthreads=8, java_int_wrap=False, trace=True)
    g = dp.run(mbsv.MODE_FAST, 1500, n_chains=n_chains, trace=True)
    assert_same(o, g, trace=True)
    g = dp.run(mbsv.MODE_FAST, 1500, n_chains=n_chains)
    assert np.array_equal(o.aln_count, g.aln_count) and np.array_equal(o.rng_state, g.rng_state)
    assert np.array_equal(o.final_logprob, g.final_logprob)


@pytest.mark.parametrize("n_chains,mode_rows", [(32, -1), (5, -1), (64, -1)])
def test_memoised_subproblem_wider_than_shared_memory(oracle, mbsv, n_chains, mode_rows):
    """A one-fragment subproblem with 300 alignments in 300 clusters is memo-eligible (301 states) but its counts do not
    fit the shared-memory classes: it must run (general kernel, same full-sum arithmetic) next to small memoised ones."""
    from multibreak_sv_b200 import synth
    g = synth.generate("cfg2", n_frag=40, seed=5)
    a, sc = {k: v.copy() for k, v in g.arrays.items()}, dict(g.scalars)
    nA, nC = 300, 300
    F0, A0, K0, C0, S0 = sc["n_frag"], sc["n_aln"], sc["n_aln_esp"], sc["n_cluster"], sc["n_sub"]
    cat = np.concatenate
    a["sub_frag_ptr"] = cat([a["sub_frag_ptr"], [F0 + 1]]).astype(np.int32)
    a["sub_cluster_ptr"] = cat([a["sub_cluster_ptr"], [C0 + nC]]).astype(np.int32)
    a["sub_sv_ptr"] = cat([a["sub_sv_ptr"], [a["sub_sv_ptr"][-1]]]).astype(np.int32)
    a["frag_aln_ptr"] = cat([a["frag_aln_ptr"], [A0 + nA]]).astype(np.int32)
    rng = np.random.default_rng(8)
    a["aln_numerrors"] = cat([a["aln_numerrors"], rng.integers(0, 6, nA)]).astype(np.int32)
    a["aln_len"] = cat([a["aln_len"], np.full(nA, 200)]).astype(np.int32)
    a["aln_exp"] = cat([a["aln_exp"], np.zeros(nA)]).astype(np.int32)
    a["aln_esp_ptr"] = cat([a["aln_esp_ptr"], K0 + 1 + np.arange(nA)]).astype(np.int32)
    a["aln_esp_cluster"] = cat([a["aln_esp_cluster"], C0 + np.arange(nA)]).astype(np.int32)
    a["aln_esp_leaf"] = cat([a["aln_esp_leaf"], np.zeros(nA)]).astype(np.int32)
    a["cluster_kind"] = cat([a["cluster_kind"], np.zeros(nC)]).astype(np.uint8)
    a["supports_order"] = cat([a["supports_order"], C0 + rng.permutation(nC)]).astype(np.int32)
    one = 1 + np.arange(nC)
    a["cluster_hist_ptr"] = cat([a["cluster_hist_ptr"], a["cluster_hist_ptr"][-1] + 2 * one]).astype(np.int32)   # supports 0..1
    a["cluster_leaf_ptr"] = cat([a["cluster_leaf_ptr"], a["cluster_leaf_ptr"][-1] + one]).astype(np.int32)
    a["leaf_owner"] = cat([a["leaf_owner"], np.full(nC, F0)]).astype(np.int32)
    a["cluster_root_ptr"] = cat([a["cluster_root_ptr"], a["cluster_root_ptr"][-1] + one]).astype(np.int32)
    a["root_leaf_ptr"] = cat([a["root_leaf_ptr"], a["root_leaf_ptr"][-1] + one]).astype(np.int32)
    a["root_leaf"] = cat([a["root_leaf"], np.zeros(nC)]).astype(np.int32)
    sc.update(n_sub=S0 + 1, n_frag=F0 + 1, n_aln=A0 + nA, n_aln_esp=K0 + nA, n_cluster=C0 + nC,
              n_hist=int(a["cluster_hist_ptr"][-1]), n_leaf=sc["n_leaf"] + nC, n_root=sc["n_root"] + nC,
              n_root_leaf=sc["n_root_leaf"] + nC)
    p = mbsv.PackedProblem(sc, a)
    op = oracle.Problem(sc, a)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", 1200, n_chains=n_chains, nthreads=8, java_int_wrap=False)
    gr = dp.run(mbsv.MODE_FAST, 1200, n_chains=n_chains)
    assert_same(o, gr)
    rows = [li["rows"] for li in dp.launch_info()]
    assert 0 in rows                                      # the wide subproblem went to the global-workspace class


@pytest.mark.parametrize("n_chains", [1, 32, 64])
def test_tiny_iteration_counts_without_trace(oracle, mbsv, example_problem, n_chains):
    """The look-ahead lazy skip (non-trace kernels) at the end of a chain: every small N, generic and memoised kernels."""
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=300, seed=12)
    op = oracle.Problem(p.scalars, p.arrays)
    dp, de = mbsv.DeviceProblem(p), mbsv.DeviceProblem(to_packed(mbsv, example_problem))
    for n in (0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 16, 33):
        for prob, oprob in ((dp, op), (de, example_problem)):
            g = prob.run(mbsv.MODE_FAST, n, n_chains=n_chains)
            o = oracle.run(oprob, "incremental", n, n_chains=n_chains, java_int_wrap=False)
            assert_same(o, g)


@pytest.mark.gpu
def test_bounds_checked_build():
    """The parity cases again on libmbsv_b200_check.so (`make check`): every access to a chain's state tile and the
    memo tables' extent are range-checked on the device and trap instead of corrupting a neighbouring chain.  (This
    pool runs no compute-sanitizer; these are the kernels' own checks.)"""
    import subprocess
    import sys
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multibreak_sv_b200",
                       "libmbsv_b200_check.so")
    assert os.path.exists(lib), "libmbsv_b200_check.so missing: run `make -C multibreak_sv_b200/csrc check`"
    env = dict(os.environ, MBSV_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "not bounds_checked and not full_size and not posterior_within"],
                       env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def _with_experiments(mbsv, g, n_exp, seed):
    """The generated problem with its alignments spread over n_exp sequencing experiments (1..4 on the device)."""
    import ctypes as C_
    a, sc = {k: v.copy() for k, v in g.arrays.items()}, dict(g.scalars)
    rng = np.random.default_rng(seed)
    a["aln_exp"] = rng.integers(0, n_exp, sc["n_aln"]).astype(np.int32)
    a["exp_pseq"] = np.array([0.01, 0.15, 0.05, 0.10][:n_exp])
    a["exp_lambda"] = np.array([10.0, 3.0, 5.0, 2.0][:n_exp])
    T = np.zeros((n_exp, sc["max_support"] + 1))
    L = mbsv.load_library()
    for x in range(n_exp):
        L.mbsv_fill_logprobcov(float(a["exp_lambda"][x]), sc["max_support"], T[x].ctypes.data_as(C_.POINTER(C_.c_double)))
    a["logprobcov"] = T.ravel()
    sc["n_exp"] = n_exp
    return mbsv.PackedProblem(sc, a)


@pytest.mark.parametrize("n_exp,cc", [(3, 0.0), (4, 0.0), (4, 0.4)])
def test_three_and_four_experiments(oracle, mbsv, n_exp, cc):
    """MBSV_MAX_EXP = 4: per-experiment totals, logBinomial terms and count slots for every experiment count."""
    from multibreak_sv_b200 import synth
    p = _with_experiments(mbsv, synth.generate("cfg4", n_frag=1500, seed=31, conncomp_frac=cc), n_exp, 7)
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    for mode, omode in ((mbsv.MODE_FAST, "incremental"), (mbsv.MODE_REPLAY_EXACT, "faithful")):
        o = oracle.run(op, omode, 1500, n_chains=32, nthreads=8, java_int_wrap=False, trace=True)
        g = dp.run(mode, 1500, n_chains=32, java_int_wrap=False, trace=True)
        assert_same(o, g, trace=True)
        g = dp.run(mode, 1500, n_chains=32, java_int_wrap=False)
        assert_same(o, g)


@pytest.mark.parametrize("perr", [0.0, 1.0, 0.5, 1e-300])
def test_error_probability_extremes(oracle, mbsv, example_problem, perr):
    """perr = 0 makes log(perr) = -inf (0 * -inf = NaN log-posteriors, never accepted: Math.min(1, NaN) > u is false);
    perr = 1 makes every draw propose "unmapped"; both must follow the reference's arithmetic, not trap."""
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=400, seed=77)
    op = oracle.Problem(p.scalars, p.arrays)
    for prob, oprob in ((mbsv.DeviceProblem(p), op), (mbsv.DeviceProblem(to_packed(mbsv, example_problem)), example_problem)):
        for mode, omode in ((mbsv.MODE_FAST, "incremental"), (mbsv.MODE_REPLAY_EXACT, "faithful")):
            o = oracle.run(oprob, omode, 800, n_chains=32, perr=perr, nthreads=8, java_int_wrap=False)
            g = prob.run(mode, 800, n_chains=32, perr=perr, java_int_wrap=False)
            assert np.array_equal(o.counters, g.counters) and np.array_equal(o.rng_state, g.rng_state)
            assert np.array_equal(o.final_assign, g.final_assign) and np.array_equal(o.aln_count, g.aln_count)
            assert np.array_equal(o.cluster_hist, g.cluster_hist) and np.array_equal(o.frag_err_count, g.frag_err_count)
            # NaN log-posteriors compare by bit pattern
            assert np.array_equal(o.final_logprob.view(np.uint64), g.final_logprob.view(np.uint64))


def test_partitioned_files_to_gpu(oracle, mbsv, tmp_path):
    """Reference-format files -> mbsv_host_load_partitioned (in-memory partition + per-subset packing) -> GPU batch; the
    oracle runs the same packed batch."""
    from multibreak_sv_b200 import host, synth
    g = synth.generate("cfg4", n_frag=900, seed=123, conncomp_frac=0.2)
    paths = synth.write_reference_files(g, str(tmp_path), frag_numbers=np.random.default_rng(9).permutation(900) * 5 + 77)
    p, skipped = host.load_partitioned(paths["clusters"], paths["assignments"], paths["experiments"])
    assert skipped == 0 and p.scalars["n_frag"] == 900 and p.scalars["n_sub"] >= g.scalars["n_sub"]
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", 2000, n_chains=32, nthreads=8, java_int_wrap=False)
    assert_same(o, dp.run(mbsv.MODE_FAST, 2000, n_chains=32))
    o = oracle.run(op, "faithful", 500, n_chains=2, nthreads=8, java_int_wrap=True, trace=True)
    assert_same(o, dp.run(mbsv.MODE_REPLAY_EXACT, 500, n_chains=2, trace=True), trace=True)


@pytest.mark.parametrize("max_aln", [2, 9])
def test_gibbs_mode_samples_the_exact_posterior(oracle, mbsv, max_aln):
    """MBSV_MODE_GIBBS (statistical tier): pooled heat-bath chains against the exactly enumerated posterior,
    within 4.5 Monte-Carlo standard errors from independent batches.  Unlike the reference's MH chain the Gibbs kernel has
    no proposal asymmetry (quirk Q1), so fragments with many alignments are fair game too (max_aln = 9)."""
    from multibreak_sv_b200 import synth
    p = None
    for seed in range(600):
        q = synth.generate("cfg4", n_frag=5, seed=9000 + seed)
        nal = np.diff(q.arrays["frag_aln_ptr"])
        if q.scalars["n_sub"] == 1 and nal.max() == max_aln and np.prod(nal + 1.0) < 3000:
            p = q
            break
    assert p is not None
    frag, hist = _exact_posterior(oracle, p)
    dp = mbsv.DeviceProblem(p)
    # single-site updates cross between the modes of coupled fragments slowly: a long burn-in (measured: 1000 iterations
    # leave a visible trace of the random initial state, 10000 do not)
    B, X, N = 12, 256, 20000
    win = 10000
    est_f, est_h = [], []
    fp = p.arrays["frag_aln_ptr"]
    F = p.scalars["n_frag"]
    for b in range(B):
        g = dp.run(mbsv.MODE_GIBBS, N, n_chains=X, perr=0.05, seed=777 * (b + 1), burnin=win, per_chain=False, want_state=False)
        assert int(g.totals[0]) == N * X and int(g.totals[5]) == 0 and 0 < int(g.totals[1]) < N * X
        nwin = (N - max(N - win + 1, 0)) * X
        est_f.append(np.concatenate([np.concatenate([[g.frag_err_count[f]], g.aln_count[fp[f]:fp[f + 1]]]) for f in range(F)]) / nwin)
        est_h.append(g.cluster_hist / nwin)
        # every sample of the window has every fragment in exactly one state
        assert all(int(g.frag_err_count[f]) + int(g.aln_count[fp[f]:fp[f + 1]].sum()) == nwin for f in range(F))
    for est, exact in ((np.array(est_f), np.concatenate(frag)), (np.array(est_h), hist)):
        mean, se = est.mean(0), est.std(0, ddof=1) / np.sqrt(B)
        z = np.abs(mean - exact) / np.maximum(se, 1e-4)
        assert z.max() < 4.5, (z.max(), mean[z.argmax()], exact[z.argmax()])


def test_gibbs_mode_limits_and_determinism(mbsv, example_problem):
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=2000, seed=3)
    dp = mbsv.DeviceProblem(p)
    a = dp.run(mbsv.MODE_GIBBS, 500, n_chains=8)
    b = dp.run(mbsv.MODE_GIBBS, 500, n_chains=8)
    for k in ("aln_count", "frag_err_count", "cluster_hist", "final_assign", "counters", "rng_state", "final_logprob"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert (a.counters[:, :, 0] == 500).all() and (a.counters[:, :, 1] <= 500).all()
    # 4, 8 or 32 lanes per chain (chosen per size class) are the same chains: every lane of a group carries the same RNG
    for group in ("4", "8", "32"):
        os.environ["MBSV_GIBBS_GROUP"] = group
        try:
            c = dp.run(mbsv.MODE_GIBBS, 500, n_chains=8)
        finally:
            del os.environ["MBSV_GIBBS_GROUP"]
        for k in ("aln_count", "frag_err_count", "cluster_hist", "final_assign", "counters", "rng_state"):
            assert np.array_equal(getattr(a, k), getattr(c, k)), (group, k)
        assert np.allclose(a.final_logprob, c.final_logprob, rtol=1e-12, atol=1e-9)
    # the state in the global workspace (the path of subproblems beyond the shared-memory classes) is the same chain again
    os.environ["MBSV_GIBBS_GLOBAL"] = "1"
    try:
        c = dp.run(mbsv.MODE_GIBBS, 500, n_chains=8)
        assert [li["rows"] for li in dp.launch_info()] == [0]
    finally:
        del os.environ["MBSV_GIBBS_GLOBAL"]
    for k in ("aln_count", "frag_err_count", "cluster_hist", "final_assign", "counters", "rng_state"):
        assert np.array_equal(getattr(a, k), getattr(c, k)), ("global", k)
    assert np.allclose(a.final_logprob, c.final_logprob, rtol=1e-12, atol=1e-9)
    # a component too large for shared memory (a giant-component shape: 6000 fragments, ~18 000 state rows) runs there by itself
    big = mbsv.DeviceProblem(synth.generate("cfg5", n_frag=6000, seed=4))
    gb = big.run(mbsv.MODE_GIBBS, 3000, n_chains=4)
    assert [li["rows"] for li in big.launch_info()] == [0] and int(gb.totals[0]) == 3000 * 4
    gb2 = big.run(mbsv.MODE_GIBBS, 3000, n_chains=4)
    assert np.array_equal(gb.aln_count, gb2.aln_count) and np.array_equal(gb.final_assign, gb2.final_assign)
    nwin = 3000 - (3000 - mbsv.reference_burnin(3000) + 1)
    assert nwin > 0 and (gb.aln_count.sum() + gb.frag_err_count.sum()) == 4 * nwin * 6000     # every fragment is somewhere, every window iteration
    # the shipped example has a fragment with 32 alignments: 33 options, two lane chunks
    de = mbsv.DeviceProblem(to_packed(mbsv, example_problem))
    g = de.run(mbsv.MODE_GIBBS, 4000, n_chains=64)
    assert int(g.totals[0]) == 4000 * 64 and np.isfinite(g.final_logprob).all()
    with pytest.raises(mbsv.MbsvError) as e:
        de.run(mbsv.MODE_GIBBS, 10, trace=True)
    assert e.value.status == -1
    # connected components run (their set cover is evaluated per option); deterministic, every fragment tallied once per sample
    pc = synth.generate("cfg4", n_frag=200, seed=3, conncomp_frac=0.5)
    cc = mbsv.DeviceProblem(pc)
    c1, c2 = cc.run(mbsv.MODE_GIBBS, 400, n_chains=8), cc.run(mbsv.MODE_GIBBS, 400, n_chains=8)
    for k in ("aln_count", "frag_err_count", "cluster_hist", "final_assign", "rng_state"):
        assert np.array_equal(getattr(c1, k), getattr(c2, k)), k
    fp = pc.arrays["frag_aln_ptr"]
    nwin = (400 - (400 - mbsv.reference_burnin(400) + 1)) * 8
    assert all(int(c1.frag_err_count[f]) + int(c1.aln_count[fp[f]:fp[f + 1]].sum()) == nwin for f in range(200))


@pytest.mark.gpu
def test_gibbs_mode_with_connected_components(oracle, mbsv):
    """Gibbs mode on a tiny subproblem with a GASV connected component: the fragment marginals agree with the enumerated
    posterior, and the support histograms (set-cover multisets for the component) with the MH chain's, whose connected-
    component path is compared bit for bit with the oracle elsewhere.  Fragments have <= 2 alignments (quirk Q1 cannot bite)."""
    from multibreak_sv_b200 import synth
    p = None
    for seed in range(400):
        q = synth.generate("cfg4", n_frag=6, seed=7000 + seed, conncomp_frac=0.7)
        if q.scalars["n_sub"] == 1 and np.diff(q.arrays["frag_aln_ptr"]).max() == 2 and (q.arrays["cluster_kind"] != 0).any():
            p = q
            break
    assert p is not None
    frag, _ = _exact_posterior(oracle, p)
    dp = mbsv.DeviceProblem(p)
    B, X, N, win = 10, 256, 20000, 10000
    fp, F = p.arrays["frag_aln_ptr"], p.scalars["n_frag"]
    est = {mbsv.MODE_GIBBS: ([], []), mbsv.MODE_FAST: ([], [])}
    for mode in est:
        for b in range(B):
            g = dp.run(mode, N, n_chains=X, perr=0.05, seed=555 * (b + 1), burnin=win, per_chain=False, want_state=False)
            assert int(g.totals[5]) == 0
            nwin = (N - max(N - win + 1, 0)) * X
            est[mode][0].append(np.concatenate([np.concatenate([[g.frag_err_count[f]], g.aln_count[fp[f]:fp[f + 1]]]) for f in range(F)]) / nwin)
            est[mode][1].append(g.cluster_hist / nwin)
    gf, gh = (np.array(v) for v in est[mbsv.MODE_GIBBS])
    mf, mh = (np.array(v) for v in est[mbsv.MODE_FAST])
    z = np.abs(gf.mean(0) - np.concatenate(frag)) / np.maximum(gf.std(0, ddof=1) / np.sqrt(B), 1e-4)
    assert z.max() < 4.5, ("fragment marginals", z.max())
    se = np.sqrt(gh.var(0, ddof=1) / B + mh.var(0, ddof=1) / B)
    z = np.abs(gh.mean(0) - mh.mean(0)) / np.maximum(se, 1e-4)
    assert z.max() < 4.5, ("support histograms vs the MH chain", z.max())
    assert gh.mean(0)[p.arrays["cluster_hist_ptr"][int(np.nonzero(p.arrays["cluster_kind"])[0][0])]:].sum() > 0


def test_rhat_across_chain_groups(mbsv):
    """diagnostics.rhat on the device: groups of Gibbs chains agree (R-hat ~ 1) after a long run and visibly disagree when
    the run is far too short for the random initial states to be forgotten."""
    from multibreak_sv_b200 import diagnostics, synth
    dp = mbsv.DeviceProblem(synth.generate("cfg4", n_frag=300, seed=8))
    good = diagnostics.rhat(dp, mbsv.MODE_GIBBS, 20000, chains_per_group=32, groups=4, burnin=10000, perr=0.05)
    assert good["samples_per_group"] == 9999 * 32 and 0.99 < good["max"] < 1.05
    mh = diagnostics.rhat(dp, mbsv.MODE_FAST, 20000, chains_per_group=32, groups=4, burnin=10000, perr=0.05)
    assert 0.99 < mh["max"] < 1.1
    short = diagnostics.rhat(dp, mbsv.MODE_GIBBS, 12, chains_per_group=2, groups=4, burnin=10, perr=0.05)
    assert short["max"] > 1.2


def _many_slot_problem(mbsv, seed):
    """Subproblems whose alignments carry 1..7 ESPs each (several of them in one cluster now and then): more counted slots
    than the 16-byte slot record of an alignment holds (aln_ent4 marks those and the kernel walks aln_esp_slot), and the
    same slot twice in one record.  Sizes chosen to land in the 48-row, the 96-row and the global-workspace class."""
    from multibreak_sv_b200 import synth
    rng = np.random.default_rng(seed)
    base = synth.generate("cfg4", n_frag=50, seed=3)
    E, pseq, lam = 2, np.array([0.01, 0.15]), np.array([10.0, 3.0])
    sub_frag_ptr, sub_cluster_ptr, frag_aln_ptr, aln_esp_ptr = [0], [0], [0], [0]
    numerr, alen, aexp, esp_cluster, esp_leaf, sup_order = [], [], [], [], [], []
    leaves = []                                               # per cluster: owning fragment of every leaf
    for F_s, C_s in ((14, 12), (30, 20), (60, 50)):
        f0, c0 = sub_frag_ptr[-1], sub_cluster_ptr[-1]
        leaves += [[] for _ in range(C_s)]
        spread = list(rng.permutation(C_s))                   # every cluster gets at least one leaf
        for f in range(F_s):
            x = int(rng.integers(0, E))
            for _ in range(int(rng.integers(2, 4))):
                numerr.append(int(rng.integers(0, 9))); alen.append(int(rng.integers(200, 3000))); aexp.append(x)
                for _ in range(int(rng.integers(1, 8))):
                    c = c0 + (int(spread.pop()) if spread else int(rng.integers(0, C_s)))
                    esp_cluster.append(c); esp_leaf.append(len(leaves[c])); leaves[c].append(f0 + f)
                aln_esp_ptr.append(len(esp_cluster))
            frag_aln_ptr.append(len(numerr))
        sub_frag_ptr.append(f0 + F_s); sub_cluster_ptr.append(c0 + C_s)
        sup_order += list(c0 + rng.permutation(C_s))
    C, nl = len(leaves), np.array([len(l) for l in leaves])
    assert nl.min() >= 1
    max_support = int(nl.max())
    cum = lambda v: np.concatenate([[0], np.cumsum(v)]).astype(np.int32)
    a = dict(sub_frag_ptr=sub_frag_ptr, sub_cluster_ptr=sub_cluster_ptr, sub_sv_ptr=np.zeros(len(sub_frag_ptr)),
             frag_aln_ptr=frag_aln_ptr, aln_numerrors=numerr, aln_len=alen, aln_exp=aexp, aln_esp_ptr=aln_esp_ptr,
             aln_esp_cluster=esp_cluster, aln_esp_leaf=esp_leaf, cluster_kind=np.zeros(C), supports_order=sup_order,
             cluster_hist_ptr=cum(nl + 1), cluster_leaf_ptr=cum(nl), leaf_owner=[f for l in leaves for f in l],
             cluster_root_ptr=np.arange(C + 1), root_leaf_ptr=cum(nl), root_leaf=[i for l in leaves for i in range(len(l))],
             sv_cluster=np.zeros(0), sv_frag_ptr=np.zeros(1), sv_frag=np.zeros(0), sv_numchanged=np.zeros(0),
             exp_pseq=pseq, exp_lambda=lam,
             logprobcov=np.concatenate([synth.logprobcov_table(float(l), max_support) for l in lam]))
    a = {k: np.asarray(v, dtype=base.arrays[k].dtype) for k, v in a.items()}
    sc = dict(base.scalars)
    sc.update(n_sub=len(sub_frag_ptr) - 1, n_frag=sub_frag_ptr[-1], n_aln=len(numerr), n_aln_esp=len(esp_cluster), n_cluster=C,
              n_exp=E, max_support=max_support, n_sv=0, n_sv_frag=0, n_leaf=int(nl.sum()), n_root=C, n_root_leaf=int(nl.sum()),
              n_hist=int((nl + 1).sum()))
    return sc, a


@pytest.mark.gpu
@pytest.mark.parametrize("n_chains", [3, 64])
def test_alignments_with_more_slots_than_the_slot_record(oracle, mbsv, n_chains):
    sc, a = _many_slot_problem(mbsv, seed=11)
    nesp = np.diff(a["aln_esp_ptr"])
    assert nesp.max() >= 6 and (nesp <= 4).any()
    p, op = mbsv.PackedProblem(sc, a), oracle.Problem(sc, a)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", 3000, n_chains=n_chains, nthreads=8, java_int_wrap=False, memo_states=0)
    g = dp.run(mbsv.MODE_FAST, 3000, n_chains=n_chains, memo_states=0)
    assert_same(o, g)
    assert {abs(li["rows"]) for li in dp.launch_info()} >= ({48, 96, 0} if n_chains == 64 else {0})
    o = oracle.run(op, "faithful", 1500, n_chains=n_chains, nthreads=8, trace=True, memo_states=0)
    g = dp.run(mbsv.MODE_REPLAY_EXACT, 1500, n_chains=n_chains, trace=True, memo_states=0)
    assert_same(o, g, trace=True)


@pytest.mark.gpu
def test_device_memory_pool_across_problems(oracle, mbsv):
    """Blocks of a destroyed problem serve the next one (mbsv_core.cu DevPool): results do not depend on whether a buffer is
    fresh or reused (every run initialises what it reads), and mbsv_trim gives the kept blocks back."""
    import torch
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=60000, seed=21)          # large enough for blocks above the pool's 1 MB threshold
    q = synth.generate("cfg4", n_frag=45000, seed=22)
    def run(pp):
        dp = mbsv.DeviceProblem(pp)
        g = dp.run(mbsv.MODE_FAST, 600, n_chains=32)
        dp.close()
        return g
    mbsv.trim()
    first = run(p)
    free_kept = torch.cuda.mem_get_info()[0]
    other = run(q)                                            # a different problem in between dirties the kept blocks
    again = run(p)
    for k in ("aln_count", "frag_err_count", "cluster_hist", "final_assign", "counters", "rng_state", "final_logprob"):
        assert np.array_equal(getattr(first, k), getattr(again, k)), k
    mbsv.trim()
    assert torch.cuda.mem_get_info()[0] > free_kept           # the blocks went back to the driver
    o = oracle.run(oracle.Problem(q.scalars, q.arrays), "incremental", 600, n_chains=32, nthreads=8, java_int_wrap=False)
    assert_same(o, other)


@pytest.mark.gpu
@pytest.mark.parametrize("n_chains", [32, 64])
def test_ratio_tables_of_small_multi_fragment_subproblems(oracle, mbsv, n_chains):
    """Full warps of one subproblem (chains a multiple of 32): subproblems of several fragments without AlterSV candidates
    and at most 32 joint states take their acceptance ratios from a table built in shared memory (memo_kernel.cuh `rt`),
    one-fragment ones likewise (`f1`); both against the oracle, in both modes."""
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=4000, seed=17, shuffle_orders=True)
    a = p.arrays
    F, nsv = np.diff(a["sub_frag_ptr"]), np.diff(a["sub_sv_ptr"])
    nst = np.exp(np.add.reduceat(np.log(np.diff(a["frag_aln_ptr"]) + 1.0), a["sub_frag_ptr"][:-1]))
    assert ((F >= 2) & (nsv == 0) & (nst <= 32.5)).sum() >= 20 and ((F == 1) & (nsv == 0)).sum() >= 100
    op = oracle.Problem(p.scalars, p.arrays)
    dp = mbsv.DeviceProblem(p)
    o = oracle.run(op, "incremental", 2500, n_chains=n_chains, nthreads=8, java_int_wrap=False)
    g = dp.run(mbsv.MODE_FAST, 2500, n_chains=n_chains)
    assert_same(o, g)
    assert any(li["rows"] < 0 for li in dp.launch_info())          # the memoised kernel ran
    o = oracle.run(op, "faithful", 800, n_chains=n_chains, nthreads=8, trace=True)
    g = dp.run(mbsv.MODE_REPLAY_EXACT, 800, n_chains=n_chains, trace=True)
    assert_same(o, g, trace=True)
