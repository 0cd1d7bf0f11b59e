"""Pins the CPU oracle: public java.util.Random constants, the derived known-answers of SURVEY.md §8c
(KAT-1..KAT-6: an independent line-by-line emulation of the reference), glibc log/exp within 1 ulp, exact
enumeration of tiny posteriors, and faithful == incremental on every integer output."""
import hashlib
import math
import os

import numpy as np
import pytest

from conftest import EXAMPLE

KAT2_FRAGS = ["longread_118671", "longread_118673", "longread_118674", "longread_118669"]
KAT2_CLUSTERS = ("c4136 c4135 c4134 c4133 c4155 c4139 c4138 c4137 c4140 c4125 c4147 c4146 c4145 c4129 c4128 c4127 "
                 "c4149 c4126 c4148 c4150 c4132 c4154 c4131 c4153 c4130 c4152 c4151").split()
KAT6_SHA = "27aa9c30b1a4ac191a2942d3a8930d0fc723a978ee800960c2a7ed375e6bfcc6"


def test_kat1_java_random(oracle):
    assert oracle.random_nextint(42) == -1170105035          # public constant
    assert oracle.random_nextint(0) == -1155484576           # public constant
    assert oracle.random_doubles(0, 1)[0] == 0.730967787376657
    assert oracle.random_doubles(123456, 6).tolist() == [0.4132192266296165, 0.990414373129817, 0.2522554407939305,
                                                         0.7056636259919211, 0.07343842024000635, 0.5360645079196139]
    # nextInt(bound): power-of-two shortcut and rejection loop stay in range, deterministic
    for bound in (1, 2, 3, 7, 16, 33, 1 << 20, (1 << 30) + 1):
        v = oracle.random_bounded(5, bound, 1000)
        assert v.min() >= 0 and v.max() < bound
    # new Random(0).nextInt(10) sequence (public: 0, 8, 9, 7, 5 ... from java docs examples "60 48 29 47 15 53 91 61" for 100)
    assert oracle.random_bounded(0, 100, 8).tolist() == [60, 48, 29, 47, 15, 53, 91, 61]


def _jhash(s):
    h = 0
    for c in s:
        h = (31 * h + ord(c)) & 0xFFFFFFFF
    return h


def _closed_form_order(keys, cap):
    """Without removals/tree bins, HashMap iteration order == sort by (bucket at the final capacity, insertion)."""
    return [k for _, _, k in sorted(((_jhash(k) ^ (_jhash(k) >> 16)) & (cap - 1), i, k) for i, k in enumerate(keys))]


def test_string_hash_and_hashmap_order(oracle):
    assert oracle.string_hash("") == 0
    assert oracle.string_hash("a") == 97
    assert oracle.string_hash("hello") == 99162322           # public constant
    h = sum(ord(c) * 31 ** (14 - i) for i, c in enumerate("longread_118671")) & 0xFFFFFFFF
    assert oracle.string_hash("longread_118671") == (h - (1 << 32) if h >= (1 << 31) else h)
    order, cap, tree = oracle.hashmap_order(["longread_118669", "longread_118671", "longread_118673", "longread_118674"])
    assert order == KAT2_FRAGS and cap == 16 and not tree
    # Integer-like keys "0".."12": hash = char code(s); order by bucket
    keys = [str(i) for i in range(13)]
    order, cap, _ = oracle.hashmap_order(keys)
    assert cap == 32 and order == ["11", "12", "0", "1", "2", "3", "4", "5", "6", "7", "8", "9", "10"]
    assert order == _closed_form_order(keys, 32)
    rng = np.random.default_rng(11)
    for n in (5, 12, 13, 24, 25, 100, 1000, 5000):
        ks = [f"longread_{v}" for v in rng.choice(10 ** 6, n, replace=False)]
        order, cap, tree = oracle.hashmap_order(ks)
        want_cap = 16
        while n > 0.75 * want_cap:
            want_cap *= 2
        assert not tree and cap == want_cap and order == _closed_form_order(ks, cap)
    # removal does not shrink the table and keeps the relative order of the others
    order1, _, _ = oracle.hashmap_order(keys)
    order2, cap2, _ = oracle.hashmap_order(keys, removed=["3", "11"])
    assert cap2 == 32 and order2 == [k for k in order1 if k not in ("3", "11")]


def test_kat2_kat3_kat4_example_structure(oracle, example_problem):
    p = example_problem
    assert p.scalars["n_frag"] == 4 and p.scalars["n_aln"] == 69 and p.scalars["n_aln_esp"] == 86
    assert p.scalars["n_cluster"] == 27 and p.scalars["n_exp"] == 1 and p.scalars["max_support"] == 26
    assert p.names["frag_id"] == KAT2_FRAGS
    assert p.names["cluster_id"] == KAT2_CLUSTERS
    assert np.diff(p.arrays["frag_aln_ptr"]).tolist() == [32, 21, 2, 14]
    assert p.scalars["n_sv"] == 0                                    # KAT-4
    assert p.names["unsupported"] == "" and not p.names["tree_bin"]
    assert p.names["hashmap_caps"].tolist() == [16, 64, 64, 16]
    assert (p.arrays["cluster_kind"] == 0).all()
    assert len(p.names["esp_names"]) == 69
    r = oracle.run(p, "faithful", 0)
    assert r.initial_assign[0].tolist() == [31, 1, 0, 0]              # KAT-3
    assert abs(r.initial_logprob[0, 0] - (-35.263569332642774)) < 1e-12


def test_kat5_kat6_example_trajectory(oracle, example_problem):
    p = example_problem
    r = oracle.run(p, "faithful", 10000, trace=True)
    assert r.rc == 0
    mf, ma, lp = r.trace_move_frag[0, 0], r.trace_move_assign[0, 0], r.trace_logprob[0, 0]
    assert (mf[0], mf[1]) == (-1, -1)
    assert (mf[2], ma[2]) == (0, 29) and abs(lp[2] - (-34.38401938806149)) < 1e-12
    assert (mf[6], ma[6]) == (0, 14) and abs(lp[6] - (-30.894579772399503)) < 1e-12
    assert (mf[20], ma[20]) == (3, 1)
    assert (mf[21], ma[21]) == (1, 11) and abs(lp[21] - (-30.403994473303637)) < 1e-12
    assert r.counters[0, 0].tolist() == [5006, 1174, 0, 0, 5006]
    assert int(r.rng_state[0, 0]) == 0x226A00D3B97D
    assert r.final_assign[0].tolist() == [8, -1, 0, 0]
    assert abs(r.final_logprob[0, 0] - (-13.19841525274474)) < 1e-12
    fp = p.arrays["frag_aln_ptr"]
    assert r.frag_err_count.tolist() == [0, 999, 0, 0]
    assert r.aln_count[fp[2]:fp[3]].tolist() == [608, 391]
    c669 = r.aln_count[fp[3]:fp[4]]
    assert {i: int(v) for i, v in enumerate(c669) if v} == {0: 325, 1: 341, 2: 15, 7: 144, 8: 168, 12: 6}
    hp = p.arrays["cluster_hist_ptr"]
    ci = p.names["cluster_id"].index("c4125")
    hist = r.cluster_hist[hp[ci]:hp[ci + 1]]
    assert {i: int(v) for i, v in enumerate(hist) if v} == {2: 165, 3: 834}
    st = oracle.replay_states(p, r.initial_assign[0], mf, ma)
    h = hashlib.sha256()
    for row in st:
        h.update((",".join(str(row[i]) for i in p.names["sorted_frag"]) + "\n").encode())
    assert h.hexdigest() == KAT6_SHA


def test_fdlibm_within_one_ulp_of_glibc(oracle):
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(1e-300, 1e-290, 200), rng.uniform(1e-3, 1e6, 4000), rng.uniform(0.5, 2, 4000),
                         np.arange(1, 500, dtype=float), [1.0, 2.0, math.e, 1e300, 5e-324]])
    exact = 0
    for x in xs:
        a, b = oracle.jlog(float(x)), math.log(float(x))
        assert abs(a - b) <= math.ulp(b)
        exact += a == b
    assert exact > 0.95 * len(xs)
    for x in np.concatenate([rng.uniform(-745, 709, 3000), rng.uniform(-2, 2, 3000), [0.0, -0.0, 1e-30, 709.78, -745.2]]):
        a, b = oracle.jexp(float(x)), math.exp(float(x))
        assert abs(a - b) <= math.ulp(b) or (b == 0.0 and a < 1e-320)
    assert oracle.jlog(1.0) == 0.0 and oracle.jexp(0.0) == 1.0
    assert oracle.jlog(2.0) == 0.6931471805599453
    assert oracle.jexp(1000.0) == math.inf and oracle.jexp(-1000.0) == 0.0
    assert math.isnan(oracle.jlog(-1.0)) and oracle.jlog(0.0) == -math.inf


def test_faithful_equals_incremental_on_integers(oracle, mbsv):
    from multibreak_sv_b200 import synth
    for cfg, nf, chains in (("cfg2", 4000, 1), ("cfg3", 1500, 3), ("cfg4", 5000, 2)):
        p = synth.generate(cfg, n_frag=nf, shuffle_orders=True)
        op = oracle.Problem(p.scalars, p.arrays)
        a = oracle.run(op, "faithful", 1500, n_chains=chains, trace=True, nthreads=8)
        b = oracle.run(op, "incremental", 1500, n_chains=chains, trace=True, nthreads=8)
        assert a.rc == 0 and b.rc == 0
        for k in ("trace_move_frag", "trace_move_assign", "aln_count", "frag_err_count", "cluster_hist", "final_assign",
                  "counters", "rng_state"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), k
        assert np.allclose(a.trace_logprob, b.trace_logprob, rtol=0, atol=1e-8)
        assert np.array_equal(a.final_logprob, b.final_logprob)     # both recompute the final state in Java order


def test_tallies_sum_to_window(oracle, example_problem):
    for n, chains in ((10000, 1), (137, 3), (7, 2)):
        r = oracle.run(example_problem, "incremental", n, n_chains=chains)
        burnin = oracle.reference_burnin(n)
        nwin = n - max(n - burnin + 1, 0)
        fp = example_problem.arrays["frag_aln_ptr"]
        for f in range(4):
            assert int(r.frag_err_count[f]) + int(r.aln_count[fp[f]:fp[f + 1]].sum()) == nwin * chains


def _exact_marginals(oracle, prob, perr):
    lp = oracle.enumerate_logprobs(prob, 0, perr)
    w = np.exp(lp - lp.max())
    w /= w.sum()
    nal = np.diff(prob.arrays["frag_aln_ptr"])
    marg = [np.zeros(n + 1) for n in nal]
    idx = np.arange(len(lp))
    for f, n in enumerate(nal):
        d = idx % (n + 1)
        idx = idx // (n + 1)
        marg[f] = np.bincount(d, weights=w, minlength=n + 1)
    return marg


def test_exact_enumeration_stationarity(oracle, mbsv, tmp_path):
    """Long multi-chain runs of the oracle reproduce the exactly enumerated posterior of a tiny problem whose
    fragments have <= 2 alignments (where quirk Q1's proposal bias cannot bite)."""
    from multibreak_sv_b200 import synth
    rng = np.random.default_rng(5)
    tried = 0
    for seed in range(200):
        p = synth.generate("cfg4", n_frag=6, seed=1000 + seed)
        nal = np.diff(p.arrays["frag_aln_ptr"])
        if p.scalars["n_sub"] != 1 or nal.max() > 2 or nal.max() < 2:
            continue
        tried += 1
        op = oracle.Problem(p.scalars, p.arrays)
        perr = 0.05
        marg = _exact_marginals(oracle, op, perr)
        n, chains = 200000, 8
        r = oracle.run(op, "incremental", n, n_chains=chains, perr=perr, burnin=n - 2000, nthreads=8)
        nwin = (n - max(n - (n - 2000) + 1, 0)) * chains
        fp = p.arrays["frag_aln_ptr"]
        for f in range(6):
            est = np.concatenate([[r.frag_err_count[f]], r.aln_count[fp[f]:fp[f + 1]]]) / nwin
            # autocorrelated samples: generous 5-sigma band with an effective sample size of nwin/50
            tol = 5 * np.sqrt(np.maximum(marg[f] * (1 - marg[f]), 1e-4) / (nwin / 50))
            assert np.all(np.abs(est - marg[f]) < tol), (seed, f, est, marg[f])
        if tried >= 3:
            break
    assert tried >= 1
