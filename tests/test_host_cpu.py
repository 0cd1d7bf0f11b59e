"""The native host mirror (include/mbsv_host.h, multibreak_sv_b200.host.MultiBreakSV) against the oracle's literal
restatement of the reference parser: identical packed arrays (integer-exact, including the HashMap orders), identical
rejection of malformed input, and output files in the reference's format (Double.toString)."""
import os

import numpy as np
import pytest

from conftest import EXAMPLE

ARGS = ["--numiterations", "10000", "--perr", "0.001"]


def _host(mbsv, c, a, e, cf=False, extra=()):
    from multibreak_sv_b200 import host
    args = ["--clusterfile", c, "--assignmentfile", a, "--experimentfile", e] + ARGS + list(extra)
    if cf:
        args.append("--readclustersfirst")
    return host.MultiBreakSV(args)


def _same(p, q):
    assert p.scalars == q.scalars
    for k in p.arrays:
        assert np.array_equal(p.arrays[k], q.arrays[k]), k


def test_example_matches_oracle(mbsv, oracle, example_problem):
    sv = _host(mbsv, os.path.join(EXAMPLE, "gasv.clusters"), os.path.join(EXAMPLE, "assignments.txt"),
               os.path.join(EXAMPLE, "experiments.txt"))
    sv.getFragmentAlignments(); sv.constructClusterDiagrams(); sv.precomputePriors(); sv.fillNonSingletonArray()
    p = sv.packed()
    _same(p, example_problem)
    assert sv.names(0, 4) == example_problem.names["frag_id"]
    assert sv.names(1, 27) == example_problem.names["cluster_id"]
    assert sv.names(3, 86) == example_problem.names["aln_esp_name"]
    assert sv.sortedFragments(4).tolist() == example_problem.names["sorted_frag"].tolist()
    assert sv.burnin == 1000 and sv.rejected() == ""


@pytest.mark.parametrize("cf", [False, True])
@pytest.mark.parametrize("cfg,nfrag,cc", [("cfg4", 600, 0.4), ("cfg2", 1500, 0.0), ("cfg3", 300, 0.3)])
def test_generated_files_match_oracle(mbsv, oracle, tmp_path, cf, cfg, nfrag, cc):
    from multibreak_sv_b200 import synth
    g = synth.generate(cfg, n_frag=nfrag, seed=31337, conncomp_frac=cc)
    paths = synth.write_reference_files(g, str(tmp_path))
    sv = _host(mbsv, paths["clusters"], paths["assignments"], paths["experiments"], cf).load()
    q = oracle.load_files(paths["clusters"], paths["assignments"], paths["experiments"], clusters_first=cf)
    p = sv.packed()
    _same(p, q)
    assert sv.names(0, p.scalars["n_frag"]) == q.names["frag_id"]
    assert sv.names(1, p.scalars["n_cluster"]) == q.names["cluster_id"]
    assert sv.names(2, p.scalars["n_exp"]) == q.names["exp_label"]
    assert sv.warning() == q.names["warning"]


def test_subset_of_clusters_with_readclustersfirst(mbsv, oracle, tmp_path):
    """A subproblem file (a few rows of the clusters file) + --readclustersfirst: fragments outside it vanish."""
    from multibreak_sv_b200 import host, synth
    g = synth.generate("cfg4", n_frag=500, seed=5)
    paths = synth.write_reference_files(g, str(tmp_path))
    rows = open(paths["clusters"]).read().splitlines()
    hosts = []
    for k, chunk in enumerate((rows[:40], rows[40:45], rows[45:300])):
        sub = os.path.join(str(tmp_path), f"subset-{k + 1}.clusters")
        open(sub, "w").write("\n".join(chunk) + "\n")
        sv = _host(mbsv, sub, paths["assignments"], paths["experiments"], cf=True).load()
        _same(sv.packed(), oracle.load_files(sub, paths["assignments"], paths["experiments"], clusters_first=True))
        hosts.append(sv)
    b = host.batch(hosts)
    assert b.scalars["n_sub"] == 3
    assert b.scalars["n_frag"] == sum(h.packed().scalars["n_frag"] for h in hosts)
    # the batch is the concatenation with re-based indices: running it on the oracle equals running the parts
    ob = oracle.Problem(b.scalars, b.arrays)
    rb = oracle.run(ob, "faithful", 500, n_chains=2)
    off_a = 0
    for s, h in enumerate(hosts):
        ps = h.packed()
        rs = oracle.run(oracle.Problem(ps.scalars, ps.arrays), "faithful", 500, n_chains=2)
        n = ps.scalars["n_aln"]
        assert np.array_equal(rb.aln_count[off_a:off_a + n], rs.aln_count)
        assert np.array_equal(rb.counters[s], rs.counters[0])
        off_a += n


def test_malformed_inputs_rejected_like_the_oracle(mbsv, oracle, tmp_path):
    from multibreak_sv_b200 import host

    def w(name, text):
        p = os.path.join(str(tmp_path), name)
        open(p, "w").write(text)
        return p
    exp = w("e.txt", "ExperimentLabel\tPseq\tLambdaD\nx 0.1 3\n")
    asg = ("FragmentID\tDiscordantPairs\tNumErrors\tBasesInAlignment\tExperimentLabel\n"
           "longread_1\tlongread_1_0.0-1.0\t5\t100\tx\nlongread_1\tlongread_1_0.1-1.1\t7\t100\tx\n"
           "longread_2\tlongread_2_0.0-1.0\t5\t100\tx\n")
    good = "c1\t2\t10\tD\tlongread_1_0.0-1.0, longread_2_0.0-1.0\t1\t1\t1, 2\nc2\t1\t10\tD\tlongread_1_0.1-1.1\t1\t1\t1, 2\n"
    a, c = w("a.txt", asg), w("c.txt", good)
    _same(_host(mbsv, c, a, exp).load().packed(), oracle.load_files(c, a, exp))
    with pytest.raises(mbsv.MbsvError, match="not in Experiments file") as e:
        _host(mbsv, c, w("a2.txt", asg.replace("\tx\n", "\ty\n", 1)), exp).load()
    assert e.value.status == -7
    with pytest.raises(mbsv.MbsvError, match="ArrayIndexOutOfBounds"):
        _host(mbsv, w("c2.txt", good.replace("\nc2", "\n\nc2")), a, exp).load()
    dup = _host(mbsv, w("c3.txt", good + "c3\t1\t10\tD\tlongread_2_0.0-1.0\t1\t1\t1, 2\n"), a, exp).load()
    assert "more than one cluster" in dup.rejected()
    with pytest.raises(mbsv.MbsvError) as e:
        dup.packed()
    assert e.value.status == -6
    with pytest.raises(mbsv.MbsvError, match="cannot open"):
        _host(mbsv, "/nonexistent", a, exp).load()
    for bad in (["--numiterations", "5"], ["--clusterfile", c, "--assignmentfile", a, "--experimentfile", exp, "--perr", "2",
                                            "--numiterations", "5"], ["--bogus"]):
        with pytest.raises(ValueError):
            host.MultiBreakSV(bad)


def test_hashmap_order_matches_oracle_including_collisions(mbsv, oracle):
    from multibreak_sv_b200 import host
    rng = np.random.default_rng(3)
    for n in (1, 12, 13, 24, 25, 49, 97, 1000, 20000):
        ks = [f"longread_{v}" for v in rng.choice(10 ** 7, n, replace=False)]
        o_order, o_cap, o_tree = oracle.hashmap_order(ks)
        h_order, h_cap = host.java_hashmap_order(ks)
        assert not o_tree and h_order == o_order and h_cap == o_cap
    # "Aa" and "BB" share a hashCode: 2^k colliding keys exercise treeifyBin()'s resize below capacity 64
    coll = ["".join(p) for p in __import__("itertools").product(("Aa", "BB"), repeat=4)]      # 16 equal hashes
    for n in (8, 9, 10):
        o_order, o_cap, o_tree = oracle.hashmap_order(coll[:n])
        assert not o_tree
        h_order, h_cap = host.java_hashmap_order(coll[:n])
        assert h_order == o_order and h_cap == o_cap
    assert oracle.hashmap_order(coll[:10])[1] == 64                      # 9th and 10th entries doubled 16 -> 32 -> 64
    assert oracle.hashmap_order(coll[:11])[2] is True                    # next one lands in a 10-entry bin at capacity 64
    # ... which the host mirror no longer refuses: it replays the map through the literal red-black emulation
    # (csrc/java_hashmap.hpp; tests/test_hashmap_tree_bins.py); the order is NOT the plain-list order any more
    import java_hashmap_ref
    m = java_hashmap_ref.JavaHashMap()
    for k in coll[:11]:
        m.put(k)
    h_order, h_cap = host.java_hashmap_order(coll[:11])
    assert h_order == m.keys() and h_cap == 64 and h_order != coll[:11]


def test_double_to_string(mbsv):
    from multibreak_sv_b200 import host
    want = {834.0: "834.0", 0.391: "0.391", 0.999: "0.999", 1e-4: "1.0E-4", 1e7: "1.0E7", 9999999.0: "9999999.0",
            0.001: "0.001", 123456789.125: "1.23456789125E8", -35.263569332642774: "-35.263569332642774",
            0.0: "0.0", 1 / 3: "0.3333333333333333", 100.0: "100.0", 1.5e-5: "1.5E-5", -2.5e10: "-2.5E10",
            float("inf"): "Infinity", 608 / 1000: "0.608"}
    for v, s in want.items():
        assert host.java_double_to_string(v) == s, (v, s)
    assert host.java_double_to_string(float("nan")) == "NaN"


def test_write_final_files_format(mbsv, oracle, example_problem, tmp_path):
    """writeFinalFiles (MultiBreakSV.java:1754-1792) from the oracle's tallies of the 10,000-iteration example run."""
    sv = _host(mbsv, os.path.join(EXAMPLE, "gasv.clusters"), os.path.join(EXAMPLE, "assignments.txt"),
               os.path.join(EXAMPLE, "experiments.txt"), extra=["--prefix", os.path.join(str(tmp_path), "out")]).load()
    r = oracle.run(example_problem, "faithful", 10000)
    sv.setResults(r.aln_count, r.frag_err_count, r.cluster_hist)
    sv.writeFinalFiles()
    bp = open(os.path.join(str(tmp_path), "out.finalbreakpoints.longburnin")).read().splitlines()
    assert bp[0] == "ClusterID\tNumIters\tSupport0\tSupport1\tSupport2..."
    assert [l.split("\t")[0] for l in bp[1:]] == example_problem.names["cluster_id"]          # breakpoints.keySet() order
    row = dict((l.split("\t")[0], l.split("\t")) for l in bp[1:])["c4125"]
    assert row[1] == "1000" and len(row) == 2 + 27
    assert row[2:6] == ["1.0", "0.0", "165.0", "834.0"]       # bin 0 = burnin - sum(bins >= 1): 999 samples / 1000
    asg = open(os.path.join(str(tmp_path), "out.finalassignment.longburnin")).read().splitlines()
    assert asg[0] == "FragmentID\tAssignmentIndex\tAvgAssignment\tDiscordantPairs"
    assert asg[1] == "longread_118669\t-1\t0.0\terror"
    assert asg[2] == "longread_118669\t0\t0.325\tlongread_118669_0.0.0-0.1.0"
    body = [l.split("\t") for l in asg[1:]]
    assert [b[0] for b in body] == sorted(b[0] for b in body) and len(body) == 69 + 4
    l74 = [b for b in body if b[0] == "longread_118674"]
    assert [b[2] for b in l74] == ["0.0", "0.608", "0.391"]
    l73 = [b for b in body if b[0] == "longread_118673" and b[1] == "-1"][0]
    assert l73[2] == "0.999"
    multi = [b for b in body if "," in b[3]]
    assert len(multi) == 17 and all(", " not in b[3] for b in multi)


PERL_SPLITTER = "/root/reference/src/generateIndependentSubproblems.pl"


@pytest.mark.skipif(not os.path.exists(PERL_SPLITTER), reason="the reference's Perl script is only present in the build container")
@pytest.mark.parametrize("cfg,nfrag,cc", [("cfg4", 700, 0.3), ("cfg2", 2000, 0.0)])
def test_partition_matches_the_reference_perl_script(mbsv, tmp_path, cfg, nfrag, cc):
    """Live oracle for the partition: the unmodified generateIndependentSubproblems.pl (line order inside a subset
    file is Perl-hash-random, so files are compared as sets of lines; the numbering must match exactly)."""
    import subprocess
    from multibreak_sv_b200 import host, synth
    g = synth.generate(cfg, n_frag=nfrag, seed=2718, conncomp_frac=cc)
    # long-read numbers in shuffled order so that "smallest unseen id" differs from file order
    nums = np.random.default_rng(1).permutation(nfrag) * 3 + 1000
    paths = synth.write_reference_files(g, str(tmp_path), frag_numbers=nums)
    pdir, ndir = os.path.join(str(tmp_path), "perl"), os.path.join(str(tmp_path), "native")
    os.makedirs(pdir); os.makedirs(ndir)
    subprocess.run(["perl", PERL_SPLITTER, paths["clusters"], pdir], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    n, rows = host.generateIndependentSubproblems(paths["clusters"], ndir)
    pfiles = sorted(os.listdir(pdir))
    assert len(pfiles) == n >= g.scalars["n_sub"] and sorted(os.listdir(ndir)) == pfiles
    for f in pfiles:
        assert sorted(open(os.path.join(pdir, f)).read().splitlines()) == sorted(open(os.path.join(ndir, f)).read().splitlines()), f
    lines = open(paths["clusters"]).read().splitlines()
    assert all((r == 0) == ("." in l.split("\t")[0]) for r, l in zip(rows, lines))


def test_partition_then_batch_refines_generator_subproblems(mbsv, oracle, tmp_path):
    """partition -> one host per subset file (--readclustersfirst) -> batch: every fragment lands in exactly one
    subproblem, and each connected component lies inside one generated subproblem (the generator does not force
    its subproblems to be connected, so the partition may be finer)."""
    from multibreak_sv_b200 import host, synth
    g = synth.generate("cfg4", n_frag=300, seed=99)
    paths = synth.write_reference_files(g, str(tmp_path))
    sdir = os.path.join(str(tmp_path), "subs")
    os.makedirs(sdir)
    n, _ = host.generateIndependentSubproblems(paths["clusters"], sdir)
    assert n >= g.scalars["n_sub"]
    hosts = [_host(mbsv, os.path.join(sdir, f"subset-{k}.clusters"), paths["assignments"], paths["experiments"], cf=True).load()
             for k in range(1, n + 1)]
    b = host.batch(hosts)
    assert b.scalars["n_sub"] == n and b.scalars["n_frag"] == 300 and b.scalars["n_aln"] == g.scalars["n_aln"]
    fp = g.arrays["sub_frag_ptr"]
    gen_sub = {f"longread_{(f + 1) * 7 + 100000}": s for s in range(g.scalars["n_sub"]) for f in range(fp[s], fp[s + 1])}
    seen = set()
    for h in hosts:
        ids = h.names(0, h.packed().scalars["n_frag"])
        assert len({gen_sub[i] for i in ids}) == 1 and not (seen & set(ids))
        seen |= set(ids)
    assert seen == set(gen_sub)


@pytest.mark.parametrize("cfg,nfrag,cc,seed", [("cfg4", 400, 0.3, 99), ("cfg2", 900, 0.0, 5), ("cfg3", 250, 0.4, 17)])
def test_load_partitioned_equals_one_host_per_subset_file(mbsv, tmp_path, cfg, nfrag, cc, seed):
    """mbsv_host_load_partitioned (partition + load in memory, assignments file read once) must give exactly the batch
    of: generateIndependentSubproblems -> subset-k.clusters files -> one --readclustersfirst host per file -> batch."""
    from multibreak_sv_b200 import host, synth
    g = synth.generate(cfg, n_frag=nfrag, seed=seed, conncomp_frac=cc)
    nums = np.random.default_rng(seed).permutation(nfrag) * 3 + 1000          # subset numbering != file order
    paths = synth.write_reference_files(g, str(tmp_path), frag_numbers=nums)
    sdir = os.path.join(str(tmp_path), "subs")
    os.makedirs(sdir)
    n, _ = host.generateIndependentSubproblems(paths["clusters"], sdir)
    hosts = [_host(mbsv, os.path.join(sdir, f"subset-{k}.clusters"), paths["assignments"], paths["experiments"], cf=True).load()
             for k in range(1, n + 1)]
    want = host.batch([h for h in hosts if h.packed().scalars["n_frag"] > 0])
    for threads in ("1", "3", "7"):                  # worker threads load ranges of subsets, merged in subset order
        os.environ["MBSV_HOST_THREADS"] = threads
        try:
            got, skipped = host.load_partitioned(paths["clusters"], paths["assignments"], paths["experiments"])
        finally:
            del os.environ["MBSV_HOST_THREADS"]
        assert skipped == sum(1 for h in hosts if h.packed().scalars["n_frag"] == 0)
        _same(got, want)
        assert got.scalars["n_sub"] == n - skipped and got.scalars["n_frag"] == nfrag


def test_batch_final_files_are_the_concatenated_per_subset_outputs(mbsv, oracle, tmp_path):
    """PartitionedBatch.writeFinalFiles == header + the rows of every `subset-k` run's own output files, in subset order."""
    from multibreak_sv_b200 import host, synth
    g = synth.generate("cfg4", n_frag=300, seed=41, conncomp_frac=0.25)
    paths = synth.write_reference_files(g, str(tmp_path), frag_numbers=np.random.default_rng(2).permutation(300) * 3 + 500)
    b = host.PartitionedBatch(paths["clusters"], paths["assignments"], paths["experiments"])
    p = b.packed
    r = oracle.run(oracle.Problem(p.scalars, p.arrays), "incremental", 2000, n_chains=4, nthreads=4, java_int_wrap=False)
    burnin = 4 * 200                                            # pooled over the 4 chains: window length x chains
    prefix = os.path.join(str(tmp_path), "batch")
    b.writeFinalFiles(prefix, burnin, r.aln_count, r.frag_err_count, r.cluster_hist)
    sdir = os.path.join(str(tmp_path), "subs")
    os.makedirs(sdir)
    n, _ = host.generateIndependentSubproblems(paths["clusters"], sdir)
    assert b.numbers.tolist() == list(range(1, n + 1)) and b.skipped == 0
    a = p.arrays
    bp, asg = [], []
    for i, k in enumerate(b.numbers):
        sv = _host(mbsv, os.path.join(sdir, f"subset-{k}.clusters"), paths["assignments"], paths["experiments"], cf=True,
                   extra=["--prefix", os.path.join(sdir, f"out-{k}")]).load()
        f0, f1 = a["sub_frag_ptr"][i], a["sub_frag_ptr"][i + 1]
        c0, c1 = a["sub_cluster_ptr"][i], a["sub_cluster_ptr"][i + 1]
        sv.setResults(r.aln_count[a["frag_aln_ptr"][f0]:a["frag_aln_ptr"][f1]], r.frag_err_count[f0:f1],
                      r.cluster_hist[a["cluster_hist_ptr"][c0]:a["cluster_hist_ptr"][c1]], numiters=8000)     # burnin 800
        assert sv.burnin == burnin
        sv.writeFinalFiles()
        bp += open(os.path.join(sdir, f"out-{k}.finalbreakpoints.longburnin")).read().splitlines()[1:]
        asg += open(os.path.join(sdir, f"out-{k}.finalassignment.longburnin")).read().splitlines()[1:]
    got_bp = open(prefix + ".finalbreakpoints.longburnin").read().splitlines()
    got_asg = open(prefix + ".finalassignment.longburnin").read().splitlines()
    assert got_bp[0] == "ClusterID\tNumIters\tSupport0\tSupport1\tSupport2..." and got_bp[1:] == bp
    assert got_asg[0] == "FragmentID\tAssignmentIndex\tAvgAssignment\tDiscordantPairs" and got_asg[1:] == asg
    b.close()


def test_load_partitioned_on_the_shipped_second_data_set(mbsv, oracle):
    """Ten independent clusters: each subproblem equals the oracle's load of the shipped iterK.clusters file."""
    from conftest import GOLDEN
    from multibreak_sv_b200 import host
    pre = os.path.join(GOLDEN, "preproc_infiles")
    files = [os.path.join(pre, "gasv.in.clusters"), os.path.join(pre, "assignments.txt"), os.path.join(pre, "experiments.txt")]
    got, skipped = host.load_partitioned(*files)
    assert (got.scalars["n_sub"], skipped, got.scalars["n_frag"], got.scalars["n_cluster"]) == (10, 0, 16, 10)
    # subset 1 is c7 (longread_1, two alignments): iter6.clusters in the shipped numbering
    q = oracle.load_files(os.path.join(pre, "cluster-subproblems", "iter6.clusters"), files[1], files[2], clusters_first=True)
    a = got.arrays
    assert a["sub_frag_ptr"][1] == 1 and a["frag_aln_ptr"][1] == 2 and q.scalars["n_frag"] == 1 and q.scalars["n_aln"] == 2
    assert a["aln_numerrors"][:2].tolist() == q.arrays["aln_numerrors"].tolist() == [1007, 1002]


def test_matchfile_initial_state(mbsv, tmp_path):
    """--matchfile (MultiBreakSV.java:1030-1085): the shipped initialset.txt never matches (its IDs lack `longread_`), so
    every fragment starts unassigned; a matching single-ESP alignment is set, a two-ESP alignment never is (thisval <= 1)."""
    from multibreak_sv_b200 import host
    sv = _host(mbsv, os.path.join(EXAMPLE, "gasv.clusters"), os.path.join(EXAMPLE, "assignments.txt"),
               os.path.join(EXAMPLE, "experiments.txt"), extra=["--matchfile", os.path.join(EXAMPLE, "initialset.txt")]).load()
    p = sv.packed()
    assert sv.initialAssignmentFromMatchfile(p).tolist() == [-1, -1, -1, -1]
    names = sv.names(3, p.scalars["n_aln_esp"])
    a = p.arrays
    # pick a single-ESP alignment of fragment 2 and a two-ESP alignment of fragment 0
    single = next(al for al in range(a["frag_aln_ptr"][2], a["frag_aln_ptr"][3]) if a["aln_esp_ptr"][al + 1] - a["aln_esp_ptr"][al] == 1)
    double = next(al for al in range(a["frag_aln_ptr"][0], a["frag_aln_ptr"][1]) if a["aln_esp_ptr"][al + 1] - a["aln_esp_ptr"][al] == 2)
    mf = os.path.join(str(tmp_path), "match.txt")
    open(mf, "w").write(names[a["aln_esp_ptr"][single]] + "\n" + names[a["aln_esp_ptr"][double]] + "\n" + names[a["aln_esp_ptr"][double] + 1] + "\n")
    sv.matchfile = mf
    init = sv.initialAssignmentFromMatchfile(p)
    assert init[2] == single - a["frag_aln_ptr"][2]
    # the first alignment of fragment 0 that contains a listed ESP and has exactly one ESP wins, two-ESP ones never do
    al0 = a["frag_aln_ptr"][0] + init[0] if init[0] >= 0 else None
    assert al0 is None or a["aln_esp_ptr"][al0 + 1] - a["aln_esp_ptr"][al0] == 1
    c = sv.matchfile_counts
    assert c["esps"] == 3 and c["numset"] == int((init >= 0).sum()) and c["numset"] + c["numerr"] + c["numbad"] == 4
    # the same rule, restated: first alignment whose ESP count equals min(1, listed ESPs it contains)
    listed = set(open(mf).read().split())
    for f in range(4):
        want = -1
        for al in range(a["frag_aln_ptr"][f], a["frag_aln_ptr"][f + 1]):
            j0, j1 = a["aln_esp_ptr"][al], a["aln_esp_ptr"][al + 1]
            if int(any(names[j] in listed for j in range(j0, j1))) == j1 - j0:
                want = al - a["frag_aln_ptr"][f]
                break
        assert init[f] == want


def test_cli_usage_and_loud_failure_without_gpu(mbsv):
    """mbsv_cli mirrors the reference's option checks (MultiBreakSV.java:211-267) and never falls back to a CPU path."""
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multibreak_sv_b200", "mbsv_cli")
    assert os.path.exists(cli), "run make -C multibreak_sv_b200/csrc"
    files = ["--clusterfile", os.path.join(EXAMPLE, "gasv.clusters"), "--assignmentfile", os.path.join(EXAMPLE, "assignments.txt"),
             "--experimentfile", os.path.join(EXAMPLE, "experiments.txt")]
    for args, msg in ((["--clusterfile", "x"], "Assignment file is required."),
                      (files, "Number of iterations is required to sample."),
                      (files + ["--numiterations", "10"], "Mapping error probability (perr) is required."),
                      (files + ["--numiterations", "10", "--perr", "1.5"], "--perr must be a value between 0 and 1."),
                      (files + ["--numiterations", "ten", "--perr", "0.1"], "--numiterations must be a positive integer."),
                      (files + ["--bogus"], "--bogus is not a valid option."),
                      (files + ["--enumerate"], "--enumerate is outside the sampler path")):
        r = subprocess.run([cli] + args, capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([cli] + files + ["--numiterations", "10", "--perr", "0.001", "--prefix", "/tmp/never"], capture_output=True, text=True)
        assert r.returncode == 2 and "runMCMC:" in r.stderr and not os.path.exists("/tmp/never.finalassignment.longburnin")


def test_load_partitioned_reports_the_failing_subset(mbsv, tmp_path):
    """A rejected subset (an ESP listed twice in one cluster row) fails the whole load and names the subset; an unknown
    experiment label and a missing file fail as in the single-problem path."""
    from multibreak_sv_b200 import host
    from multibreak_sv_b200._abi import MbsvError
    d = str(tmp_path)
    open(os.path.join(d, "e.txt"), "w").write("ExperimentLabel\tPseq\tLambdaD\npacbio\t0.15\t3\n")
    open(os.path.join(d, "a.txt"), "w").write("FragmentID\tDiscordantPairs\tNumErrors\tLen\tExperimentLabel\n"
                                              "longread_1\tlongread_1_a\t5\t100\tpacbio\nlongread_2\tlongread_2_a\t7\t120\tpacbio\n"
                                              "longread_2\tlongread_2_b\t9\t130\tpacbio\n")
    good = "c1\t1\t10.0\tD\tlongread_1_a\t1\t1\t1, 2, 3, 4\nc2\t2\t10.0\tD\tlongread_2_a, longread_2_b\t1\t1\t1, 2, 3, 4\n"
    open(os.path.join(d, "c.txt"), "w").write(good)
    p, skipped = host.load_partitioned(os.path.join(d, "c.txt"), os.path.join(d, "a.txt"), os.path.join(d, "e.txt"))
    assert (p.scalars["n_sub"], p.scalars["n_frag"], p.scalars["n_aln"], skipped) == (2, 2, 3, 0)
    open(os.path.join(d, "bad.txt"), "w").write(good.replace("longread_2_a, longread_2_b", "longread_2_a, longread_2_a"))
    with pytest.raises(MbsvError) as e:
        host.load_partitioned(os.path.join(d, "bad.txt"), os.path.join(d, "a.txt"), os.path.join(d, "e.txt"))
    assert e.value.status == -6 and "subset 2" in str(e.value) and "listed twice" in str(e.value)
    open(os.path.join(d, "a2.txt"), "w").write(open(os.path.join(d, "a.txt")).read().replace("130\tpacbio", "130\tillumina"))
    with pytest.raises(MbsvError) as e:
        host.load_partitioned(os.path.join(d, "c.txt"), os.path.join(d, "a2.txt"), os.path.join(d, "e.txt"))
    assert e.value.status == -7 and "illumina" in str(e.value)
    with pytest.raises(MbsvError) as e:
        host.load_partitioned(os.path.join(d, "c.txt"), os.path.join(d, "missing.txt"), os.path.join(d, "e.txt"))
    assert e.value.status == -7
