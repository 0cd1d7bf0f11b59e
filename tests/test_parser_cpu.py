"""The oracle's restatement of the reference parser (MultiBreakSV.java:319-771) on generated text files:
the parsed problem must equal the generator's CSR up to the HashMap orders, the count-based supports must equal
the literal setBreakpoints (from ESP identities), and malformed inputs must be flagged the way the reference fails."""
import os

import numpy as np
import pytest

from conftest import EXAMPLE


def _canon(prob_arrays, frag_names, cluster_names):
    """Order-independent view: {fragment name: [(numerrors, len, exp, [cluster names of contributing ESPs])...]}"""
    a = prob_arrays
    out = {}
    for f, fn in enumerate(frag_names):
        alns = []
        for al in range(a["frag_aln_ptr"][f], a["frag_aln_ptr"][f + 1]):
            cl = [cluster_names[c] if c >= 0 else None for c in a["aln_esp_cluster"][a["aln_esp_ptr"][al]:a["aln_esp_ptr"][al + 1]]]
            alns.append((int(a["aln_numerrors"][al]), int(a["aln_len"][al]), int(a["aln_exp"][al]), cl))
        out[fn] = alns
    return out


@pytest.mark.parametrize("clusters_first", [False, True])
def test_roundtrip_generated_files(oracle, mbsv, tmp_path, clusters_first):
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=400, seed=4242, conncomp_frac=0.4)
    nums = np.arange(1, 401) * 7 + 100000
    paths = synth.write_reference_files(p, str(tmp_path), frag_numbers=nums)
    q = oracle.load_files(paths["clusters"], paths["assignments"], paths["experiments"], clusters_first=clusters_first)
    assert q.names["unsupported"] == "" and not q.names["tree_bin"]
    for k in ("n_frag", "n_aln", "n_aln_esp", "n_cluster", "n_exp", "max_support", "n_root", "n_root_leaf"):
        assert q.scalars[k] == p.scalars[k], k
    gen_frag = [f"longread_{n}" for n in nums]
    gen_cl = [f"c{c + 1}" for c in range(p.scalars["n_cluster"])]
    # experiments.keySet() order may differ from the generator's: compare through labels
    lab_gen, lab_par = p.exp_labels, q.names["exp_label"]
    ca = _canon(p.arrays, gen_frag, gen_cl)
    cb = _canon(q.arrays, q.names["frag_id"], q.names["cluster_id"])
    assert set(ca) == set(cb)
    for fn in ca:
        assert len(ca[fn]) == len(cb[fn])
        for x, y in zip(ca[fn], cb[fn]):
            assert x[:2] == y[:2] and lab_gen[x[2]] == lab_par[y[2]] and x[3] == y[3]
    # HashMap orders: fragments / clusters are permutations given by the independent closed form
    from test_oracle_kat import _closed_form_order
    assert q.names["frag_id"] == _closed_form_order(gen_frag, int(q.names["hashmap_caps"][0]))
    assert q.names["cluster_id"] == _closed_form_order(gen_cl, int(q.names["hashmap_caps"][1]))
    sup = [q.names["cluster_id"][i] for i in q.arrays["supports_order"]]
    assert sup == _closed_form_order(q.names["cluster_id"], int(q.names["hashmap_caps"][2]))
    # kinds and AlterSV lists survive
    kind_gen = dict(zip(gen_cl, p.arrays["cluster_kind"]))
    assert all(kind_gen[c] == k for c, k in zip(q.names["cluster_id"], q.arrays["cluster_kind"]))
    assert q.scalars["n_sv"] == p.scalars["n_sv"] and sorted(q.arrays["sv_numchanged"]) == sorted(p.arrays["sv_numchanged"])


def test_counts_equal_literal_set_breakpoints(oracle, mbsv, tmp_path):
    """Count-based support bookkeeping == literal setBreakpoints (leaves -> owner -> assigned alignment contains
    the key -> greedy set cover), for random assignments, simple clusters and connected components."""
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=300, seed=77, conncomp_frac=0.5)
    paths = synth.write_reference_files(p, str(tmp_path))
    q = oracle.load_files(paths["clusters"], paths["assignments"], paths["experiments"])
    m = oracle.LoadedModel(paths["clusters"], paths["assignments"], paths["experiments"])
    a = q.arrays
    E, C, F = q.scalars["n_exp"], q.scalars["n_cluster"], q.scalars["n_frag"]
    nal = np.diff(a["frag_aln_ptr"])
    rng = np.random.default_rng(1)
    for _ in range(5):
        assign = np.array([rng.integers(-1, n) for n in nal], np.int32)
        # supports implied by the run path: a 0-iteration faithful run with this initial state, then tallies via
        # a 1-iteration window is overkill; recompute counts here the way the sampler does
        cnt = np.zeros((C, E), np.int64)
        for f in range(F):
            if assign[f] < 0:
                continue
            al = a["frag_aln_ptr"][f] + assign[f]
            for j in range(a["aln_esp_ptr"][al], a["aln_esp_ptr"][al + 1]):
                c = a["aln_esp_cluster"][j]
                if c >= 0:
                    cnt[c, a["aln_exp"][al]] += 1
        for c in range(C):
            lit = m.literal_supports(c, assign, E)
            if a["cluster_kind"][c] == 0:
                for x in range(E):
                    assert lit[x] == ([int(cnt[c, x])] if cnt[c, x] > 0 else [])
            else:
                # set cover covers every selected ESP exactly once in total
                for x in range(E):
                    assert sum(lit[x]) == cnt[c, x] and lit[x] == sorted(lit[x])


def _write(tmp_path, name, text):
    p = os.path.join(str(tmp_path), name)
    open(p, "w").write(text)
    return p


def test_malformed_inputs(oracle, tmp_path):
    exp = _write(tmp_path, "e.txt", "ExperimentLabel\tPseq\tLambdaD\nx 0.1 3\n")
    asg = ("FragmentID\tDiscordantPairs\tNumErrors\tBasesInAlignment\tExperimentLabel\n"
           "longread_1\tlongread_1_0.0-1.0\t5\t100\tx\nlongread_1\tlongread_1_0.1-1.1\t7\t100\tx\n"
           "longread_2\tlongread_2_0.0-1.0\t5\t100\tx\n")
    good_cl = "c1\t2\t10\tD\tlongread_1_0.0-1.0, longread_2_0.0-1.0\t1\t1\t1, 2\nc2\t1\t10\tD\tlongread_1_0.1-1.1\t1\t1\t1, 2\n"
    a, c = _write(tmp_path, "a.txt", asg), _write(tmp_path, "c.txt", good_cl)
    p = oracle.load_files(c, a, exp)
    assert p.scalars["n_frag"] == 2 and p.scalars["n_cluster"] == 2 and p.names["unsupported"] == ""
    # unknown experiment label: the reference exits (MultiBreakSV.java:355-358)
    bad = _write(tmp_path, "a2.txt", asg.replace("\tx\n", "\ty\n", 1))
    with pytest.raises(RuntimeError, match="not in Experiments file"):
        oracle.load_files(c, bad, exp)
    # blank line inside the clusters file: ArrayIndexOutOfBounds in the reference
    with pytest.raises(RuntimeError, match="ArrayIndexOutOfBounds"):
        oracle.load_files(_write(tmp_path, "c2.txt", good_cl.replace("\nc2", "\n\nc2")), a, exp)
    # an ESP listed in two clusters: the reference's incremental setBreakpoints becomes path dependent
    dup = good_cl + "c3\t1\t10\tD\tlongread_2_0.0-1.0\t1\t1\t1, 2\n"
    q = oracle.load_files(_write(tmp_path, "c3.txt", dup), a, exp)
    assert "more than one cluster" in q.names["unsupported"]
    # alignments whose ESPs are in no cluster are dropped and indices shift (quirk Q8); empty fragments vanish
    few = "c2\t1\t10\tD\tlongread_1_0.1-1.1\t1\t1\t1, 2\n"
    q = oracle.load_files(_write(tmp_path, "c4.txt", few), a, exp)
    assert q.names["frag_id"] == ["longread_1"] and q.scalars["n_aln"] == 1 and q.arrays["aln_numerrors"].tolist() == [7]
    # header row "cID" is skipped (MultiBreakSV.java:417)
    q = oracle.load_files(_write(tmp_path, "c5.txt", "cID\tn\tloc\ttype\tesps\n" + good_cl), a, exp)
    assert q.scalars["n_cluster"] == 2


def test_example_matchfile_quirk(oracle, example_problem):
    """initialset.txt lacks the longread_ prefix, so fa.contains(esp) never matches: the matchfile run
    of the manual starts with every fragment unassigned."""
    esps = open(os.path.join(EXAMPLE, "initialset.txt")).read().split()
    assert not any(e in example_problem.names["esp_names"] for e in esps)
    r = oracle.run(example_problem, "faithful", 100, init_assign=np.full(4, -1, np.int32))
    assert r.initial_assign[0].tolist() == [-1, -1, -1, -1]
