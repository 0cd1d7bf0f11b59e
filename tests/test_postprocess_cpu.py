"""Cluster scores: multibreak_sv_b200.postprocess against the reference's own quick-post-process.py.

The expected files under tests/golden/postprocess/ were written by the reference script itself (executed from
/root/reference by tests/golden/make_postprocess_golden.py); the inputs beside them are this repository's writers'
output for the CPU oracle's tallies.  No GPU, no /root/reference at test time.
"""
import os

import numpy as np
import pytest

from conftest import EXAMPLE

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postprocess")


def _case(name):
    d = os.path.join(GOLD, name)
    clusters = os.path.join(d, "gasv.clusters")
    if not os.path.exists(clusters):
        clusters = os.path.join(EXAMPLE, "gasv.clusters")
    return d, clusters, int(open(os.path.join(d, "minsupport")).read())


@pytest.mark.parametrize("name", ["example", "conncomp"])
def test_text_path_matches_the_reference_script(name, tmp_path):
    from multibreak_sv_b200 import postprocess
    d, clusters, minsupport = _case(name)
    out = os.path.join(str(tmp_path), "scores.txt")
    postprocess.quick_post_process(clusters, os.path.join(d, "run.finalbreakpoints.longburnin"),
                                   os.path.join(d, "run.finalassignment.longburnin"), minsupport, out)
    assert open(out).read() == open(os.path.join(d, "expected.txt")).read()


def test_cli_and_errors(tmp_path, capsys):
    from multibreak_sv_b200 import postprocess
    d, clusters, minsupport = _case("example")
    out = os.path.join(str(tmp_path), "o.txt")
    postprocess.main(["--clustersfile", clusters, "--finalbreakpointfile", os.path.join(d, "run.finalbreakpoints.longburnin"),
                      "--finalassignmentfile", os.path.join(d, "run.finalassignment.longburnin"), "--outfile", out])
    assert capsys.readouterr().out == "Wrote to %s\n" % out
    assert open(out).read() == open(os.path.join(d, "expected.txt")).read()          # default --minsupport is 2
    with pytest.raises(SystemExit) as e:
        postprocess.main(["--clustersfile", clusters])
    assert "requires clusters file" in str(e.value)
    # a cluster whose id is missing from the finalbreakpoints file is the reference's ERROR exit (:119-120)
    bp = open(os.path.join(d, "run.finalbreakpoints.longburnin")).read().splitlines(True)
    short = os.path.join(str(tmp_path), "short.bp")
    open(short, "w").writelines(l for l in bp if not l.startswith("c4126\t"))
    with pytest.raises(SystemExit) as e:
        postprocess.quick_post_process(clusters, short, os.path.join(d, "run.finalassignment.longburnin"), 2, out)
    assert 'cid="c4126"' in str(e.value)


@pytest.mark.parametrize("name,iters", [("example", 10000), ("conncomp", 4000)])
def test_scores_from_tallies_equal_the_text_path(mbsv, oracle, name, iters, tmp_path):
    """The same scores without the text round trip, from the integer tallies the sampler returns."""
    from multibreak_sv_b200 import host, postprocess, synth
    d, clusters, minsupport = _case(name)
    if name == "example":
        files = dict(clusters=clusters, assignments=os.path.join(EXAMPLE, "assignments.txt"),
                     experiments=os.path.join(EXAMPLE, "experiments.txt"))
    else:
        files = synth.write_reference_files(synth.generate("cfg4", n_frag=60, seed=99, conncomp_frac=0.5), str(tmp_path))
        assert open(files["clusters"]).read() == open(clusters).read()
    sv = host.MultiBreakSV(["--clusterfile", files["clusters"], "--assignmentfile", files["assignments"],
                            "--experimentfile", files["experiments"], "--numiterations", str(iters), "--perr", "0.001"]).load()
    p = sv.packed()
    r = oracle.run(oracle.load_files(files["clusters"], files["assignments"], files["experiments"]), "faithful", iters)
    a = p.arrays
    sc = p.scalars
    got = postprocess.scores_from_tallies(
        sv.names(1, sc["n_cluster"]), a["cluster_hist_ptr"], r.cluster_hist, sv.sortedFragments(sc["n_frag"]), a["frag_aln_ptr"],
        a["aln_esp_ptr"], sv.names(3, sc["n_aln_esp"]), r.aln_count, sv.burnin, postprocess.read_clusters(clusters), minsupport)
    want = [l.split("\t") for l in open(os.path.join(d, "expected.txt")).read().splitlines()[1:]]
    assert [w[0] for w in want] == list(got)
    for w in want:
        assert ("%.4f" % got[w[0]][0], "%.4f" % got[w[0]][1]) == (w[1], w[2]), w[0]
