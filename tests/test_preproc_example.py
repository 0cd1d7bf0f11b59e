"""The reference's SECOND shipped data set (Preprocessor-example/compressed-outfiles/example-output-run.tgz:
out-MBSVinputs/{assignments,experiments}.txt, out-RunGASV/gasv.in.clusters and the shipped per-cluster
cluster-subproblems/iter*.clusters; data only, copied to tests/golden/preproc_infiles/).

Unlike MultiBreak-SV-example it is mostly single-alignment long reads, so `bpsWithNonSingletonSupport` is not empty and
AlterSV proposals (MultiBreakSV.java:1291-1353) run on real data: 441 of the first 10,000 iterations.  Golden vectors:
tests/golden/preproc_expected.json (oracle-made, tests/golden/make_golden.py; the partition part was written by the
reference's own generateIndependentSubproblems.pl)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, to_packed
from test_golden import RUNS, _check

PRE = os.path.join(GOLDEN, "preproc_infiles")
FILES = [os.path.join(PRE, "gasv.in.clusters"), os.path.join(PRE, "assignments.txt"), os.path.join(PRE, "experiments.txt")]
GOLD = json.load(open(os.path.join(GOLDEN, "preproc_expected.json")))


@pytest.fixture(scope="module")
def preproc_problem(oracle):
    return oracle.load_files(*FILES)


def test_shape_and_altersv_candidates(preproc_problem):
    p = preproc_problem
    sc = p.scalars
    # 36 long reads in the assignments file, 16 of them in the 10 clusters; longread_1 has two alignments
    assert (sc["n_frag"], sc["n_aln"], sc["n_cluster"], sc["n_exp"], sc["max_support"]) == (16, 17, 10, 1, 4)
    assert p.names["frag_id"] == GOLD["frag_id"] and p.names["cluster_id"] == GOLD["cluster_id"]
    # fillNonSingletonArray (:748-771): clusters with at least two ESPs ... whose single-alignment fragments AlterSV toggles
    assert [p.names["cluster_id"][c] for c in p.arrays["sv_cluster"]] == GOLD["sv_cluster"] == ["c5", "c6", "c10"]
    assert p.arrays["sv_numchanged"].tolist() == GOLD["sv_numchanged"] == [2, 3, 4]


@pytest.mark.parametrize("mode,n,chains", RUNS)
def test_oracle_reproduces_golden(oracle, preproc_problem, mode, n, chains):
    r = oracle.run(preproc_problem, mode, n, n_chains=chains, trace=(chains == 1))
    _check(r, GOLD["runs"][f"{mode}-{n}-{chains}"], chains)
    if chains == 1 and n == 10000:
        assert r.counters[0, 0].tolist() == [4557, 284, 441, 0, 5884]      # 441 AlterSV proposals, 2-4 fragments each


def test_host_parser_matches_oracle(mbsv, oracle, preproc_problem):
    from multibreak_sv_b200 import host
    for cf in (False, True):
        sv = host.MultiBreakSV(["--clusterfile", FILES[0], "--assignmentfile", FILES[1], "--experimentfile", FILES[2],
                                "--numiterations", "10000", "--perr", "0.001"] + (["--readclustersfirst"] if cf else [])).load()
        q = oracle.load_files(*FILES, clusters_first=cf)
        p = sv.packed()
        assert p.scalars == q.scalars
        for k in p.arrays:
            assert np.array_equal(p.arrays[k], q.arrays[k]), k
        assert sv.names(0, 16) == q.names["frag_id"] and sv.names(1, 10) == q.names["cluster_id"]
        assert sv.warning() == q.names["warning"]
    assert q.scalars == preproc_problem.scalars


def test_partition_matches_perl_and_the_shipped_split(mbsv, tmp_path):
    """Every cluster of this data set is its own subproblem.  Numbering = the Perl script's (subset k holds the smallest
    long-read number not yet placed: longread_1 -> c7 first); the shipped cluster-subproblems/iter*.clusters (numbered
    by another tool) hold the same ten one-line files."""
    from multibreak_sv_b200 import host
    out = str(tmp_path)
    n, rows = host.generateIndependentSubproblems(FILES[0], out)
    assert n == 10 and sorted(rows.tolist()) == list(range(1, 11))
    got = {f: sorted(l.split("\t")[0] for l in open(os.path.join(out, f)).read().splitlines()) for f in sorted(os.listdir(out))}
    assert got == GOLD["perl_subsets"]
    shipped = os.path.join(PRE, "cluster-subproblems")
    assert sorted(open(os.path.join(shipped, f)).read() for f in os.listdir(shipped)) == \
        sorted(open(os.path.join(out, f)).read() for f in os.listdir(out))


def test_one_shipped_subproblem_with_readclustersfirst(mbsv, oracle):
    """iter4.clusters (c6: three single-alignment reads) run the way the manual says (:464-473): the whole assignments
    file, --readclustersfirst."""
    from multibreak_sv_b200 import host
    c = os.path.join(PRE, "cluster-subproblems", "iter4.clusters")
    sv = host.MultiBreakSV(["--clusterfile", c, "--assignmentfile", FILES[1], "--experimentfile", FILES[2],
                            "--numiterations", "1000", "--perr", "0.001", "--readclustersfirst"]).load()
    q = oracle.load_files(c, FILES[1], FILES[2], clusters_first=True)
    p = sv.packed()
    assert p.scalars == q.scalars and p.scalars["n_frag"] == 3 and p.scalars["n_cluster"] == 1 and p.scalars["n_sv"] == 1
    assert sorted(sv.names(0, 3)) == ["longread_63", "longread_64", "longread_67"]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,n,chains", RUNS)
def test_cuda_reproduces_golden(mbsv, preproc_problem, mode, n, chains):
    dp = mbsv.DeviceProblem(to_packed(mbsv, preproc_problem))
    m = mbsv.MODE_REPLAY_EXACT if mode == "faithful" else mbsv.MODE_FAST
    g = dp.run(m, n, n_chains=chains, trace=(chains == 1), java_int_wrap=True)
    _check(g, GOLD["runs"][f"{mode}-{n}-{chains}"], chains)


@pytest.mark.gpu
def test_cli_file_to_file(mbsv, oracle, preproc_problem, tmp_path):
    """mbsv_cli on the data set: the move summary the reference prints (:842-843) and the long-burn-in files."""
    from multibreak_sv_b200 import host
    prefix = os.path.join(str(tmp_path), "pre")
    cli = os.path.join(os.path.dirname(os.path.abspath(host.__file__)), "mbsv_cli")
    r = subprocess.run([cli, "--clusterfile", FILES[0], "--assignmentfile", FILES[1], "--experimentfile", FILES[2],
                        "--numiterations", "10000", "--perr", "0.001", "--prefix", prefix, "--savespace"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "284 of 4557(" in r.stdout and "0 of 441(0.0) AlterSV Moves" in r.stdout
    assert not os.path.exists(prefix + ".mcmc.txt")                       # --savespace
    o = oracle.run(preproc_problem, "faithful", 10000)
    ref = host.MultiBreakSV(["--clusterfile", FILES[0], "--assignmentfile", FILES[1], "--experimentfile", FILES[2],
                             "--numiterations", "10000", "--perr", "0.001", "--prefix", prefix + "_ref"]).load()
    ref.setResults(o.aln_count, o.frag_err_count, o.cluster_hist)
    ref.writeFinalFiles()
    for ext in (".finalbreakpoints.longburnin", ".finalassignment.longburnin"):
        assert open(prefix + ext).read() == open(prefix + "_ref" + ext).read()


@pytest.mark.gpu
def test_cli_partition_mode(mbsv, oracle, tmp_path):
    """mbsv_cli --partition: partition + load in memory, one GPU batch over the ten subproblems, concatenated output files;
    expected = the oracle's run of the same batch through the same batch writer."""
    from multibreak_sv_b200 import host
    prefix = os.path.join(str(tmp_path), "part")
    cli = os.path.join(os.path.dirname(os.path.abspath(host.__file__)), "mbsv_cli")
    r = subprocess.run([cli, "--clusterfile", FILES[0], "--assignmentfile", FILES[1], "--experimentfile", FILES[2],
                        "--numiterations", "10000", "--perr", "0.001", "--prefix", prefix, "--partition"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "10 independent subproblems (0 empty skipped): 16 fragments, 17 alignments, 10 clusters" in r.stdout
    b = host.PartitionedBatch(*FILES)
    p = b.packed
    o = oracle.run(oracle.Problem(p.scalars, p.arrays), "faithful", 10000)
    assert f"{int(o.counters[:, :, 1].sum())} of {int(o.counters[:, :, 0].sum())}(" in r.stdout
    b.writeFinalFiles(prefix + "_ref", 1000, o.aln_count, o.frag_err_count, o.cluster_hist)
    for ext in (".finalbreakpoints.longburnin", ".finalassignment.longburnin"):
        assert open(prefix + ext).read() == open(prefix + "_ref" + ext).read()
    assert len(open(prefix + ".finalbreakpoints.longburnin").read().splitlines()) == 11


@pytest.mark.gpu
def test_python_partition_mode_equals_the_cli(mbsv, tmp_path):
    """python -m multibreak_sv_b200.host --partition and mbsv_cli --partition write the same files."""
    from multibreak_sv_b200 import host
    args = ["--clusterfile", FILES[0], "--assignmentfile", FILES[1], "--experimentfile", FILES[2], "--numiterations", "5000",
            "--perr", "0.001", "--partition", "--chains", "3"]
    py = os.path.join(str(tmp_path), "py")
    host.main(args + ["--prefix", py])
    cli = os.path.join(os.path.dirname(os.path.abspath(host.__file__)), "mbsv_cli")
    cp = os.path.join(str(tmp_path), "cli")
    r = subprocess.run([cli] + args + ["--prefix", cp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for ext in (".finalbreakpoints.longburnin", ".finalassignment.longburnin"):
        assert open(py + ext).read() == open(cp + ext).read()
    assert open(py + ".finalbreakpoints.longburnin").read().splitlines()[1].split("\t")[1] == "1500"      # 3 chains x burnin 500
