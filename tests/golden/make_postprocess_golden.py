#!/usr/bin/env python
"""Regenerates tests/golden/postprocess/: inputs and expected output of the cluster-score post-processing.

    python tests/golden/make_postprocess_golden.py          (build container only: needs /root/reference)

For each case: the three input files -> host mirror (parser) -> CPU oracle tallies -> the host writers'
.finalbreakpoints.longburnin / .finalassignment.longburnin (committed as the fixture's inputs) -> the REFERENCE's own
src/quick-post-process.py, executed here from /root/reference, -> expected.txt.  The script is Python 2; the single
`print` statement on its last line is rewritten in memory so that Python 3 can compile it -- nothing else is touched and
no reference source is written into this repository.  Under Python 3 its dict iteration is insertion (file) order.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/src/quick-post-process.py"

from multibreak_sv_b200 import host, synth  # noqa: E402
from oracle import oracle as o  # noqa: E402

RUNNER = r"""
import re, sys
src = open(sys.argv[1]).read()
src, n = re.subn(r"(?m)^(\s*)print '([^']*)' % \(([^)]*)\)\s*$", r"\1print('\2' % (\3))", src)
assert n == 1, n
sys.argv = sys.argv[1:]
exec(compile(src, sys.argv[0], "exec"))
"""


def case(name, clusters, assignments, experiments, iters, minsupport):
    out = os.path.join(HERE, "postprocess", name)
    os.makedirs(out, exist_ok=True)
    sv = host.MultiBreakSV(["--clusterfile", clusters, "--assignmentfile", assignments, "--experimentfile", experiments,
                            "--numiterations", str(iters), "--perr", "0.001", "--prefix", os.path.join(out, "run")]).load()
    p = o.load_files(clusters, assignments, experiments)
    r = o.run(p, "faithful", iters)
    sv.setResults(r.aln_count, r.frag_err_count, r.cluster_hist)
    sv.writeFinalFiles()
    subprocess.run([sys.executable, "-c", RUNNER, REF, "--clustersfile", clusters, "--finalbreakpointfile",
                    os.path.join(out, "run.finalbreakpoints.longburnin"), "--finalassignmentfile",
                    os.path.join(out, "run.finalassignment.longburnin"), "--minsupport", str(minsupport), "--outfile",
                    os.path.join(out, "expected.txt")], check=True)
    with open(os.path.join(out, "minsupport"), "w") as fh:
        fh.write(f"{minsupport}\n")
    print("wrote", out)


def main():
    ex = os.path.join(HERE, "example_infiles")
    # (the example's clusters file is tests/golden/example_infiles/gasv.clusters; it is not duplicated here)
    case("example", os.path.join(ex, "gasv.clusters"), os.path.join(ex, "assignments.txt"), os.path.join(ex, "experiments.txt"),
         10000, 2)
    # a synthetic batch with connected components (dotted maximal-cluster rows map to their component's frequency)
    tmp = os.path.join(HERE, "postprocess", "conncomp")
    os.makedirs(tmp, exist_ok=True)
    g = synth.generate("cfg4", n_frag=60, seed=99, conncomp_frac=0.5)
    paths = synth.write_reference_files(g, tmp)
    case("conncomp", paths["clusters"], paths["assignments"], paths["experiments"], 4000, 1)
    for k in ("assignments", "experiments"):
        os.remove(paths[k])


if __name__ == "__main__":
    main()
