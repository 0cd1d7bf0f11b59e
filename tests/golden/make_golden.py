#!/usr/bin/env python
"""Regenerates tests/golden/example_expected.json and preproc_expected.json (the two shipped data sets).

    python tests/golden/make_golden.py

Inputs: tests/golden/example_infiles/ (the reference's MultiBreak-SV-example/infiles, data only).  The reference ships
no expected outputs and cannot run here (no JDK), so these fixtures are produced by the CPU oracle (oracle/), whose
agreement with SURVEY.md §8c's independently derived known-answers is checked in tests/test_oracle_kat.py.  They pin the
oracle AND the CUDA path against accidental drift: both must keep reproducing exactly these numbers.
"""
import hashlib
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as o  # noqa: E402


def hexd(x):
    return struct.pack(">d", float(x)).hex()


def record(p, runs):
    out = {"frag_id": p.names["frag_id"], "cluster_id": p.names["cluster_id"], "runs": {}}
    for mode, n, chains in runs:
        r = o.run(p, mode, n, n_chains=chains, trace=(chains == 1))
        rec = {"counters": r.counters.reshape(-1, 5).tolist(), "rng_state": [hex(int(x)) for x in r.rng_state.ravel()],
               "final_assign": r.final_assign.tolist(), "final_logprob_bits": [hexd(x) for x in r.final_logprob.ravel()],
               "initial_logprob_bits": [hexd(x) for x in r.initial_logprob.ravel()],
               "aln_count": r.aln_count.tolist(), "frag_err_count": r.frag_err_count.tolist(),
               "cluster_hist": r.cluster_hist.tolist()}
        if chains == 1:
            h = hashlib.sha256()
            h.update(r.trace_move_frag.tobytes()); h.update(r.trace_move_assign.tobytes()); h.update(r.trace_logprob.tobytes())
            rec["trace_sha256"] = h.hexdigest()
        out["runs"][f"{mode}-{n}-{chains}"] = rec
    return out


RUNS = (("faithful", 10000, 1), ("incremental", 10000, 1), ("incremental", 3000, 64), ("faithful", 137, 2))


def main():
    ex = os.path.join(HERE, "example_infiles")
    p = o.load_files(os.path.join(ex, "gasv.clusters"), os.path.join(ex, "assignments.txt"), os.path.join(ex, "experiments.txt"))
    with open(os.path.join(HERE, "example_expected.json"), "w") as fh:
        json.dump(record(p, RUNS), fh, indent=1)
    print("wrote example_expected.json")
    # the second shipped data set (Preprocessor-example/compressed-outfiles/example-output-run.tgz: out-MBSVinputs/ and
    # out-RunGASV/gasv.in.clusters): mostly single-alignment long reads, so AlterSV proposals are exercised on real data
    pre = os.path.join(HERE, "preproc_infiles")
    q = o.load_files(os.path.join(pre, "gasv.in.clusters"), os.path.join(pre, "assignments.txt"), os.path.join(pre, "experiments.txt"))
    rec = record(q, RUNS)
    rec["sv_cluster"] = [q.names["cluster_id"][c] for c in q.arrays["sv_cluster"].tolist()]
    rec["sv_numchanged"] = q.arrays["sv_numchanged"].tolist()
    # generateIndependentSubproblems.pl on its clusters file (run here from /root/reference when present): cluster ids per subset
    perl = "/root/reference/src/generateIndependentSubproblems.pl"
    if os.path.exists(perl):
        import subprocess
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            subprocess.run(["perl", perl, os.path.join(pre, "gasv.in.clusters"), td], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL)
            rec["perl_subsets"] = {f: sorted(l.split("\t")[0] for l in open(os.path.join(td, f)).read().splitlines())
                                   for f in sorted(os.listdir(td))}
    with open(os.path.join(HERE, "preproc_expected.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print("wrote preproc_expected.json")


if __name__ == "__main__":
    main()
