#!/usr/bin/env python
"""Regenerates tests/golden/example_expected.json and the expected output files of the shipped example.

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


def main():
    ex = os.path.join(HERE, "example_infiles")
    p = o.load_files(os.path.join(ex, "gasv.clusters"), os.path.join(ex, "assignments.txt"), os.path.join(ex, "experiments.txt"))
    out = {"frag_id": p.names["frag_id"], "cluster_id": p.names["cluster_id"], "runs": {}}
    for mode, n, chains in (("faithful", 10000, 1), ("incremental", 10000, 1), ("incremental", 3000, 64), ("faithful", 137, 2)):
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
    with open(os.path.join(HERE, "example_expected.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote example_expected.json")


if __name__ == "__main__":
    main()
