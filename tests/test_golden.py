"""Committed golden vectors (tests/golden/example_expected.json, made by tests/golden/make_golden.py from the oracle)
for the shipped example: the oracle must keep reproducing them (CPU), and so must the CUDA path (GPU)."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, to_packed

GOLD = json.load(open(os.path.join(GOLDEN, "example_expected.json")))
RUNS = [("faithful", 10000, 1), ("incremental", 10000, 1), ("incremental", 3000, 64), ("faithful", 137, 2)]


def _check(r, rec, chains):
    assert r.counters.reshape(-1, 5).tolist() == rec["counters"]
    assert [hex(int(x)) for x in r.rng_state.ravel()] == rec["rng_state"]
    assert r.final_assign.tolist() == rec["final_assign"]
    assert [struct.pack(">d", float(x)).hex() for x in r.final_logprob.ravel()] == rec["final_logprob_bits"]
    assert [struct.pack(">d", float(x)).hex() for x in r.initial_logprob.ravel()] == rec["initial_logprob_bits"]
    assert r.aln_count.tolist() == rec["aln_count"] and r.frag_err_count.tolist() == rec["frag_err_count"]
    assert r.cluster_hist.tolist() == rec["cluster_hist"]
    if chains == 1:
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(r.trace_move_frag).tobytes()); h.update(np.ascontiguousarray(r.trace_move_assign).tobytes())
        h.update(np.ascontiguousarray(r.trace_logprob).tobytes())
        assert h.hexdigest() == rec["trace_sha256"]


@pytest.mark.parametrize("mode,n,chains", RUNS)
def test_oracle_reproduces_golden(oracle, example_problem, mode, n, chains):
    assert example_problem.names["frag_id"] == GOLD["frag_id"] and example_problem.names["cluster_id"] == GOLD["cluster_id"]
    _check(oracle.run(example_problem, mode, n, n_chains=chains, trace=(chains == 1)), GOLD["runs"][f"{mode}-{n}-{chains}"], chains)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,n,chains", RUNS)
def test_cuda_reproduces_golden(mbsv, example_problem, mode, n, chains):
    dp = mbsv.DeviceProblem(to_packed(mbsv, example_problem))
    m = mbsv.MODE_REPLAY_EXACT if mode == "faithful" else mbsv.MODE_FAST
    g = dp.run(m, n, n_chains=chains, trace=(chains == 1), java_int_wrap=True)
    _check(g, GOLD["runs"][f"{mode}-{n}-{chains}"], chains)
