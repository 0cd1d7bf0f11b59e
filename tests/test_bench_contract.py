"""bench.py: the JSON line of both arms must carry the contract keys (the reference arm runs on the CPU; the GPU arm prints the same
keys plus roofline/clocks; it is exercised on the GPU box)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n-frag", "20000",
                        "--cpu-iterations", "2e6", "--iters", "500", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "fragment-reassignments/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cfg4-mixed-5M" and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--n-frag", "20000"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_gpu_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--n-frag", "30000", "--iters", "400", "--steps", "2",
                        "--warmup", "3", "--cpu-iterations", "2e6"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["scaling"] == "weak"
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "sm-issue" and rf["peak"] == 100.0 and "traffic" in rf
    hb = rf["hbm_notional"]
    assert hb["bound"] == "hbm" and hb["peak"] > 0 and abs(hb["frac"] - hb["achieved"] / hb["peak"]) < 1e-9
    assert sum(c["reassignments"] for c in d["per_class"]) * d["steps"] > 0
    assert d["e2e"]["steps"] == d["steps"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_stratified_sample_is_a_valid_sub_batch(oracle):
    """bench.py's CPU sample: the gathered sub-batch of arbitrary subproblems runs exactly like those subproblems inside
    the full batch (tallies of the picked subproblems, counters, RNG states)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from multibreak_sv_b200 import synth
    p = synth.generate("cfg4", n_frag=4000, seed=5, conncomp_frac=0.3)
    a = p.arrays
    full = oracle.run(oracle.Problem(p.scalars, a), "incremental", 400, n_chains=2, nthreads=4, java_int_wrap=False)
    pick = np.array([0, 1, 2, 7, 8, 30, 31, 32, 33, p.scalars["n_sub"] - 1])
    q = bench.gather_subproblems(p, pick)
    sub = oracle.run(oracle.Problem(q.scalars, q.arrays), "incremental", 400, n_chains=2, nthreads=4, java_int_wrap=False)
    assert np.array_equal(sub.counters, full.counters[pick]) and np.array_equal(sub.rng_state, full.rng_state[pick])
    fptr, aptr, hptr = a["sub_frag_ptr"], a["frag_aln_ptr"], a["cluster_hist_ptr"]
    cptr = a["sub_cluster_ptr"]
    al = np.concatenate([full.aln_count[aptr[fptr[s]]:aptr[fptr[s + 1]]] for s in pick])
    fe = np.concatenate([full.frag_err_count[fptr[s]:fptr[s + 1]] for s in pick])
    hi = np.concatenate([full.cluster_hist[hptr[cptr[s]]:hptr[cptr[s + 1]]] for s in pick])
    assert np.array_equal(sub.aln_count, al) and np.array_equal(sub.frag_err_count, fe) and np.array_equal(sub.cluster_hist, hi)
    strata = bench.stratified_sample(p, 2, 400, 2 * 400 * 60, 4)
    assert sum(N for N, *_ in strata) == p.scalars["n_sub"] and all(n >= min(N, 8) for N, n, *_ in strata)


def test_generator_does_not_need_the_product_library(mbsv):
    """synth fills logprobcov with its own fdlibm log; it must agree bit for bit with mbsv_fill_logprobcov."""
    import ctypes as C
    import numpy as np
    from multibreak_sv_b200 import synth
    L = mbsv.load_library()
    for lam, ms in ((10.0, 60), (3.0, 17), (0.7, 5)):
        T = np.zeros(ms + 1)
        L.mbsv_fill_logprobcov(lam, ms, T.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.array_equal(T, synth.logprobcov_table(lam, ms))
    src = open(os.path.join(ROOT, "multibreak_sv_b200", "synth.py")).read()
    assert "load_library" not in src
