"""The C-ABI library loads without a GPU, exports every symbol include/mbsv_b200.h declares, validates
descriptors, and refuses to compute (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, to_packed


def _declared_functions():
    names = set()
    for h in ("mbsv_b200.h", "mbsv_host.h"):
        hdr = open(os.path.join(ROOT, "include", h)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        names |= set(re.findall(r"\b(mbsv_[a-z0-9_]+)\s*\(", hdr))
    return sorted(names)


def test_header_symbols_exported(mbsv):
    L = mbsv.load_library()
    names = _declared_functions()
    assert "mbsv_run" in names and "mbsv_problem_create" in names and "mbsv_host_run_mcmc" in names
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mbsv_b200.h but not exported"
    assert L.mbsv_abi_version() == 1


def test_struct_layouts_match_header(mbsv):
    """ctypes mirrors vs the C compiler's view of include/mbsv_b200.h."""
    from multibreak_sv_b200 import _abi
    src = r'''
    #include "mbsv_b200.h"
    #include <stdio.h>
    #include <stddef.h>
    int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(mbsv_problem_desc), sizeof(mbsv_run_params),
      sizeof(mbsv_results), offsetof(mbsv_problem_desc, sub_frag_ptr), offsetof(mbsv_run_params, init_assign),
      offsetof(mbsv_results, kernel_ms)); return 0; }'''
    d = os.path.join(ROOT, ".pytest_cache")
    os.makedirs(d, exist_ok=True)
    cfile, exe = os.path.join(d, "layout.c"), os.path.join(d, "layout")
    open(cfile, "w").write(src)
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
    got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [C.sizeof(_abi.ProblemDesc), C.sizeof(_abi.RunParams), C.sizeof(_abi.Results),
            _abi.ProblemDesc.sub_frag_ptr.offset, _abi.RunParams.init_assign.offset, _abi.Results.kernel_ms.offset]
    assert got == want
    # the literal offsets integration/java/MultiBreakSVGpu.java writes through (Panama FFM MemorySegment.set)
    assert _abi.ProblemDesc.sub_frag_ptr.offset == 56 and C.sizeof(_abi.ProblemDesc) == 14 * 4 + 25 * 8
    assert (_abi.RunParams.perr.offset, _abi.RunParams.seed.offset, _abi.RunParams.device.offset,
            _abi.RunParams.init_assign.offset, _abi.RunParams.trace.offset, _abi.RunParams.memo_states.offset,
            C.sizeof(_abi.RunParams)) == (16, 24, 32, 40, 48, 52, 56)
    assert _abi.Results.kernel_ms.offset == 14 * 8 and C.sizeof(_abi.Results) == 14 * 8 + 8
    java = open(os.path.join(ROOT, "integration", "java", "MultiBreakSVGpu.java")).read()
    for needle in ("arena.allocate(14 * 4 + 25 * 8, 8)", "56L + 8L * i", "arena.allocate(56, 8)", "rp.set(JAVA_INT, 52, 65536)",
                   "arena.allocate(14 * 8 + 8, 8)"):
        assert needle in java


def test_no_cpu_fallback(mbsv, example_problem):
    L = mbsv.load_library()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert L.mbsv_device_count() == -2
    with pytest.raises(mbsv.MbsvError) as e:
        mbsv.DeviceProblem(to_packed(mbsv, example_problem))
    assert e.value.status == -2


def test_descriptor_validation(mbsv, example_problem):
    L = mbsv.load_library()
    from multibreak_sv_b200 import _abi
    p = to_packed(mbsv, example_problem)
    h = C.c_void_p()

    def create(pp):
        d = pp.desc()
        return L.mbsv_problem_create(C.byref(d), 0, C.byref(h))

    bad = to_packed(mbsv, example_problem)
    bad.scalars["abi_version"] = 99
    assert create(bad) == -1 and b"abi_version" in L.mbsv_last_error()
    bad = to_packed(mbsv, example_problem)
    bad.arrays["frag_aln_ptr"] = p.arrays["frag_aln_ptr"][::-1].copy()
    assert create(bad) == -1
    bad = to_packed(mbsv, example_problem)
    bad.arrays["cluster_kind"] = np.full_like(p.arrays["cluster_kind"], 2)
    assert create(bad) == -1
    bad = to_packed(mbsv, example_problem)                 # connected components need their leaf/root arrays
    bad.arrays["cluster_kind"] = np.ones_like(p.arrays["cluster_kind"])
    bad.arrays["root_leaf"] = p.arrays["root_leaf"] + 100
    assert create(bad) == -1 and b"root_leaf" in L.mbsv_last_error()
    bad = to_packed(mbsv, example_problem)
    so = p.arrays["supports_order"].copy()
    so[0] = so[1]
    bad.arrays["supports_order"] = so
    assert create(bad) == -1
    bad = to_packed(mbsv, example_problem)
    bad.scalars["max_support"] = 2                      # the 26-ESP cluster would overflow logprobcov
    bad.arrays["logprobcov"] = p.arrays["logprobcov"][:3]
    assert create(bad) == -6
    assert L.mbsv_problem_create(None, 0, C.byref(h)) == -1
    assert L.mbsv_status_string(-5) == b"reference invariant violated"
    # an empty problem (no subproblem / fragment / alignment) is an argument error: the reference cannot run it either
    # (nextInt(0) and an index into an empty fragment list throw, MultiBreakSV.java:1013,1198)
    for key in ("n_sub", "n_frag", "n_aln"):
        bad = to_packed(mbsv, example_problem)
        bad.scalars[key] = 0
        assert create(bad) == -1 and b"out of range" in L.mbsv_last_error()
    # Experiment.logprobcov[0] is always 0 in the reference; the Delta arithmetic relies on it
    bad = to_packed(mbsv, example_problem)
    lp = p.arrays["logprobcov"].copy()
    lp[0] = -1.5
    bad.arrays["logprobcov"] = lp
    assert create(bad) == -1 and b"logprobcov" in L.mbsv_last_error()
    # a fragment without any alignment is ragged input the packer rejects too
    bad = to_packed(mbsv, example_problem)
    fp = p.arrays["frag_aln_ptr"].copy()
    fp[1] = fp[0]
    bad.arrays["frag_aln_ptr"] = fp
    assert create(bad) == -1


def test_host_math_matches_oracle(mbsv, oracle):
    L = mbsv.load_library()
    rng = np.random.default_rng(3)
    for x in np.concatenate([rng.uniform(1e-9, 1e7, 20000), np.arange(1, 2000, dtype=float)]):
        assert L.mbsv_log(float(x)) == oracle.jlog(float(x))
    for x in rng.uniform(-750, 710, 20000):
        assert L.mbsv_exp(float(x)) == oracle.jexp(float(x))
    # the selected (formerly early-return) cases of e_exp.c: thresholds, subnormal results, tiny arguments, non-finite
    bits = lambda v: np.float64(v).view(np.uint64)
    special = [0.0, -0.0, 1e-10, -1e-10, 2.0 ** -28, 2.0 ** -29, -2.0 ** -29, 709.782712893384, 709.7827128933841, 709.79,
               -745.1332191019411, -745.1332191019412, -745.14, -708.0, -708.4, -720.0, -744.9, 1e300, -1e300, 5e3, -5e3,
               float("inf"), float("-inf"), float("nan"), 0.5 * np.log(2.0), 1.5 * np.log(2.0), -0.34, 0.35, 1.04]
    special += list(rng.uniform(-746, -707, 2000)) + list(rng.uniform(-1e-7, 1e-7, 2000)) + list(rng.uniform(-1e5, 1e5, 2000))
    for x in special:
        assert bits(L.mbsv_exp(float(x))) == bits(oracle.jexp(float(x))), x
    t = np.zeros(27)
    assert L.mbsv_fill_logprobcov(3.0, 26, t.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert t[0] == 0.0 and abs(t[3] - (-3 + 3 * np.log(3) - np.log(6))) < 1e-12


def test_no_fma_contraction_in_ptx(mbsv):
    """The sampler must not be compiled with FMA contraction (jmath.cuh): every f64 op in PTX is .rn add/mul/div."""
    from multibreak_sv_b200 import _abi
    try:
        ptx = subprocess.check_output(["cuobjdump", "-ptx", _abi.LIB_PATH], stderr=subprocess.DEVNULL).decode()
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump unavailable")
    if "mul.rn.f64" not in ptx:
        pytest.skip("library was built without embedded PTX")
    assert "fma.rn.f64" not in ptx and "mad.rn.f64" not in ptx
    assert " mul.f64" not in ptx and " add.f64" not in ptx     # unqualified ops could still be contracted by ptxas
