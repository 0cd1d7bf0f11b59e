"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/libmbsv_oracle.so, the CPU restatement of the reference sampler
(MultiBreakSV.java / MultiBreakSVAssignmentMatrix.java).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  The product package
(multibreak_sv_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmbsv_oracle.so")
_lib = None

I32P = C.POINTER(C.c_int32)
U8P = C.POINTER(C.c_uint8)
U32P = C.POINTER(C.c_uint32)
I64P = C.POINTER(C.c_int64)
U64P = C.POINTER(C.c_uint64)
F64P = C.POINTER(C.c_double)

# Field order/type is the contract shared with include/mbsv_b200.h (mbsv_problem_desc).
DESC_SCALARS = ["abi_version", "n_sub", "n_frag", "n_aln", "n_aln_esp", "n_cluster", "n_exp", "max_support",
                "n_sv", "n_sv_frag", "n_leaf", "n_root", "n_root_leaf", "n_hist"]
DESC_ARRAYS = [("sub_frag_ptr", np.int32), ("sub_cluster_ptr", np.int32), ("sub_sv_ptr", np.int32),
               ("frag_aln_ptr", np.int32), ("aln_numerrors", np.int32), ("aln_len", np.int32),
               ("aln_exp", np.int32), ("aln_esp_ptr", np.int32), ("aln_esp_cluster", np.int32),
               ("aln_esp_leaf", np.int32), ("cluster_kind", np.uint8), ("supports_order", np.int32),
               ("cluster_hist_ptr", np.int32), ("cluster_leaf_ptr", np.int32), ("leaf_owner", np.int32),
               ("cluster_root_ptr", np.int32), ("root_leaf_ptr", np.int32), ("root_leaf", np.int32),
               ("sv_cluster", np.int32), ("sv_frag_ptr", np.int32), ("sv_frag", np.int32),
               ("sv_numchanged", np.int32), ("exp_pseq", np.float64), ("exp_lambda", np.float64),
               ("logprobcov", np.float64)]
_CT = {np.int32: I32P, np.uint8: U8P, np.float64: F64P}


class OrcDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in DESC_SCALARS] + [(n, _CT[t]) for n, t in DESC_ARRAYS]


class OrcParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n_chains", C.c_int32), ("num_iters", C.c_int32), ("burnin", C.c_int32),
                ("perr", C.c_double), ("seed", C.c_int64), ("device", C.c_int32), ("java_int_wrap", C.c_int32),
                ("init_assign", I32P), ("trace", C.c_int32), ("memo_states", C.c_int32)]


class OrcResults(C.Structure):
    _fields_ = [("aln_count", U32P), ("frag_err_count", U32P), ("cluster_hist", U32P), ("final_assign", I32P),
                ("final_logprob", F64P), ("counters", I64P), ("rng_state", U64P), ("error", I32P),
                ("trace_logprob", F64P), ("trace_move_frag", I32P), ("trace_move_assign", I32P),
                ("initial_assign", I32P), ("initial_logprob", F64P), ("totals", I64P), ("kernel_ms", C.c_double)]


def build(force: bool = False) -> str:
    """Compile the oracle (g++, seconds).  Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in ("mbsv_oracle.cpp", "jsemantics.hpp", "refmodel.hpp", "sampler.hpp")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    if "_native" not in _LIB_PATH:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def use_native_build() -> str:
    """Switch this process to the -O3 -march=native build of the same sources (bench.py's CPU baseline).  It is compiled
    on the machine that runs it, into a directory named after that machine's CPU flags, and never shipped (.gpurunignore):
    results are bit-identical to the portable build (floating-point contraction stays off)."""
    global _LIB_PATH, _lib
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    rel = os.path.join("_native", hashlib.sha1(flags.encode()).hexdigest()[:12], "libmbsv_oracle_native.so")
    subprocess.check_call(["make", "-C", _HERE, "-s", "native", f"NATIVE={rel}"])
    _LIB_PATH = os.path.join(_HERE, rel)
    _lib = None
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_run.argtypes = [C.POINTER(OrcDesc), C.POINTER(OrcParams), C.POINTER(OrcResults), C.c_int]
        L.orc_run.restype = C.c_int
        L.orc_enumerate.argtypes = [C.POINTER(OrcDesc), C.c_int, C.c_double, C.c_int, F64P, C.c_int64]
        L.orc_enumerate.restype = C.c_int64
        L.orc_log.argtypes = [C.c_double]; L.orc_log.restype = C.c_double
        L.orc_exp.argtypes = [C.c_double]; L.orc_exp.restype = C.c_double
        L.orc_random_doubles.argtypes = [C.c_int64, C.c_int, F64P]
        L.orc_random_nextint.argtypes = [C.c_int64]; L.orc_random_nextint.restype = C.c_int32
        L.orc_random_bounded.argtypes = [C.c_int64, C.c_int32, C.c_int, I32P]
        L.orc_string_hash.argtypes = [C.c_char_p]; L.orc_string_hash.restype = C.c_int32
        L.orc_hashmap_order.argtypes = [C.c_char_p, C.c_int, U8P, I32P, I32P, I32P]
        L.orc_hashmap_order.restype = C.c_int
        L.orc_load.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.orc_load.restype = C.c_void_p
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_i32_len.argtypes = [C.c_void_p, C.c_char_p]; L.orc_i32_len.restype = C.c_int
        L.orc_i32_get.argtypes = [C.c_void_p, C.c_char_p, I32P]; L.orc_i32_get.restype = C.c_int
        L.orc_scalars.argtypes = [C.c_void_p, I32P]
        L.orc_u8_cluster_kind.argtypes = [C.c_void_p, U8P]; L.orc_u8_cluster_kind.restype = C.c_int
        L.orc_f64_get.argtypes = [C.c_void_p, C.c_char_p, F64P]; L.orc_f64_get.restype = C.c_int
        L.orc_str_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int]; L.orc_str_get.restype = C.c_char_p
        L.orc_literal_supports.argtypes = [C.c_void_p, C.c_int, I32P, I32P, C.c_int]
        L.orc_literal_supports.restype = C.c_int
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ct) if a is not None else ct()


# ------------------------------------------------------------------------------------------------
# Java-semantics hooks
# ------------------------------------------------------------------------------------------------
def jlog(x: float) -> float:
    return lib().orc_log(x)


def jexp(x: float) -> float:
    return lib().orc_exp(x)


def random_doubles(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.float64)
    lib().orc_random_doubles(seed, n, _p(out, F64P))
    return out


def random_nextint(seed: int) -> int:
    return lib().orc_random_nextint(seed)


def random_bounded(seed: int, bound: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.int32)
    lib().orc_random_bounded(seed, bound, n, _p(out, I32P))
    return out


def string_hash(s: str) -> int:
    return lib().orc_string_hash(s.encode())


def hashmap_order(keys, removed=()):
    """keySet() iteration order of a java.util.HashMap<String,?> after put(keys...) then remove(removed...).
    Returns (ordered list of surviving keys, table capacity, tree_bin flag)."""
    blob = b"".join(k.encode() + b"\0" for k in keys)
    rem = set(removed)
    mask = np.array([1 if k in rem else 0 for k in keys], np.uint8)
    out = np.zeros(max(1, len(keys)), np.int32)
    cap = C.c_int32(0)
    tb = C.c_int32(0)
    n = lib().orc_hashmap_order(blob, len(keys), _p(mask, U8P), _p(out, I32P), C.byref(cap), C.byref(tb))
    return [keys[i] for i in out[:n]], cap.value, bool(tb.value)


# ------------------------------------------------------------------------------------------------
# Problems
# ------------------------------------------------------------------------------------------------
@dataclass
class Problem:
    """Flat arrays of one or more independent problems (same field names as mbsv_problem_desc)."""
    scalars: dict
    arrays: dict
    names: dict = field(default_factory=dict)   # frag_id, cluster_id, exp_label, sorted_frag, aln_esp_name ...

    def desc(self, struct_type=OrcDesc):
        d = struct_type()
        for k in DESC_SCALARS:
            setattr(d, k, int(self.scalars[k]))
        self._keep = []
        for k, t in DESC_ARRAYS:
            a = np.ascontiguousarray(self.arrays[k], dtype=t)
            self._keep.append(a)
            ptr_t = dict(struct_type._fields_)[k]
            setattr(d, k, a.ctypes.data_as(ptr_t) if a.size else ptr_t())
        return d


def load_files(clusterfile: str, assignmentfile: str, experimentfile: str, clusters_first: bool = False) -> Problem:
    """The reference's input side (MultiBreakSV.java:114-128) up to fillNonSingletonArray, packed."""
    L = lib()
    err = C.create_string_buffer(512)
    h = L.orc_load(clusterfile.encode(), assignmentfile.encode(), experimentfile.encode(), int(clusters_first), err, 512)
    if not h:
        raise RuntimeError(err.value.decode())
    try:
        sc = np.zeros(8, np.int32)
        L.orc_scalars(h, _p(sc, I32P))
        F, A, K, Cn, E, P, maxsup, tree = (int(x) for x in sc)

        def gi(name):
            n = L.orc_i32_len(h, name.encode())
            out = np.zeros(max(n, 0), np.int32)
            if n > 0:
                L.orc_i32_get(h, name.encode(), _p(out, I32P))
            return out

        def gf(name, n):
            out = np.zeros(n, np.float64)
            if n:
                L.orc_f64_get(h, name.encode(), _p(out, F64P))
            return out

        def gs(name, n):
            return [L.orc_str_get(h, name.encode(), i).decode() for i in range(n)]

        kind = np.zeros(Cn, np.uint8)
        if Cn:
            L.orc_u8_cluster_kind(h, _p(kind, U8P))
        arrays = {k: gi(k) for k in ["frag_aln_ptr", "aln_numerrors", "aln_len", "aln_exp", "aln_esp_ptr",
                                     "aln_esp_cluster", "aln_esp_leaf", "supports_order", "cluster_leaf_ptr",
                                     "leaf_owner", "cluster_root_ptr", "root_leaf_ptr", "root_leaf", "sv_cluster",
                                     "sv_frag_ptr", "sv_frag", "sv_numchanged"]}
        arrays["cluster_kind"] = kind
        arrays["sub_frag_ptr"] = np.array([0, F], np.int32)
        arrays["sub_cluster_ptr"] = np.array([0, Cn], np.int32)
        arrays["sub_sv_ptr"] = np.array([0, len(arrays["sv_cluster"])], np.int32)
        nleaf = np.diff(arrays["cluster_leaf_ptr"])
        arrays["cluster_hist_ptr"] = np.concatenate([[0], np.cumsum(nleaf + 1)]).astype(np.int32)
        arrays["exp_pseq"] = gf("exp_pseq", E)
        arrays["exp_lambda"] = gf("exp_lambda", E)
        arrays["logprobcov"] = gf("logprobcov", E * (maxsup + 1))
        scalars = dict(abi_version=1, n_sub=1, n_frag=F, n_aln=A, n_aln_esp=K, n_cluster=Cn, n_exp=E,
                       max_support=maxsup, n_sv=len(arrays["sv_cluster"]), n_sv_frag=len(arrays["sv_frag"]),
                       n_leaf=len(arrays["leaf_owner"]), n_root=len(arrays["root_leaf_ptr"]) - 1,
                       n_root_leaf=len(arrays["root_leaf"]), n_hist=int(arrays["cluster_hist_ptr"][-1]))
        names = dict(frag_id=gs("frag_id", F), cluster_id=gs("cluster_id", Cn), exp_label=gs("exp_label", E),
                     aln_esp_name=gs("aln_esp_name", K), esp_names=gs("esp_names", P),
                     sorted_frag=gi("sorted_frag"), aln_esp_rawcluster=gi("aln_esp_rawcluster"),
                     aln_esp_id=gi("aln_esp_id"), leaf_esp_id=gi("leaf_esp_id"), hashmap_caps=gi("hashmap_caps"),
                     unsupported=L.orc_str_get(h, b"unsupported", 0).decode(),
                     warning=L.orc_str_get(h, b"warning", 0).decode(), tree_bin=bool(tree))
        return Problem(scalars, arrays, names)
    finally:
        L.orc_free(h)


class LoadedModel:
    """Keeps the parsed reference object graph alive for literal cross-checks."""

    def __init__(self, clusterfile, assignmentfile, experimentfile, clusters_first=False):
        L = lib()
        err = C.create_string_buffer(512)
        self.h = L.orc_load(clusterfile.encode(), assignmentfile.encode(), experimentfile.encode(),
                            int(clusters_first), err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())

    def literal_supports(self, c: int, assign: np.ndarray, n_exp: int, stride: int = 64):
        out = np.zeros(n_exp * stride, np.int32)
        a = np.ascontiguousarray(assign, np.int32)
        rc = lib().orc_literal_supports(self.h, c, _p(a, I32P), _p(out, I32P), stride)
        if rc:
            raise RuntimeError("reference Error in setBreakpoints")
        return [list(out[x * stride + 1: x * stride + 1 + out[x * stride]]) for x in range(n_exp)]

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_free(self.h)
            self.h = None


@dataclass
class RunResult:
    aln_count: np.ndarray
    frag_err_count: np.ndarray
    cluster_hist: np.ndarray
    final_assign: np.ndarray      # [n_chains, n_frag]
    final_logprob: np.ndarray     # [n_sub, n_chains]
    counters: np.ndarray          # [n_sub, n_chains, 5] naive_prop, naive_acc, sv_prop, sv_acc, reassignments
    rng_state: np.ndarray         # [n_sub, n_chains]
    error: np.ndarray
    trace_logprob: np.ndarray | None = None
    trace_move_frag: np.ndarray | None = None
    trace_move_assign: np.ndarray | None = None
    initial_assign: np.ndarray | None = None    # [n_chains, n_frag]
    initial_logprob: np.ndarray | None = None   # [n_sub, n_chains]
    totals: np.ndarray | None = None            # [6]
    kernel_ms: float = 0.0
    rc: int = 0


def reference_burnin(num_iters: int) -> int:
    """MultiBreakSV.java:186-193."""
    b = num_iters // 10
    return b if b != 0 else 10


def alloc_results(prob: Problem, n_chains: int, num_iters: int, trace: bool, results_type=OrcResults):
    s = prob.scalars
    S = s["n_sub"]
    r = RunResult(
        aln_count=np.zeros(s["n_aln"], np.uint32), frag_err_count=np.zeros(s["n_frag"], np.uint32),
        cluster_hist=np.zeros(s["n_hist"], np.uint32), final_assign=np.zeros((n_chains, s["n_frag"]), np.int32),
        final_logprob=np.zeros((S, n_chains), np.float64), counters=np.zeros((S, n_chains, 5), np.int64),
        rng_state=np.zeros((S, n_chains), np.uint64), error=np.zeros((S, n_chains), np.int32),
        initial_assign=np.zeros((n_chains, s["n_frag"]), np.int32), initial_logprob=np.zeros((S, n_chains), np.float64),
        totals=np.zeros(6, np.int64))
    if trace:
        r.trace_logprob = np.zeros((S, n_chains, num_iters), np.float64)
        r.trace_move_frag = np.zeros((S, n_chains, num_iters), np.int32)
        r.trace_move_assign = np.zeros((S, n_chains, num_iters), np.int32)
    cr = results_type()
    ft = dict(results_type._fields_)
    for k in ["aln_count", "frag_err_count", "cluster_hist", "final_assign", "final_logprob", "counters", "rng_state",
              "error", "trace_logprob", "trace_move_frag", "trace_move_assign", "initial_assign", "initial_logprob", "totals"]:
        a = getattr(r, k)
        setattr(cr, k, a.ctypes.data_as(ft[k]) if a is not None else ft[k]())
    return r, cr


def run(prob: Problem, mode: str = "faithful", num_iters: int = 10000, n_chains: int = 1, perr: float = 0.001,
        seed: int = 123456, burnin: int | None = None, java_int_wrap: bool = True, init_assign=None,
        trace: bool = False, nthreads: int = 1, memo_states: int = 65536) -> RunResult:
    """MultiBreakSV.runMCMC (MultiBreakSV.java:773-877) on every (subproblem, chain) of `prob`, on the CPU."""
    d = prob.desc(OrcDesc)
    p = OrcParams()
    p.mode = {"faithful": 0, "incremental": 1}[mode]
    p.n_chains = n_chains
    p.num_iters = num_iters
    p.burnin = reference_burnin(num_iters) if burnin is None else burnin
    p.perr = perr
    p.seed = seed
    p.java_int_wrap = int(java_int_wrap)
    ia = None
    if init_assign is not None:
        ia = np.ascontiguousarray(init_assign, np.int32)
        p.init_assign = ia.ctypes.data_as(I32P)
    p.trace = int(trace)
    p.memo_states = memo_states
    r, cr = alloc_results(prob, n_chains, num_iters, trace)
    r.rc = lib().orc_run(C.byref(d), C.byref(p), C.byref(cr), nthreads)
    return r


def enumerate_logprobs(prob: Problem, sub: int = 0, perr: float = 0.001, java_int_wrap: bool = True,
                       cap: int = 2_000_000) -> np.ndarray:
    d = prob.desc(OrcDesc)
    out = np.zeros(cap, np.float64)
    n = lib().orc_enumerate(C.byref(d), sub, perr, int(java_int_wrap), _p(out, F64P), cap)
    if n < 0:
        raise ValueError("state space larger than cap")
    return out[:n]


def replay_states(prob: Problem, init_assign: np.ndarray, move_frag: np.ndarray, move_assign: np.ndarray,
                  sub: int = 0) -> np.ndarray:
    """Rebuild the per-iteration assignment vectors of one chain from its accepted-move log."""
    a = prob.arrays
    f0, f1 = int(a["sub_frag_ptr"][sub]), int(a["sub_frag_ptr"][sub + 1])
    nal = np.diff(a["frag_aln_ptr"])
    cur = np.array(init_assign[f0:f1], np.int32)
    out = np.zeros((len(move_frag), f1 - f0), np.int32)
    sv0 = int(a["sub_sv_ptr"][sub])
    for i, (mf, ma) in enumerate(zip(move_frag, move_assign)):
        if mf >= 0:
            cur[mf - f0] = ma
        elif mf <= -2:
            k = sv0 + (-mf - 2)
            for q in range(a["sv_frag_ptr"][k], a["sv_frag_ptr"][k + 1]):
                f = a["sv_frag"][q] - f0
                assert nal[f + f0] == 1
                cur[f] = 0 if cur[f] == -1 else -1
        out[i] = cur
    return out
