"""ctypes binding of libmbsv_b200.so (include/mbsv_b200.h).

The library is the product: if it is missing, or there is no CUDA device, every compute call fails loudly —
there is no CPU fallback and nothing here ever touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MBSV_B200_LIB selects an alternative build of the same library (kernel-tuning experiments); it is still a
# CUDA build of this repository's sources, never a fallback.
LIB_PATH = os.environ.get("MBSV_B200_LIB") or os.path.join(_HERE, "libmbsv_b200.so")

I32P = C.POINTER(C.c_int32)
U8P = C.POINTER(C.c_uint8)
U32P = C.POINTER(C.c_uint32)
I64P = C.POINTER(C.c_int64)
U64P = C.POINTER(C.c_uint64)
F64P = C.POINTER(C.c_double)

MBSV_ABI_VERSION = 1
MODE_REPLAY_EXACT = 0
MODE_FAST = 1
MODE_GIBBS = 2          # heat-bath sampler of the statistical tier (not the reference's chain)

STATUS = {0: "MBSV_OK", -1: "MBSV_ERR_ARG", -2: "MBSV_ERR_CUDA", -3: "MBSV_ERR_NCCL", -4: "MBSV_ERR_OOM",
          -5: "MBSV_ERR_INVARIANT", -6: "MBSV_ERR_UNSUPPORTED", -7: "MBSV_ERR_IO"}

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


class ProblemDesc(C.Structure):
    """mbsv_problem_desc"""
    _fields_ = [(n, C.c_int32) for n in DESC_SCALARS] + [(n, _CT[t]) for n, t in DESC_ARRAYS]


class RunParams(C.Structure):
    """mbsv_run_params"""
    _fields_ = [("mode", C.c_int32), ("n_chains", C.c_int32), ("num_iters", C.c_int32), ("burnin", C.c_int32),
                ("perr", C.c_double), ("seed", C.c_int64), ("device", C.c_int32), ("java_int_wrap", C.c_int32),
                ("init_assign", I32P), ("trace", C.c_int32), ("memo_states", C.c_int32)]


class Results(C.Structure):
    """mbsv_results"""
    _fields_ = [("aln_count", U32P), ("frag_err_count", U32P), ("cluster_hist", U32P), ("final_assign", I32P),
                ("final_logprob", F64P), ("counters", I64P), ("rng_state", U64P), ("error", I32P),
                ("trace_logprob", F64P), ("trace_move_frag", I32P), ("trace_move_assign", I32P),
                ("initial_assign", I32P), ("initial_logprob", F64P), ("totals", I64P), ("kernel_ms", C.c_double)]


EXPORTS = ["mbsv_abi_version", "mbsv_device_count", "mbsv_last_error", "mbsv_status_string", "mbsv_problem_create",
           "mbsv_problem_destroy", "mbsv_run", "mbsv_run_device", "mbsv_problem_stats", "mbsv_last_launch_count",
           "mbsv_log", "mbsv_exp", "mbsv_fill_logprobcov", "mbsv_lpt_shard", "mbsv_last_launch_info",
           "mbsv_last_launch_totals", "mbsv_split_chains", "mbsv_comm_unique_id", "mbsv_comm_create", "mbsv_comm_destroy", "mbsv_reduce_tallies",
           "mbsv_run_multi", "mbsv_trim"]
COMM_ID_BYTES = 128

_lib = None


class MbsvError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


def load_library():
    """Load libmbsv_b200.so; raises (never falls back) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C multibreak_sv_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.mbsv_abi_version.restype = C.c_int
    L.mbsv_device_count.restype = C.c_int
    L.mbsv_last_error.restype = C.c_char_p
    L.mbsv_status_string.argtypes = [C.c_int]; L.mbsv_status_string.restype = C.c_char_p
    L.mbsv_problem_create.argtypes = [C.POINTER(ProblemDesc), C.c_int32, C.POINTER(C.c_void_p)]
    L.mbsv_problem_create.restype = C.c_int
    L.mbsv_problem_destroy.argtypes = [C.c_void_p]; L.mbsv_problem_destroy.restype = C.c_int
    L.mbsv_trim.argtypes = [C.c_int32]; L.mbsv_trim.restype = C.c_int
    L.mbsv_run.argtypes = [C.c_void_p, C.POINTER(RunParams), C.POINTER(Results)]; L.mbsv_run.restype = C.c_int
    L.mbsv_run_device.argtypes = [C.c_void_p, C.POINTER(RunParams), C.POINTER(Results)]
    L.mbsv_run_device.restype = C.c_int
    L.mbsv_problem_stats.argtypes = [C.c_void_p, I64P, C.c_int32]; L.mbsv_problem_stats.restype = C.c_int
    L.mbsv_last_launch_count.restype = C.c_int64
    L.mbsv_log.argtypes = [C.c_double]; L.mbsv_log.restype = C.c_double
    L.mbsv_exp.argtypes = [C.c_double]; L.mbsv_exp.restype = C.c_double
    L.mbsv_fill_logprobcov.argtypes = [C.c_double, C.c_int32, F64P]; L.mbsv_fill_logprobcov.restype = C.c_int
    L.mbsv_last_launch_info.argtypes = [F64P, C.c_int32]; L.mbsv_last_launch_info.restype = C.c_int
    L.mbsv_lpt_shard.argtypes = [I64P, C.c_int64, C.c_int32, I32P]; L.mbsv_lpt_shard.restype = C.c_int
    L.mbsv_last_launch_totals.argtypes = [I64P, C.c_int32]; L.mbsv_last_launch_totals.restype = C.c_int
    L.mbsv_split_chains.argtypes = [C.c_int32, C.c_int32, C.c_int32, I32P, I32P]; L.mbsv_split_chains.restype = C.c_int
    L.mbsv_comm_unique_id.argtypes = [U8P]; L.mbsv_comm_unique_id.restype = C.c_int
    L.mbsv_comm_create.argtypes = [U8P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    L.mbsv_comm_create.restype = C.c_int
    L.mbsv_comm_destroy.argtypes = [C.c_void_p]; L.mbsv_comm_destroy.restype = C.c_int
    L.mbsv_reduce_tallies.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, F64P]
    L.mbsv_reduce_tallies.restype = C.c_int
    L.mbsv_run_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(RunParams), C.POINTER(Results), F64P]
    L.mbsv_run_multi.restype = C.c_int
    if L.mbsv_abi_version() != MBSV_ABI_VERSION:
        raise ImportError("libmbsv_b200.so ABI version mismatch")
    _lib = L
    return L


def check(status: int):
    if status != 0:
        raise MbsvError(status, load_library().mbsv_last_error().decode())


@dataclass
class PackedProblem:
    """Host copy of a mbsv_problem_desc: the flat CSR of one or more independent subproblems."""
    scalars: dict
    arrays: dict

    def desc(self) -> ProblemDesc:
        d = ProblemDesc()
        for k in DESC_SCALARS:
            setattr(d, k, int(self.scalars[k]))
        self._keep = []
        ft = dict(ProblemDesc._fields_)
        for k, t in DESC_ARRAYS:
            a = np.ascontiguousarray(self.arrays[k], dtype=t)
            self._keep.append(a)
            setattr(d, k, a.ctypes.data_as(ft[k]) if a.size else ft[k]())
        return d

    @property
    def host_bytes(self) -> int:
        return int(sum(np.asarray(self.arrays[k]).nbytes for k, _ in DESC_ARRAYS))


@dataclass
class RunOutput:
    aln_count: np.ndarray
    frag_err_count: np.ndarray
    cluster_hist: np.ndarray
    final_assign: np.ndarray
    final_logprob: np.ndarray
    counters: np.ndarray
    rng_state: np.ndarray
    error: np.ndarray
    initial_assign: np.ndarray
    initial_logprob: np.ndarray
    trace_logprob: np.ndarray | None
    trace_move_frag: np.ndarray | None
    trace_move_assign: np.ndarray | None
    kernel_ms: float
    launches: int
    totals: np.ndarray | None = None     # [6] proposals/moves/reassignments summed over chains, first chain status

    @property
    def result_bytes(self) -> int:
        return int(sum(a.nbytes for a in (self.aln_count, self.frag_err_count, self.cluster_hist, self.final_assign,
                                          self.final_logprob, self.counters, self.rng_state, self.error, self.totals)
                       if a is not None))


def trim(device: int = -1) -> None:
    """mbsv_trim: return the device memory the library keeps for reuse by the next problem (device < 0: every device)."""
    check(load_library().mbsv_trim(device))


def reference_burnin(num_iters: int) -> int:
    """burnin = numiters/10, 0 -> 10 (MultiBreakSV.java:186-193)."""
    b = num_iters // 10
    return b if b != 0 else 10


class DeviceProblem:
    """A problem resident on one GPU (mbsv_problem handle)."""

    def __init__(self, packed: PackedProblem, device: int = 0):
        self.lib = load_library()
        self.packed = packed
        self.device = device
        self.handle = C.c_void_p()
        d = packed.desc()
        check(self.lib.mbsv_problem_create(C.byref(d), device, C.byref(self.handle)))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.mbsv_problem_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, mode: int = MODE_FAST, num_iters: int = 10000, n_chains: int = 1, perr: float = 0.001,
            seed: int = 123456, burnin: int | None = None, java_int_wrap: bool | None = None, init_assign=None,
            trace: bool = False, want_state: bool = True, per_chain: bool = True, memo_states: int = 65536,
            allow_invariant: bool = False) -> RunOutput:
        """MultiBreakSV.runMCMC (MultiBreakSV.java:773-877) for every (subproblem, chain), on the GPU.

        A chain that hits an invariant the reference turns into an Error (MBSV_ERR_INVARIANT: its tallies are not
        balanced and must not be used) raises, like every other failure; allow_invariant=True returns the outputs with
        the per-chain `error` array instead (tests of that path)."""
        s = self.packed.scalars
        S, F = s["n_sub"], s["n_frag"]
        rp = RunParams()
        rp.mode = mode
        rp.n_chains = n_chains
        rp.num_iters = num_iters
        rp.burnin = reference_burnin(num_iters) if burnin is None else burnin
        rp.perr = perr
        rp.seed = seed
        rp.device = self.device
        rp.java_int_wrap = int(mode == MODE_REPLAY_EXACT if java_int_wrap is None else java_int_wrap)
        ia = None
        if init_assign is not None:
            ia = np.ascontiguousarray(init_assign, np.int32)
            rp.init_assign = ia.ctypes.data_as(I32P)
        rp.trace = int(trace)
        rp.memo_states = memo_states
        out = RunOutput(
            aln_count=np.zeros(s["n_aln"], np.uint32), frag_err_count=np.zeros(F, np.uint32),
            cluster_hist=np.zeros(s["n_hist"], np.uint32),
            final_assign=np.zeros((n_chains, F) if want_state else (0, 0), np.int32),
            final_logprob=np.zeros((S, n_chains) if per_chain else (0, 0), np.float64),
            counters=np.zeros((S, n_chains, 5) if per_chain else (0, 0, 0), np.int64),
            rng_state=np.zeros((S, n_chains) if per_chain else (0, 0), np.uint64),
            error=np.zeros((S, n_chains) if per_chain else (0, 0), np.int32),
            initial_assign=np.zeros((n_chains, F) if want_state else (0, 0), np.int32),
            initial_logprob=np.zeros((S, n_chains) if per_chain else (0, 0), np.float64),
            totals=np.zeros(6, np.int64),
            trace_logprob=np.zeros((S, n_chains, num_iters), np.float64) if trace else None,
            trace_move_frag=np.zeros((S, n_chains, num_iters), np.int32) if trace else None,
            trace_move_assign=np.zeros((S, n_chains, num_iters), np.int32) if trace else None,
            kernel_ms=0.0, launches=0)
        r = Results()
        ft = dict(Results._fields_)
        for k in ["aln_count", "frag_err_count", "cluster_hist", "final_assign", "final_logprob", "counters",
                  "rng_state", "error", "trace_logprob", "trace_move_frag", "trace_move_assign", "initial_assign",
                  "initial_logprob", "totals"]:
            a = getattr(out, k)
            setattr(r, k, a.ctypes.data_as(ft[k]) if a is not None and a.size else ft[k]())
        st = self.lib.mbsv_run(self.handle, C.byref(rp), C.byref(r))
        out.kernel_ms = r.kernel_ms
        out.launches = int(self.lib.mbsv_last_launch_count())
        self.last_status = st
        if st != 0 and not (st == -5 and allow_invariant):
            check(st)
        return out

    def launch_info(self) -> list[dict]:
        buf = np.zeros(20 * 5, np.float64)
        n = self.lib.mbsv_last_launch_info(buf.ctypes.data_as(F64P), 20)
        tot = np.zeros(20 * 6, np.int64)
        self.lib.mbsv_last_launch_totals(tot.ctypes.data_as(I64P), 20)
        return [dict(rows=int(buf[i * 5]), warps=int(buf[i * 5 + 1]), grid=int(buf[i * 5 + 2]),
                     smem_bytes=int(buf[i * 5 + 3]), ms=float(buf[i * 5 + 4]), proposals=int(tot[i * 6] + tot[i * 6 + 2]),
                     accepted=int(tot[i * 6 + 1] + tot[i * 6 + 3]), reassignments=int(tot[i * 6 + 4])) for i in range(n)]

    def run_device(self, tensors: dict, mode: int = MODE_FAST, num_iters: int = 10000, n_chains: int = 1,
                   perr: float = 0.001, seed: int = 123456, burnin: int | None = None, java_int_wrap: bool = False,
                   memo_states: int = 65536, allow_invariant: bool = False):
        """mbsv_run_device: `tensors` maps result-field names to DEVICE pointers (e.g. torch tensor.data_ptr()) on
        this problem's GPU; aln_count, frag_err_count, cluster_hist and totals are required.  Returns
        (status, kernel_ms)."""
        rp = RunParams()
        rp.mode = mode
        rp.n_chains = n_chains
        rp.num_iters = num_iters
        rp.burnin = reference_burnin(num_iters) if burnin is None else burnin
        rp.perr = perr
        rp.seed = seed
        rp.device = self.device
        rp.java_int_wrap = int(java_int_wrap)
        rp.memo_states = memo_states
        r = Results()
        ft = dict(Results._fields_)
        for k, ptr in tensors.items():
            setattr(r, k, C.cast(C.c_void_p(int(ptr)), ft[k]))
        st = self.lib.mbsv_run_device(self.handle, C.byref(rp), C.byref(r))
        if st != 0 and not (st == -5 and allow_invariant):
            check(st)
        return st, r.kernel_ms


def _host_nccl_first():
    """The library takes the NCCL the process already has (same SONAME); a Python host that also uses PyTorch must have
    torch's bundled NCCL loaded BEFORE the library opens the system one, or `import torch` fails later on missing symbols."""
    import importlib.util
    if importlib.util.find_spec("torch") is not None:
        import torch  # noqa: F401


class Comm:
    """mbsv_comm: the NCCL communicator of the giant-component path, one process per GPU.  `exchange` hands rank 0's
    128-byte id to every rank (e.g. a torch.distributed broadcast of a uint8 tensor, or a file)."""

    def __init__(self, n_ranks: int, rank: int, device: int, exchange):
        _host_nccl_first()
        self.lib = load_library()
        ident = np.zeros(COMM_ID_BYTES, np.uint8)
        if rank == 0:
            check(self.lib.mbsv_comm_unique_id(ident.ctypes.data_as(U8P)))
        ident = np.ascontiguousarray(exchange(ident), np.uint8)
        self.handle = C.c_void_p()
        check(self.lib.mbsv_comm_create(ident.ctypes.data_as(U8P), n_ranks, rank, device, C.byref(self.handle)))

    def reduce_tallies(self, dev_ptr: int, n_words: int, root: int = 0) -> float:
        """One ncclReduce(sum, u32) of the contiguous device tallies onto `root`, in place; returns its device time (ms)."""
        ms = C.c_double(0.0)
        check(self.lib.mbsv_reduce_tallies(self.handle, C.c_void_p(int(dev_ptr)), int(n_words), root, C.byref(ms)))
        return ms.value

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mbsv_comm_destroy(self.handle)
            self.handle = C.c_void_p()


def run_multi(problems, mode: int = MODE_FAST, num_iters: int = 10000, n_chains: int = 8, perr: float = 0.001,
              seed: int = 123456, burnin: int | None = None, java_int_wrap: bool = False, memo_states: int = 65536):
    """mbsv_run_multi: `problems` = the same PackedProblem created on several GPUs of this process (DeviceProblem
    objects); chains split over them, tallies pooled with one ncclReduce.  Returns (RunOutput with the pooled tallies and
    totals, reduce_ms)."""
    _host_nccl_first()
    L = load_library()
    s = problems[0].packed.scalars
    rp = RunParams()
    rp.mode, rp.n_chains, rp.num_iters = mode, n_chains, num_iters
    rp.burnin = reference_burnin(num_iters) if burnin is None else burnin
    rp.perr, rp.seed, rp.device = perr, seed, problems[0].device
    rp.java_int_wrap, rp.memo_states = int(java_int_wrap), memo_states
    out = RunOutput(aln_count=np.zeros(s["n_aln"], np.uint32), frag_err_count=np.zeros(s["n_frag"], np.uint32),
                    cluster_hist=np.zeros(s["n_hist"], np.uint32), final_assign=np.zeros((0, 0), np.int32),
                    final_logprob=np.zeros((0, 0)), counters=np.zeros((0, 0, 0), np.int64),
                    rng_state=np.zeros((0, 0), np.uint64), error=np.zeros((0, 0), np.int32),
                    initial_assign=np.zeros((0, 0), np.int32), initial_logprob=np.zeros((0, 0)), trace_logprob=None,
                    trace_move_frag=None, trace_move_assign=None, kernel_ms=0.0, launches=0, totals=np.zeros(6, np.int64))
    r = Results()
    ft = dict(Results._fields_)
    for k in ("aln_count", "frag_err_count", "cluster_hist", "totals"):
        a = getattr(out, k)
        setattr(r, k, a.ctypes.data_as(ft[k]) if a.size else ft[k]())
    handles = (C.c_void_p * len(problems))(*[p.handle for p in problems])
    ms = C.c_double(0.0)
    check(L.mbsv_run_multi(handles, len(problems), C.byref(rp), C.byref(r), C.byref(ms)))
    out.kernel_ms = r.kernel_ms
    return out, ms.value
