"""Host-side mirror of the reference driver class `MultiBreakSV` (src/MultiBreakSV.java), backed by the C++ mirror
in libmbsv_b200.so (include/mbsv_host.h).  Method names, argument meaning and failure behaviour follow the reference:

    sv = MultiBreakSV(["--clusterfile", c, "--assignmentfile", a, "--experimentfile", e,
                       "--numiterations", "10000", "--perr", "0.001", "--prefix", "out"])
    sv.getFragmentAlignments(); sv.constructClusterDiagrams(); sv.precomputePriors(); sv.fillNonSingletonArray()
    sv.runMCMC(); sv.writeFinalFiles(); sv.writeStateCounts()          # == MultiBreakSV.main (:114-153)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import (DESC_ARRAYS, DESC_SCALARS, F64P, I32P, I64P, MODE_REPLAY_EXACT, U32P, PackedProblem, ProblemDesc, Results, check,
                   load_library, reference_burnin)


def _bind(L):
    if getattr(L, "_host_bound", False):
        return L
    vp = C.c_void_p
    L.mbsv_host_create.argtypes = [C.POINTER(vp)]
    L.mbsv_host_destroy.argtypes = [vp]
    L.mbsv_host_set_files.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32]
    for n in ("mbsv_host_get_fragment_alignments", "mbsv_host_construct_cluster_diagrams", "mbsv_host_precompute_priors",
              "mbsv_host_fill_non_singleton_array", "mbsv_host_load"):
        getattr(L, n).argtypes = [vp]
    L.mbsv_host_desc.argtypes = [vp, C.POINTER(ProblemDesc)]
    L.mbsv_host_name.argtypes = [vp, C.c_int32, C.c_int32]; L.mbsv_host_name.restype = C.c_char_p
    L.mbsv_host_sorted_fragments.argtypes = [vp, I32P]
    L.mbsv_host_run_mcmc.argtypes = [vp, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_int32, I32P, C.POINTER(Results)]
    L.mbsv_host_run_summary.argtypes = [vp, I64P, F64P, F64P]
    L.mbsv_host_matchfile_initial.argtypes = [vp, C.c_char_p, I32P, I32P]
    L.mbsv_host_set_results.argtypes = [vp, C.c_int32, C.c_int32, U32P, U32P, U32P]
    L.mbsv_host_write_final_files.argtypes = [vp, C.c_char_p]
    L.mbsv_host_write_mcmc_file.argtypes = [vp, C.c_char_p]
    L.mbsv_host_write_state_counts.argtypes = [vp, C.c_char_p]
    L.mbsv_host_batch_create.argtypes = [C.POINTER(vp), C.c_int32, C.POINTER(vp)]
    L.mbsv_host_batch_desc.argtypes = [vp, C.POINTER(ProblemDesc)]
    L.mbsv_host_batch_destroy.argtypes = [vp]
    L.mbsv_host_load_partitioned.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(vp), I32P]
    L.mbsv_host_batch_write_final_files.argtypes = [vp, C.c_int32, U32P, U32P, U32P, C.c_char_p]
    L.mbsv_host_batch_numbers.argtypes = [vp, I32P]
    L.mbsv_host_partition.argtypes = [C.c_char_p, C.c_char_p, I32P, C.c_int32]
    L.mbsv_java_double_to_string.argtypes = [C.c_double, C.c_char_p, C.c_int32]
    L.mbsv_java_hashmap_order.argtypes = [C.c_char_p, C.c_int32, I32P, I32P]
    L._host_bound = True
    return L


def java_double_to_string(v: float) -> str:
    buf = C.create_string_buffer(64)
    _bind(load_library()).mbsv_java_double_to_string(float(v), buf, 64)
    return buf.value.decode()


def java_hashmap_order(keys):
    L = _bind(load_library())
    blob = b"".join(k.encode() + b"\0" for k in keys)
    out = np.zeros(max(1, len(keys)), np.int32)
    cap = C.c_int32(0)
    n = L.mbsv_java_hashmap_order(blob, len(keys), out.ctypes.data_as(I32P), C.byref(cap))
    return [keys[i] for i in out[:n]], cap.value


def generateIndependentSubproblems(clusterfile: str, outdir: str | None = None):
    """src/generateIndependentSubproblems.pl: returns (number of subsets, subset id of every physical line — 0 for
    the dotted maximal-cluster rows the script drops) and writes <outdir>/subset-N.clusters when outdir is given."""
    L = _bind(load_library())
    nlines = sum(1 for _ in open(clusterfile))
    rows = np.zeros(max(1, nlines), np.int32)
    n = L.mbsv_host_partition(clusterfile.encode(), outdir.encode() if outdir else None, rows.ctypes.data_as(I32P), nlines)
    check(min(n, 0))
    return n, rows[:nlines]


def _desc_to_packed(d: ProblemDesc) -> PackedProblem:
    sc = {k: int(getattr(d, k)) for k in DESC_SCALARS}
    n = dict(sub_frag_ptr=sc["n_sub"] + 1, sub_cluster_ptr=sc["n_sub"] + 1, sub_sv_ptr=sc["n_sub"] + 1,
             frag_aln_ptr=sc["n_frag"] + 1, aln_numerrors=sc["n_aln"], aln_len=sc["n_aln"], aln_exp=sc["n_aln"],
             aln_esp_ptr=sc["n_aln"] + 1, aln_esp_cluster=sc["n_aln_esp"], aln_esp_leaf=sc["n_aln_esp"],
             cluster_kind=sc["n_cluster"], supports_order=sc["n_cluster"], cluster_hist_ptr=sc["n_cluster"] + 1,
             cluster_leaf_ptr=sc["n_cluster"] + 1, leaf_owner=sc["n_leaf"], cluster_root_ptr=sc["n_cluster"] + 1,
             root_leaf_ptr=sc["n_root"] + 1, root_leaf=sc["n_root_leaf"], sv_cluster=sc["n_sv"], sv_frag_ptr=sc["n_sv"] + 1,
             sv_frag=sc["n_sv_frag"], sv_numchanged=sc["n_sv"], exp_pseq=sc["n_exp"], exp_lambda=sc["n_exp"],
             logprobcov=sc["n_exp"] * (sc["max_support"] + 1))
    arrays = {}
    for k, t in DESC_ARRAYS:
        p = getattr(d, k)
        arrays[k] = np.ctypeslib.as_array(p, shape=(n[k],)).copy() if n[k] and p else np.zeros(0, t)
    return PackedProblem(sc, arrays)


class MultiBreakSV:
    """Mirror of the reference's public driver interface for the sampler path."""

    def __init__(self, args=None, **kw):
        self.lib = _bind(load_library())
        self.h = C.c_void_p()
        check(self.lib.mbsv_host_create(C.byref(self.h)))
        self.clusterfile = self.assignmentfile = self.experimentfile = self.matchfile = None
        self.outputname = "out"
        self.numiters, self.perr = -1, -1.0
        self.SAVE_SPACE = self.CLUSTERS_FIRST = False
        self.mode, self.device = MODE_REPLAY_EXACT, 0
        if args is not None:
            self.parseOptions(list(args))
        for k, v in kw.items():
            setattr(self, k, v)
        if self.clusterfile is not None:
            self._set_files()

    # MultiBreakSV.java:211-267
    def parseOptions(self, args):
        i = 0
        try:
            while i < len(args):
                a = args[i]
                if a == "--clusterfile": self.clusterfile = args[i + 1]; i += 1
                elif a == "--assignmentfile": self.assignmentfile = args[i + 1]; i += 1
                elif a == "--experimentfile": self.experimentfile = args[i + 1]; i += 1
                elif a == "--matchfile": self.matchfile = args[i + 1]; i += 1
                elif a == "--prefix": self.outputname = args[i + 1]; i += 1
                elif a == "--numiterations": self.numiters = int(args[i + 1]); i += 1
                elif a == "--perr": self.perr = float(args[i + 1]); i += 1
                elif a == "--savespace": self.SAVE_SPACE = True
                elif a == "--readclustersfirst": self.CLUSTERS_FIRST = True
                elif a == "--enumerate": raise ValueError("--enumerate is outside the sampler path")
                else: raise ValueError(f"{a} is not a valid option.")
                i += 1
        except (IndexError, TypeError):
            raise ValueError("--numiterations must be a positive integer, --perr must be a float between 0 and 1.")
        if self.clusterfile is None: raise ValueError("Cluster file is required.")
        if self.assignmentfile is None: raise ValueError("Assignment file is required.")
        if self.experimentfile is None: raise ValueError("Experiment file is required.")
        if self.numiters == -1: raise ValueError("Number of iterations is required to sample.")
        if self.perr == -1: raise ValueError("Mapping error probability (perr) is required.")
        if self.numiters < 0: raise ValueError("--numiterations must be a positive integer.")
        if self.perr < 0 or self.perr > 1: raise ValueError("--perr must be a value between 0 and 1.")

    def _set_files(self):
        check(self.lib.mbsv_host_set_files(self.h, self.clusterfile.encode(), self.assignmentfile.encode(),
                                           self.experimentfile.encode(), int(self.CLUSTERS_FIRST)))

    @property
    def burnin(self):
        return reference_burnin(self.numiters)

    def getFragmentAlignments(self, verbose=False):
        self._set_files()
        check(self.lib.mbsv_host_get_fragment_alignments(self.h))

    def constructClusterDiagrams(self, verbose=False):
        check(self.lib.mbsv_host_construct_cluster_diagrams(self.h))

    def precomputePriors(self):
        check(self.lib.mbsv_host_precompute_priors(self.h))

    def fillNonSingletonArray(self):
        check(self.lib.mbsv_host_fill_non_singleton_array(self.h))

    def load(self):
        """getFragmentAlignments .. fillNonSingletonArray in main()'s order."""
        self._set_files()
        check(self.lib.mbsv_host_load(self.h))
        return self

    def packed(self) -> PackedProblem:
        d = ProblemDesc()
        check(self.lib.mbsv_host_desc(self.h, C.byref(d)))
        return _desc_to_packed(d)

    def names(self, kind: int, n: int):
        return [self.lib.mbsv_host_name(self.h, kind, i).decode() for i in range(n)]

    def warning(self) -> str:
        return self.lib.mbsv_host_name(self.h, 4, 0).decode()

    def rejected(self) -> str:
        return self.lib.mbsv_host_name(self.h, 5, 0).decode()

    def sortedFragments(self, n: int) -> np.ndarray:
        out = np.zeros(n, np.int32)
        self.lib.mbsv_host_sorted_fragments(self.h, out.ctypes.data_as(I32P))
        return out

    def initialAssignmentFromMatchfile(self, p: PackedProblem = None):
        """The --matchfile path (MultiBreakSV.java:1030-1085) reduced to "initial state supplied by the host": an
        alignment is set when it has exactly one ESP and that ESP is listed (the loop `break`s after the first hit,
        so thisval <= 1, :1058-1065); every other fragment starts unassigned."""
        d = ProblemDesc()
        check(self.lib.mbsv_host_desc(self.h, C.byref(d)))
        init = np.full(max(1, d.n_frag), -1, np.int32)
        counts = np.zeros(4, np.int32)
        check(self.lib.mbsv_host_matchfile_initial(self.h, self.matchfile.encode(), init.ctypes.data_as(I32P),
                                                   counts.ctypes.data_as(I32P)))
        self.matchfile_counts = dict(zip(("numset", "numerr", "numbad", "esps"), (int(v) for v in counts)))
        return init[:d.n_frag]

    def runMCMC(self, file=None):
        """MultiBreakSV.runMCMC (:773-877) on the GPU; results stay in the host object for the writers."""
        init = None
        if self.matchfile is not None:
            init = self.initialAssignmentFromMatchfile(self.packed())
        ip = init.ctypes.data_as(I32P) if init is not None else I32P()
        check(self.lib.mbsv_host_run_mcmc(self.h, self.numiters, self.perr, self.mode, self.device, int(self.SAVE_SPACE), ip, None))
        if not self.SAVE_SPACE:
            check(self.lib.mbsv_host_write_mcmc_file(self.h, (file or self.outputname + ".mcmc.txt").encode()))

    def setResults(self, aln_count, frag_err_count, cluster_hist, numiters=None):
        if numiters is not None:
            self.numiters = numiters
        a, f, h = (np.ascontiguousarray(x, np.uint32) for x in (aln_count, frag_err_count, cluster_hist))
        check(self.lib.mbsv_host_set_results(self.h, self.numiters, self.burnin, a.ctypes.data_as(U32P), f.ctypes.data_as(U32P),
                                             h.ctypes.data_as(U32P)))

    def writeFinalFiles(self):
        check(self.lib.mbsv_host_write_final_files(self.h, self.outputname.encode()))

    def writeStateCounts(self, keys=None):
        check(self.lib.mbsv_host_write_state_counts(self.h, (self.outputname + ".samplingprobs").encode()))

    def close(self):
        if self.h:
            self.lib.mbsv_host_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def batch(hosts) -> PackedProblem:
    """Concatenate loaded hosts (one per subset-N.clusters file) into one batched PackedProblem."""
    L = _bind(load_library())
    arr = (C.c_void_p * len(hosts))(*[h.h for h in hosts])
    b = C.c_void_p()
    check(L.mbsv_host_batch_create(arr, len(hosts), C.byref(b)))
    try:
        d = ProblemDesc()
        check(L.mbsv_host_batch_desc(b, C.byref(d)))
        return _desc_to_packed(d)
    finally:
        L.mbsv_host_batch_destroy(b)


def load_partitioned(clusterfile: str, assignmentfile: str, experimentfile: str, info=None):
    """Partition + load in memory (mbsv_host_load_partitioned): one subproblem per subset of
    generateIndependentSubproblems.pl, each packed as a run on `subset-k.clusters --readclustersfirst` would see it; the
    assignments file is read once.  Returns (PackedProblem, number of skipped empty subsets); `info` (a dict) also receives
    the number of subsets one of whose java.util.HashMaps has a treeified bin (their order comes from the literal
    red-black emulation, csrc/java_hashmap.hpp)."""
    L = _bind(load_library())
    b = C.c_void_p()
    skipped = np.zeros(2, np.int32)
    n = L.mbsv_host_load_partitioned(clusterfile.encode(), assignmentfile.encode(), experimentfile.encode(),
                                     0, C.byref(b), skipped.ctypes.data_as(I32P))
    if info is not None:
        info.update(skipped=int(skipped[0]), treeified=int(skipped[1]))
    if n < 0:
        check(n)
    try:
        d = ProblemDesc()
        check(L.mbsv_host_batch_desc(b, C.byref(d)))
        return _desc_to_packed(d), int(skipped[0])
    finally:
        L.mbsv_host_batch_destroy(b)


class PartitionedBatch:
    """Files -> partition -> one batched problem that remembers its names, so that the reference's output files can be
    written for the whole batch (mbsv_host_load_partitioned with names + mbsv_host_batch_write_final_files)."""

    def __init__(self, clusterfile: str, assignmentfile: str, experimentfile: str):
        self.lib = _bind(load_library())
        self.h = C.c_void_p()
        info = np.zeros(2, np.int32)
        n = self.lib.mbsv_host_load_partitioned(clusterfile.encode(), assignmentfile.encode(), experimentfile.encode(),
                                                2, C.byref(self.h), info.ctypes.data_as(I32P))
        if n < 0:
            check(n)
        self.skipped, self.treeified = int(info[0]), int(info[1])
        d = ProblemDesc()
        check(self.lib.mbsv_host_batch_desc(self.h, C.byref(d)))
        self.packed = _desc_to_packed(d)
        self.numbers = np.zeros(n, np.int32)                 # subset number of every subproblem
        self.lib.mbsv_host_batch_numbers(self.h, self.numbers.ctypes.data_as(I32P))

    def writeFinalFiles(self, prefix: str, burnin: int, aln_count, frag_err_count, cluster_hist):
        a, f, h = (np.ascontiguousarray(x, np.uint32) for x in (aln_count, frag_err_count, cluster_hist))
        check(self.lib.mbsv_host_batch_write_final_files(self.h, burnin, a.ctypes.data_as(U32P), f.ctypes.data_as(U32P),
                                                         h.ctypes.data_as(U32P), prefix.encode()))

    def close(self):
        if self.h:
            self.lib.mbsv_host_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def main_partition(argv):
    """`--partition [--chains N] [--fast]`: the manual's multi-subproblem workflow as one GPU batch (the same
    as `mbsv_cli --partition`): partition + load in memory, one run, concatenated `.final*.longburnin` files."""
    from ._abi import DeviceProblem, MODE_FAST
    argv = list(argv)
    chains, mode = 1, MODE_REPLAY_EXACT
    argv.remove("--partition")
    if "--fast" in argv:
        argv.remove("--fast"); mode = MODE_FAST
    if "--chains" in argv:
        i = argv.index("--chains"); chains = int(argv[i + 1]); del argv[i:i + 2]
    sv = MultiBreakSV(argv)                                   # option checks and burnin as in the single-problem path
    b = PartitionedBatch(sv.clusterfile, sv.assignmentfile, sv.experimentfile)
    dp = DeviceProblem(b.packed, device=sv.device)
    r = dp.run(mode, sv.numiters, n_chains=chains, perr=sv.perr, java_int_wrap=True, want_state=False, per_chain=False)
    b.writeFinalFiles(sv.outputname, sv.burnin * chains, r.aln_count, r.frag_err_count, r.cluster_hist)
    return b, r


def main(argv):
    """MultiBreakSV.main (:114-153)."""
    if "--partition" in argv:
        return main_partition(argv)
    sv = MultiBreakSV(argv)
    sv.getFragmentAlignments(); sv.constructClusterDiagrams(); sv.precomputePriors(); sv.fillNonSingletonArray()
    sv.runMCMC(None if sv.SAVE_SPACE else sv.outputname + ".mcmc.txt")
    sv.writeFinalFiles()
    if not sv.SAVE_SPACE:
        sv.writeStateCounts()
    return sv


if __name__ == "__main__":
    import sys
    main(sys.argv[1:])
