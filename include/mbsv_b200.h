/*
 * mbsv_b200.h — C ABI of libmbsv_b200.so, the B200-native (CUDA sm_100a) drop-in for the MCMC sampler of
 * raphael-group/multibreak-sv.
 *
 * The reference has no plugin/FFI interface; the seam this library replaces is
 *     MultiBreakSV.runMCMC(String)                          src/MultiBreakSV.java:773-877
 * (initializeMatrix :990-1099, iterate :1101-1182, naiveMove :1184-1289, svMove :1291-1353,
 *  incrementAllCounters :879-988, MultiBreakSVAssignmentMatrix.computeLogProb/setBreakpoints), called once from
 * main (:137-141) after fillNonSingletonArray (:128) and before writeFinalFiles (:144).
 * A Java host binds these entry points with Panama FFM (java.lang.foreign) — see INTEGRATION.md; in this
 * repository they are driven by the C++ host mirror (mbsv_host_*) and by Python ctypes.
 *
 * Conventions
 *   - primitives and pointers only; no callbacks; no structs by value; every entry point returns int
 *     (0 = OK, negative = mbsv_status) except where noted.
 *   - the caller owns every buffer it passes (inputs and outputs); mbsv_problem_create copies its inputs,
 *     so they may be released as soon as it returns.  The library owns the device memory behind the
 *     opaque handle.  mbsv_last_error() is library-owned, thread-local, valid until the next call.
 *   - no C++ exception or abort() crosses the boundary; nothing is printed.
 *   - one caller thread per handle (the reference is single-threaded with static state,
 *     MultiBreakSV.java:50-63,78); distinct handles are independent.  mbsv_run is synchronous.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with MBSV_ERR_CUDA.
 *
 * Host setup
 *   - a mbsv_problem handle is NOT thread-safe: it owns its launch plan, chain-state workspace, streams and events;
 *     one mbsv_run / mbsv_run_device at a time per handle (distinct handles may run concurrently from distinct threads).
 *   - a run overlaps one kernel per state-size class on up to 19 streams.  CUDA's default of 8 hardware queues makes the
 *     9th stream wait for the 1st; export CUDA_DEVICE_MAX_CONNECTIONS=32 in the host's environment before its first CUDA
 *     call (the variable is read when the context is created).  The library never modifies the environment.  Results do
 *     not depend on it.
 */
#ifndef MBSV_B200_H
#define MBSV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBSV_ABI_VERSION 1

typedef enum mbsv_status {
    MBSV_OK = 0,
    MBSV_ERR_ARG = -1,         /* bad argument / malformed descriptor */
    MBSV_ERR_CUDA = -2,        /* CUDA runtime error or no device */
    MBSV_ERR_NCCL = -3,        /* NCCL missing or failing (mbsv_comm_*, mbsv_reduce_tallies, mbsv_run_multi) */
    MBSV_ERR_OOM = -4,         /* host or device allocation failed */
    MBSV_ERR_INVARIANT = -5,   /* the reference would have thrown Error (MultiBreakSV.java:1122-1130,1142-1144;
                                  MultiBreakSVAssignmentMatrix.java:375-376) */
    MBSV_ERR_UNSUPPORTED = -6, /* input the reference itself mishandles (see DESIGN.md "rejected inputs") */
    MBSV_ERR_IO = -7           /* host mirror: file could not be read / parsed */
} mbsv_status;

typedef enum mbsv_mode {
    MBSV_MODE_REPLAY_EXACT = 0, /* the reference MH chain with its quirks; every proposal re-sums the whole
                                   log-posterior in the reference's order (supports.keySet(), experiments.keySet()) */
    MBSV_MODE_FAST = 1,         /* same RNG stream and proposals; Delta-log-posterior; difference tallies */
    MBSV_MODE_GIBBS = 2         /* NOT the reference's chain (statistical tier only): random-scan heat-bath
                                   Gibbs on the same posterior, 4 / 8 / 32 lanes per (subproblem, chain), the candidate alignments of
                                   the redrawn fragment across those lanes.  One iteration = one fragment redrawn from its exact
                                   conditional (counters: [0] = [4] = updates, [1] = updates that changed the state).  Same
                                   outputs; no trace.  Connected-component clusters and states beyond shared memory (a warp
                                   per chain, state in the global workspace) are supported */
} mbsv_mode;

/*
 * One batch of S >= 1 independent problems ("subproblems": what one JVM run of the reference sees, e.g. one
 * subset-N.clusters file of generateIndependentSubproblems.pl).  Fragments, alignments, alignment-ESP
 * entries, clusters and AlterSV entries of a subproblem are contiguous; all indices below are global.
 *
 * What each array is in the reference (all file:line relative to src/):
 *   sub_frag_ptr      fragments of subproblem s = [sub_frag_ptr[s], sub_frag_ptr[s+1]), listed in
 *                     fragments.keySet() iteration order (MultiBreakSV.java:62,999,1401)
 *   sub_cluster_ptr   clusters of subproblem s, in breakpoints.keySet() order (MultiBreakSV.java:61,750,892)
 *   sub_sv_ptr        entries of bpsWithNonSingletonSupport (MultiBreakSV.java:65,748-771), construction order
 *   frag_aln_ptr      Fragment.fragmentalignments after intersectClustersAndFragments (:484-505); the index
 *                     inside a fragment is the AssignmentIndex of the output file
 *   aln_numerrors/len FragmentAlignment.numerrors / .len (FragmentAlignment.java:33-34)
 *   aln_exp           index of FragmentAlignment.exp in experiments.keySet() order
 *   aln_esp_ptr       FragmentAlignment.alignments (the ESPs of the alignment)
 *   aln_esp_cluster   cluster of esp2cid.get(esp) (MultiBreakSV.java:56,1620) when selecting the alignment
 *                     makes that ESP count towards the cluster's support (the ESP is a leaf of the cluster's
 *                     diagram and esp2fragmentID.get(esp) is this fragment, AM.java:291-315); otherwise -1
 *   aln_esp_leaf      leaf slot of that ESP inside its cluster (used for connected components only)
 *   cluster_kind      0 = cluster (GASV localization != -1), 1 = connected component (MultiBreakSV.java:452-455)
 *   supports_order    clusters of each subproblem in A.supports.keySet() order (AM.java:53-66,137): the
 *                     summation order of computeLogProb
 *   cluster_hist_ptr  where cluster c's support histogram lives in results.cluster_hist; must provide at
 *                     least (number of leaves + 1) bins (MultiBreakSVBreakpoint.supportslong)
 *   cluster_leaf_ptr, leaf_owner          diagram.leaves and esp2fragmentID of each leaf (-1: no such fragment)
 *   cluster_root_ptr, root_leaf_ptr/_leaf diagram.roots and their ESPs as leaf slots (MultiBreakSVDiagram.java:56-139)
 *   sv_cluster, sv_frag_ptr, sv_frag      AlterSV: the distinct single-alignment fragments toggled for each
 *                     eligible cluster (MultiBreakSV.java:1303-1330); sv_numchanged = numchanged (:1311)
 *   exp_pseq, exp_lambda, logprobcov      Experiment.pseq / .lambdaD / .logprobcov[0..max_support]
 *                     (Experiment.java:32-35, MultiBreakSV.java:612-614), experiments.keySet() order
 */
typedef struct mbsv_problem_desc {
    int32_t abi_version;     /* MBSV_ABI_VERSION */
    int32_t n_sub;
    int32_t n_frag;
    int32_t n_aln;
    int32_t n_aln_esp;
    int32_t n_cluster;
    int32_t n_exp;
    int32_t max_support;     /* logprobcov row length - 1 (MultiBreakSV.java:87,426-427) */
    int32_t n_sv;
    int32_t n_sv_frag;
    int32_t n_leaf;
    int32_t n_root;
    int32_t n_root_leaf;
    int32_t n_hist;          /* length of results.cluster_hist */
    const int32_t* sub_frag_ptr;     /* [n_sub+1] */
    const int32_t* sub_cluster_ptr;  /* [n_sub+1] */
    const int32_t* sub_sv_ptr;       /* [n_sub+1], may be NULL when n_sv == 0 */
    const int32_t* frag_aln_ptr;     /* [n_frag+1] */
    const int32_t* aln_numerrors;    /* [n_aln] */
    const int32_t* aln_len;          /* [n_aln] */
    const int32_t* aln_exp;          /* [n_aln] */
    const int32_t* aln_esp_ptr;      /* [n_aln+1] */
    const int32_t* aln_esp_cluster;  /* [n_aln_esp] */
    const int32_t* aln_esp_leaf;     /* [n_aln_esp], may be NULL when no connected components */
    const uint8_t* cluster_kind;     /* [n_cluster] */
    const int32_t* supports_order;   /* [n_cluster] */
    const int32_t* cluster_hist_ptr; /* [n_cluster+1] */
    const int32_t* cluster_leaf_ptr; /* [n_cluster+1] */
    const int32_t* leaf_owner;       /* [n_leaf] */
    const int32_t* cluster_root_ptr; /* [n_cluster+1] */
    const int32_t* root_leaf_ptr;    /* [n_root+1] */
    const int32_t* root_leaf;        /* [n_root_leaf] */
    const int32_t* sv_cluster;       /* [n_sv] */
    const int32_t* sv_frag_ptr;      /* [n_sv+1] */
    const int32_t* sv_frag;          /* [n_sv_frag] */
    const int32_t* sv_numchanged;    /* [n_sv] */
    const double* exp_pseq;          /* [n_exp] */
    const double* exp_lambda;        /* [n_exp] */
    const double* logprobcov;        /* [n_exp][max_support+1] */
} mbsv_problem_desc;

typedef struct mbsv_run_params {
    int32_t mode;            /* mbsv_mode */
    int32_t n_chains;        /* chains per subproblem; chain c is seeded seed + c */
    int32_t num_iters;       /* --numiterations (MultiBreakSV.java:83,228) */
    int32_t burnin;          /* numiters/10, 0 -> 10 (MultiBreakSV.java:186-193); the long-burn-in window is
                                iter > num_iters - burnin (:903,926,962) */
    double perr;             /* --perr (MultiBreakSV.java:50,229) */
    int64_t seed;            /* 123456 in the reference (MultiBreakSV.java:78) */
    int32_t device;          /* CUDA device ordinal */
    int32_t java_int_wrap;   /* 1: per-experiment error/length totals wrap like Java int (AM.java:38-39,218-219) */
    const int32_t* init_assign; /* [n_frag] initial assignment for every chain (the --matchfile path,
                                   MultiBreakSV.java:1030-1085; no RNG draws are consumed), or NULL for the
                                   random initialisation of :999-1015 */
    int32_t trace;           /* 1: fill the trace_* arrays (mcmc.txt columns + accepted-move log) */
    int32_t memo_states;     /* MBSV_MODE_FAST: subproblems without connected components whose joint state space
                                prod(n_f + 1) is at most memo_states evaluate a proposal with the full log-posterior
                                of the proposed state, i.e. REPLAY arithmetic — the reference's own memo of state
                                log-probabilities (MultiBreakSV.java:58,1133-1137); the device keeps one table per
                                subproblem (shared memory when small, else device memory, at most a quarter of the free
                                device memory in total).  The arithmetic depends only on this value, never on the chain
                                count or the GPU.  0 disables; 65536 is the tuned default. */
} mbsv_run_params;

/* Output buffers (caller-allocated host memory; NULL pointers are skipped). */
typedef struct mbsv_results {
    uint32_t* aln_count;       /* [n_aln]  long-window count per alignment, pooled over chains:
                                  Fragment.avgAssignmentslong[j] * burnin (MultiBreakSV.java:962-969,856-860) */
    uint32_t* frag_err_count;  /* [n_frag] Fragment.avgNoassignmentlong * burnin */
    uint32_t* cluster_hist;    /* [n_hist] MultiBreakSVBreakpoint.supportslong[1..]; bin 0 is left 0, the host
                                  derives it as the reference does (MultiBreakSV.java:1764-1771) */
    int32_t* final_assign;     /* [n_chains][n_frag] the returned assignment matrix */
    double* final_logprob;     /* [n_sub][n_chains] A.logprob after the last iteration (:839) */
    int64_t* counters;         /* [n_sub][n_chains][5] proposalCounts[0], moveCounts[0], proposalCounts[1],
                                  moveCounts[1] (:842-843), fragment-reassignments */
    uint64_t* rng_state;       /* [n_sub][n_chains] java.util.Random seed field after the run */
    int32_t* error;            /* [n_sub][n_chains] 0 or mbsv_status of that chain */
    double* trace_logprob;     /* [n_sub][n_chains][num_iters]   (trace only) */
    int32_t* trace_move_frag;  /* -1 no move; >= 0 fragment of an accepted Naive move; -(2+k) accepted AlterSV
                                  move of the subproblem's k-th eligible cluster */
    int32_t* trace_move_assign;/* new assignment of that fragment */
    int32_t* initial_assign;   /* [n_chains][n_frag] state after initializeMatrix */
    double* initial_logprob;   /* [n_sub][n_chains]  "Initialized Log Probability" (:789) */
    int64_t* totals;           /* [6] sums over all chains of the five counters above, then the first non-zero
                                  chain status; lets large batches skip the per-chain arrays */
    double kernel_ms;          /* out: device time of the sampler kernels (CUDA events) */
} mbsv_results;

typedef struct mbsv_problem mbsv_problem;

int mbsv_abi_version(void);
int mbsv_device_count(void);                      /* >= 0, or MBSV_ERR_CUDA */
const char* mbsv_last_error(void);
const char* mbsv_status_string(int status);

/* Validates and copies the descriptor, uploads the packed CSR to `device`. */
int mbsv_problem_create(const mbsv_problem_desc* desc, int32_t device, mbsv_problem** out);
int mbsv_problem_destroy(mbsv_problem* p);
/* The library keeps large device blocks of destroyed problems for the next problem on the same device (a caching
   allocator: one run after another does not pay a multi-gigabyte cudaMalloc / cudaFree each).  mbsv_trim returns them to
   the driver; device < 0 = every device.  An allocation that fails trims and retries by itself. */
int mbsv_trim(int32_t device);

/* Runs n_chains chains of num_iters iterations on every subproblem (blocking). */
int mbsv_run(mbsv_problem* p, const mbsv_run_params* params, mbsv_results* results);

/* Same, but every non-NULL pointer in `results` is DEVICE memory on params->device (e.g. a torch tensor);
   used by the multi-GPU path so that tallies can be reduced with NCCL without a host round trip. */
int mbsv_run_device(mbsv_problem* p, const mbsv_run_params* params, mbsv_results* results);

/* Introspection for benchmarks / tests. */
int mbsv_problem_stats(const mbsv_problem* p, int64_t* out, int32_t n);   /* {n_sub, n_frag, n_aln, n_cluster, CSR bytes on the
                                                                             device, largest chain state in rows, memoised
                                                                             subproblems of the last plan, n_hist, device} */
int64_t mbsv_last_launch_count(void);             /* kernels launched by the last mbsv_run on this thread */
/* out[i*5..] = {tile rows per warp of the size class (0 = global-workspace class, negative = memoised kernel), warps, grid, dynamic shared
   memory bytes, device milliseconds} for each of those launches; returns their number. */
int mbsv_last_launch_info(double* out, int32_t max_launches);
/* out[i*6..] = {Naive proposals, Naive moves, AlterSV proposals, AlterSV moves, fragment-reassignments, first non-zero chain
   status} summed over the chains of launch i (same order as mbsv_last_launch_info); returns their number. */
int mbsv_last_launch_totals(int64_t* out, int32_t max_launches);

/* Multi-GPU sharding of independent subproblems (the reference's only scale-out is one JVM
   per subset-N.clusters file, doc/MultiBreakSV-Manual.txt:411-473): longest-processing-time-first greedy —
   subproblems by descending weight (ties: lower index first), each to the currently least loaded of n_parts
   (ties: lowest part).  out_part[i] in [0, n_parts). */
int mbsv_lpt_shard(const int64_t* weights, int64_t n, int32_t n_parts, int32_t* out_part);

/*
 * Giant component over several GPUs.  One connected component cannot be split spatially — every
 * fragment is coupled through the per-experiment totals (MultiBreakSVAssignmentMatrix.java:203-230) — so its CHAINS are
 * split over the GPUs (mbsv_split_chains; chain c keeps the seed `seed + c` wherever it runs, so chain 0 stays the
 * reference's chain) and the integer tallies are summed with ONE ncclReduce over NVLink.  The reference has no
 * counterpart: its manual only says that a few subproblems are "quite large" (doc/MultiBreakSV-Manual.txt:420-423).
 * NCCL is loaded on first use (dlopen libnccl.so.2); without it these entry points fail with MBSV_ERR_NCCL.
 */
#define MBSV_COMM_ID_BYTES 128
typedef struct mbsv_comm mbsv_comm;

/* (first chain, number of chains) of part `part` when n_chains chains are split over n_parts GPUs */
int mbsv_split_chains(int32_t n_chains, int32_t n_parts, int32_t part, int32_t* first, int32_t* count);

/* One host process per GPU: rank 0 draws an id (ncclGetUniqueId), the host passes it to every rank by its own means,
   every rank creates its communicator, runs its chains with mbsv_run_device into ONE contiguous device buffer
   [aln_count | frag_err_count | cluster_hist] (mbsv_problem_stats: n_aln + n_frag + n_hist words) and calls
   mbsv_reduce_tallies: one ncclReduce(sum, u32) onto `root`, in place.  reduce_ms = device time of the reduce. */
int mbsv_comm_unique_id(uint8_t* out /* [MBSV_COMM_ID_BYTES] */);
int mbsv_comm_create(const uint8_t* id, int32_t n_ranks, int32_t rank, int32_t device, mbsv_comm** out);
int mbsv_comm_destroy(mbsv_comm* c);
int mbsv_reduce_tallies(mbsv_comm* c, uint32_t* tallies_dev, int64_t n_words, int32_t root, double* reduce_ms);

/* One host process, n_dev GPUs (a JVM calling through Panama FFM): problems[d] is the SAME problem created on device d.
   Splits params->n_chains over the devices, runs them concurrently, pools the tallies on problems[0]'s device with one
   ncclReduce per device inside one group call, and fills results->aln_count / frag_err_count / cluster_hist / totals
   (host buffers; every per-chain array must be NULL).  params->device is ignored. */
int mbsv_run_multi(mbsv_problem* const* problems, int32_t n_dev, const mbsv_run_params* params, mbsv_results* results,
                   double* reduce_ms);

/* Host-side math the Java host would otherwise do itself (fdlibm-equivalent, no FMA contraction). */
double mbsv_log(double x);
double mbsv_exp(double x);
/* logprobcov[k] = logPoisson(k; lambda) for k = 1..max_support, [0] = 0 (MultiBreakSV.java:612-614,1644-1649) */
int mbsv_fill_logprobcov(double lambda, int32_t max_support, double* out);

#ifdef __cplusplus
}
#endif
#endif /* MBSV_B200_H */
