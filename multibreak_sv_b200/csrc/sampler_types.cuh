// Plain-data launch arguments of the sampler kernels (shared by the host API and the kernel translation units).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mbsv {

struct DevProblem {
    int n_sub, n_frag, n_aln, n_cluster, n_exp, max_support, maxn;
    const int* sub_frag_ptr;
    const int* sub_cluster_ptr;
    const int* sub_sv_ptr;
    const double* sub_logF;          // log(F_s)  (Math.log(fragments.size()), MultiBreakSV.java:1219-1271)
    const double* sub_neglogNsv;     // -log(bpsWithNonSingletonSupport.size())  (MultiBreakSV.java:1342-1343)
    const int* frag_aln_ptr;
    const int4* aln_rec;             // {numerrors, len, first ESP entry, exp | nesp << 8}
    const int* aln_esp_slot;         // per alignment-ESP entry: local count slot cluster*E + exp, or -1
    const int* supports_order_local; // per subproblem: local cluster indices in A.supports.keySet() order
    // connected components (all NULL-safe: only dereferenced by the CC kernels)
    const int* cluster_cc_off;       // per cluster: offset of its multiset block in the chain's cc region, -1 = simple
    const int* sub_cc_scratch;       // per subproblem: offset of the set-cover scratch in the cc region, -1 = no conn-comp
    const int* sub_cc_maxleaf;       // per subproblem: max leaves of a conn-comp (size of the leaf-flag scratch)
    const int* cluster_leaf_ptr;
    const int* leaf_owner;
    const int* aln_esp_leaf;
    const int* cluster_root_ptr;
    const int* root_leaf_ptr;
    const int* root_leaf;
    const int* cluster_hist_ptr;
    const int* slot_hist_ptr;        // per (cluster, experiment) count slot: cluster_hist_ptr of its cluster (saves slot / E)
    const int* sv_frag_ptr;
    const int* sv_frag;
    const int* sv_numchanged;
    // flattened views for the speculative-window kernel (giant_kernel.cuh): every load address of a proposal's gather is
    // known one round trip after the proposal is drawn
    const int4* aln_ent4;            // per alignment: its counted slots in ESP order, -1 padded (built on first use)
    const int4* sv_mv;               // per sv_frag entry: {fragment (local), numerrors, len, exp} of its only alignment
    const int* sv_ent_ptr;           // per AlterSV candidate: its counted slot entries ...
    const int2* sv_ent;              // ... {slot, index of the fragment in the candidate's list}, in application order
    const double* T;                 // logprobcov [E][max_support+1]
    const double* logint;            // log(n), n = 1..maxn (maxn >= max alignments per fragment and the
                                     // largest n the exact-binomial branch can see, capped at 65536)
    const double* invint;            // (double)1 / n, n = 0..maxn (the cumulative thresholds of naiveMove, MultiBreakSV.java:1205,1251)
    const double* exp_pseq;
    const double* exp_logp;
    const double* exp_log1mp;
    double log2pi;                   // Math.log(2)+Math.log(Math.PI)
};

struct DevRun {
    int n_chains, num_iters, burnin, win_base, java_int_wrap;
    double perr, logperr, log1mperr;
    unsigned long long perr_thr53;   // ceil(perr * 2^53): nextDouble() < perr  <=>  bits53() < perr_thr53
    long long seed;
    const int* init_assign;
    // plan: item g = item_base + (warp - warp_base)*lanes_per_warp + lane is chain (g % n_chains) of subproblem
    // sub_order[g / n_chains]; the memoised group and the general group of a run have their own (base, lanes)
    const int* sub_order;            // subproblems by ascending state size
    const int* sub_W;                // state rows per chain of each subproblem
    const int* sub_memo;             // per subproblem: joint states if it takes the memoised path, else 0 (or NULL)
    const long long* sub_tab_off;    // memoised kernel: offset of the subproblem's table in memo_global, -1 = shared memory
    double* memo_global;
    long long n_items;               // n_sub * n_chains
    int memo_mixed;                  // memoised kernel: 0 = a warp holds chains of ONE subproblem (strides/tables shared by
                                     // the warp), 1 = every lane may hold another subproblem (strides in the lane's own
                                     // column after its state rows, all tables in memo_global)
    int lanes_per_warp;              // chains per warp: 32, or fewer for small batches (more warps hide the serial
                                     // dependency chain of a chain better than full warps do when the GPU is not full)
    const long long* warp_row_off;   // [warp - global_warp0] row offset into workspace (global class)
    int* workspace;
    int warp_begin, warp_end, rows_per_warp, global_warp0;
    int warp_base;                   // first warp of the group this launch belongs to
    long long item_base;             // first item of that group
    // outputs (device)
    unsigned int* aln_count;
    unsigned int* frag_err_count;
    unsigned int* cluster_hist;
    int* final_assign;
    double* final_logprob;
    long long* counters;
    unsigned long long* rng_state;
    int* error;
    double* trace_logprob;
    int* trace_move_frag;
    int* trace_move_assign;
    int* initial_assign;
    double* initial_logprob;
    unsigned long long* totals;      // [6]
};

// MultiBreakSV.java:1661-1695 (logBinomial, logbinomialcoefficient, logNormal)

// Memoised kernel: tiles up to MEMO_SMEM_TABLE_ROWS rows keep the log-posterior table (and, for subproblems without AlterSV
// candidates of at most MEMO_RT_MAX_STATES joint states, or with one fragment, the acceptance-ratio table) in shared memory.
// Host planner (mbsv_core.cu ensure_plan) and kernel (memo_kernel.cuh) evaluate the same predicate.
constexpr int MEMO_SMEM_TABLE_ROWS = 48;
#ifndef MBSV_RT_MAX
#define MBSV_RT_MAX 32
#endif
constexpr int MEMO_RT_MAX_STATES = MBSV_RT_MAX;

using KernelFn = void (*)(const DevProblem, const DevRun);
// kernel selectors, one per translation unit (kern_fast.cu / kern_replay.cu / kern_trace.cu)
KernelFn get_kernel_fast(bool smem, int n_exp, int rows, bool narrow_ok, size_t* extra_smem, int* state_bytes);   // MODE_FAST, no trace: the throughput path (rows = shared-memory class)
KernelFn get_kernel_replay(bool smem, int n_exp);               // MODE_REPLAY_EXACT, no trace
KernelFn get_kernel_trace(int mode, bool smem, int n_exp);      // either mode with the per-iteration trace
KernelFn get_kernel_cc(int mode, bool trace, bool smem);        // batches with connected-component clusters
KernelFn get_kernel_packed(int mode, bool cc);                   // one chain per warp, contiguous state in the global workspace
KernelFn get_kernel_gibbs(int group, bool global_state);                            // heat-bath sampler, `group` (4, 8, 32) lanes per chain; global_state: a warp per chain, state in the workspace (gibbs_kernel.cuh)
KernelFn get_kernel_memo(bool trace, bool mixed);               // memoised subproblems (memo_kernel.cuh)
// speculative-window kernel, one CTA per (subproblem, chain) (giant_kernel.cuh): FAST, no trace, no connected components
KernelFn get_kernel_giant(int n_exp, int cluster, int* threads, size_t* smem_bytes);   // cluster = CTAs per chain: 1, 2, 4, 8
int giant_window_offsets();                                      // RNG offsets one CTA evaluates per window
int giant_max_entries();                                         // most count entries / toggled fragments of one proposal it takes
int giant_max_aln_entries();                                     // most counted slots of one alignment it takes
cudaError_t launch_giant_prep(const DevProblem& P, int4* aln_ent4, cudaStream_t stream);   // fills aln_ent4
// fills the device-memory log-posterior tables of the memoised subproblems listed in tab_sub
cudaError_t launch_memo_build(const DevProblem& P, const DevRun& R, int n_tab, const long long* state_ptr, const int* tab_sub,
                              long long n_states, int scratch_rows, cudaStream_t stream);

}  // namespace mbsv
