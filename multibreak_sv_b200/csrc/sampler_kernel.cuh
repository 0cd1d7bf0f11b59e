// Device side of the MCMC sampler: one thread runs one (subproblem, chain) Metropolis-Hastings chain of
// the reference (MultiBreakSV.java:773-1353) to completion.
//
// Execution model (DESIGN.md "kernels"):
//   * a warp holds 32 chains; the per-chain state (assignment of every fragment, selected-ESP count of
//     every (cluster, experiment)) is a column of a [rows][32] int32 tile, so lane l only ever touches
//     bank l: the random accesses of the chain are shared-memory-bank-conflict free by construction;
//   * the tile lives in shared memory when it fits (SMEM=true, one launch per size class) and in an
//     L2-resident global workspace with the same interleaving for large subproblems (SMEM=false);
//   * FAST mode skips lazy iterations in a tight RNG loop so that lanes re-converge on the proposal code,
//     evaluates the Delta-log-posterior with do/undo on the counts, and keeps O(1) difference tallies
//     (an accepted move at iteration t adds +(t-base) to the bin it leaves and -(t-base) to the bin it
//     enters; the final state adds N-base) with global atomics — integer-identical to the reference's
//     per-iteration incrementAllCounters (MultiBreakSV.java:879-988) for the long-burn-in window;
//   * REPLAY_EXACT re-sums the whole log-posterior in the reference's order for every proposal.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "sampler_common.cuh"

namespace mbsv {

// ---- build parameters (tuning aids; the defaults are the measured optima, profiles/r02_general_kernel_ab.json) ----
// CTAs per SM the registers are bounded for: shared-memory classes (see MINB below for the per-class exceptions) and the
// global-workspace class.  6-8 CTAs/SM (80-64 registers) spill and are 1.6-2.1x slower.
#ifndef MBSV_MIN_BLOCKS
#define MBSV_MIN_BLOCKS 5
#endif
#ifndef MBSV_MIN_BLOCKS_GLOBAL
#define MBSV_MIN_BLOCKS_GLOBAL 5
#endif
// logBinomial / exp inlined into the proposal code of the global-workspace class (406 -> 364 ms; the shared-memory classes
// lose 20 % with it: instruction cache)
#ifndef MBSV_INLINE_MATH_GLOBAL
#define MBSV_INLINE_MATH_GLOBAL 1
#endif
// AlterSV proposals are held until this many lanes of the warp have one pending (1 = run them as they come)
#ifndef MBSV_SV_BATCH
#define MBSV_SV_BATCH 8
#endif
// Built, measured and dropped in round 2 (every class slower): read-only evaluation of Naive proposals without tentative write
// and undo pass (register lists -> ~1 KB of spills per thread; profiles/r02_readonly_eval_ab.json), and drawing the NEXT
// proposal's header right after the current acceptance coin so that its fragment loads overlap the current evaluation, with L1
// prefetches of its alignment records (15-40 % slower: the carried header costs registers the kernel does not have).  What
// stayed from the latter: the counted slots of an alignment come from one 16-byte record (aln_ent4).

// A per-thread array of N values indexed at run time: in registers (selected by compare chains) or, INSMEM, in the thread's
// column of a [N][128] block of shared memory (plain address arithmetic, and no registers held across the proposal).
template <typename T, int N, bool INSMEM> struct LaneArr;
template <typename T, int N> struct LaneArr<T, N, false> {
    T v[N];
    __device__ __forceinline__ explicit LaneArr(void*) {}
    __device__ __forceinline__ T get(int x) const {
        T r = v[0];
#pragma unroll
        for (int y = 1; y < N; ++y) if (y == x) r = v[y];
        return r;
    }
    __device__ __forceinline__ void set(int x, T val) {
#pragma unroll
        for (int y = 0; y < N; ++y) if (y == x) v[y] = val;
    }
    __device__ __forceinline__ void add(int x, T d) {
#pragma unroll
        for (int y = 0; y < N; ++y) if (y == x) v[y] += d;
    }
};
template <typename T, int N> struct LaneArr<T, N, true> {
    T* p;
    __device__ __forceinline__ explicit LaneArr(void* base) : p((T*)base + threadIdx.x) {}
    __device__ __forceinline__ T get(int x) const { return p[x * 128]; }
    __device__ __forceinline__ void set(int x, T val) { p[x * 128] = val; }
    __device__ __forceinline__ void add(int x, T d) { p[x * 128] += d; }
};

// CC = the batch contains connected-component clusters (GASV localization -1): their support is the sorted
// multiset of a greedy set cover (MultiBreakSVAssignmentMatrix.java:337-409) kept per chain next to the counts.
// MINB = CTAs per SM the register allocation is bounded for (0 = the build default).  The kernel wants ~190 registers; at 5
// CTAs/SM it gets 96 and spills, and the spills cost more than the occupancy buys once the state tile limits a class to 3 or
// 4 CTAs/SM anyway (measured per class on B200, 6000 iterations: 128 rows 87 -> 68 ms at MINB 3, 96 rows 129 -> 98 ms and
// 64 rows 118 -> 102 ms at MINB 4; 48 rows and below are fastest at 5; 6 and 8 CTAs/SM are 1.6x and 2.1x slower).
// SMTOT = the per-experiment totals and logBinomial terms of the chain live in shared memory (4 * MAXE * 8 bytes per thread
// behind the counters) instead of registers.
// ST_T = type of a state word (assignment of a fragment: -1 .. alignments - 1; count of a slot: 0 .. maxSupport).  int everywhere
// except the FAST kernels of kern_fast8.cu, which run with signed char when every fragment has at most 127 alignments and
// maxSupport <= 127 (the host checks): a quarter of the shared memory per tile (more L1 for the CSR, larger subproblems in
// shared memory) and a quarter of the global workspace (more of it stays in L2).
template <int MODE, bool TRACE, bool SMEM, int MAXE, bool CC = false, bool PACKED = false, int MINB = 0, bool SMTOT = false,
          typename ST_T = int>
__global__ void __launch_bounds__(128, MINB ? MINB : SMEM ? MBSV_MIN_BLOCKS : MBSV_MIN_BLOCKS_GLOBAL) mbsv_chain_kernel(const __grid_constant__ DevProblem P,
                                                         const __grid_constant__ DevRun R) {
    extern __shared__ int smem_tile[];
    // logBinomial / exp inlined into the proposal code: measured faster for the global-workspace class only (fewer values
    // saved around the calls; its warps wait on memory, not on instruction fetch), slower for the shared-memory classes
    constexpr bool INLINE_MATH = !SMEM && !PACKED && MODE == 1 && !TRACE && !CC && (MBSV_INLINE_MATH_GLOBAL != 0);
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int warp = R.warp_begin + blockIdx.x * 4 + wic;
    if (warp >= R.warp_end) return;
    // Lanes beyond the last item stay in the warp as idle passengers (zero-sized problem, no iterations, no
    // writes): the main loop below synchronises the full warp.
    const long long item_raw = R.item_base + (long long)(warp - R.warp_base) * R.lanes_per_warp + lane;
    const bool alive = lane < R.lanes_per_warp && item_raw < R.n_items;
    const long long item = alive ? item_raw : R.n_items - 1;
    const int sub = R.sub_order[item / R.n_chains];
    const int chain = (int)(item % R.n_chains);
    // A chain's state is a column of a [rows][32] tile (conflict-free banks in shared memory, coalesced rows in the global
    // workspace).  PACKED (one chain per warp, global workspace): the column is stored contiguously, so that the few
    // chains of a giant component take 4 bytes per row instead of a 128-byte row each and stay L2-resident.
    constexpr int st_stride = (!SMEM && PACKED) ? 1 : 32;
    ST_T* st = SMEM ? (reinterpret_cast<ST_T*>(smem_tile) + (size_t)wic * R.rows_per_warp * 32 + lane)
                    : (reinterpret_cast<ST_T*>(R.workspace) + (size_t)R.warp_row_off[warp - R.global_warp0] * st_stride + (PACKED ? 0 : lane));
    const int tile_ints = SMEM ? R.rows_per_warp * 32 * (int)sizeof(ST_T) : 0;     // the four tiles of the CTA, in ints

#ifdef MBSV_BOUNDS_CHECK
    const int st_rows = !alive ? 0 : SMEM ? min(R.sub_W[sub], R.rows_per_warp) : R.sub_W[sub];
#endif
    const int f0 = P.sub_frag_ptr[sub], F = alive ? P.sub_frag_ptr[sub + 1] - f0 : 0;
    const int c0 = P.sub_cluster_ptr[sub], C = alive ? P.sub_cluster_ptr[sub + 1] - c0 : 0;
    const int E = MAXE <= 2 ? MAXE : P.n_exp;       // the buckets 1 and 2 are exact (kern_*.cu), so E is a constant there
    const int sv0 = P.sub_sv_ptr[sub], nsv = P.sub_sv_ptr[sub + 1] - sv0;
    const int Tstride = P.max_support + 1;
    const long long sc = (long long)sub * R.n_chains + chain;
    const int N = alive ? R.num_iters : 0;
    const double logF = P.sub_logF[sub];
    const double neglogF = -logF;
    const double perr = R.perr, logperr = R.logperr;
    const unsigned long long perr_thr = R.perr_thr53;
    constexpr unsigned long long NAIVE_THR = 0x1CCCCCCCCCCCCDULL;   // 0.9 * 2^53 exactly: nextDouble() < 0.9

    JavaRandom rng;
    rng.seed(R.seed + chain);

    // Memo-eligible subproblems (mbsv_run_params.memo_states) use REPLAY arithmetic; the table-driven version lives in
    // memo_kernel.cuh, a lane that lands here (chain count not a multiple of 32) re-sums the posterior per proposal.
    const int nst = (alive && R.sub_memo) ? R.sub_memo[sub] : 0;
    const bool use_full = (MODE == 0) || (nst > 0);

    char* const tot_base = (char*)(smem_tile + tile_ints + 5 * 128);
    LaneArr<long long, MAXE, SMTOT> Etot(tot_base), Ltot(tot_base + MAXE * 1024);
    LaneArr<double, MAXE, SMTOT> LB(tot_base + 2 * MAXE * 1024), newLB(tot_base + 3 * MAXE * 1024);
#pragma unroll
    for (int x = 0; x < MAXE; ++x) { Etot.set(x, 0); Ltot.set(x, 0); LB.set(x, 0.0); }
    int nerr = 0;
    int err = 0;
    // Proposal counters.  With SMTOT only the Naive proposals are counted in a register; the others change on a minority of
    // the rounds and live in shared memory behind the state tiles ([5][128] words, one column per thread): at 96 registers
    // the kernel spills (see MINB), and every value kept out of the registers is a spill less in the proposal loop.  The
    // classes with large tiles have registers to spare and need their shared memory: more of it per CTA pushes the SM to the
    // next carve-out step and halves the L1 that serves the CSR reads (measured: 128-row class 68 -> 88 ms).
    int n_naive_prop = 0;                        // <= num_iters
    LaneArr<int, 5, SMTOT> cnt(smem_tile + tile_ints);
    constexpr int C_NACC = 0, C_SVPROP = 1, C_SVACC = 2, C_RLO = 3, C_RHI = 4;   // Naive accepted, AlterSV proposed / accepted,
#pragma unroll
    for (int i = 0; i < 5; ++i) cnt.set(i, 0);                                    // fragments toggled by AlterSV proposals (64 bit)

    auto wrapi = [&](long long t) -> long long { return R.java_int_wrap ? (long long)(int)(unsigned int)(unsigned long long)t : t; };
    auto lb_of = [&](int x, long long e, long long l) -> double {
        if (INLINE_MATH) return log_binomial_impl(wrapi(e), wrapi(l), P.exp_pseq[x], P.exp_logp[x], P.exp_log1mp[x], P.log2pi, P.logint, P.maxn);
        return log_binomial_dev(wrapi(e), wrapi(l), P.exp_pseq[x], P.exp_logp[x], P.exp_log1mp[x], P.log2pi, P.logint,
                                P.maxn);
    };
    // Apply the count changes of fragment f moving from alignment `from` to `to` (-1 = error); dir=false
    // reverts without touching dcov.  Counts, totals and nerr are updated in place.
    auto move_counts = [&](int a0, int from, int to, double& dcov, double& derr, unsigned& touched) {
        if (from != -1) {
            const int4 rec = P.aln_rec[a0 + from];
            const int x = rec.w & 0xff, ne = rec.w >> 8;
            const double* Tx = P.T + (size_t)x * Tstride;
            const int4 e4 = P.aln_ent4[a0 + from];
            if (e4.x != -3) {
                const int sl[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (sl[j] < 0) continue;
                    const int n = MBSV_ST(F + sl[j]);
                    dcov = dcov + (Tx[n - 1] - Tx[n]);
                    MBSV_ST(F + sl[j]) = n - 1;
                }
            } else
            for (int j = 0; j < ne; ++j) {
                const int slot = P.aln_esp_slot[rec.z + j];
                if (slot < 0) continue;
                const int n = MBSV_ST(F + slot);
                dcov = dcov + (Tx[n - 1] - Tx[n]);
                MBSV_ST(F + slot) = n - 1;
            }
            Etot.add(x, -(long long)rec.x); Ltot.add(x, -(long long)rec.y);
            touched |= 1u << x;
        } else {
            derr = derr - logperr;
            --nerr;
        }
        if (to != -1) {
            const int4 rec = P.aln_rec[a0 + to];
            const int x = rec.w & 0xff, ne = rec.w >> 8;
            const double* Tx = P.T + (size_t)x * Tstride;
            const int4 e4 = P.aln_ent4[a0 + to];
            if (e4.x != -3) {
                const int sl[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (sl[j] < 0) continue;
                    const int n = MBSV_ST(F + sl[j]);
                    dcov = dcov + (Tx[n + 1] - Tx[n]);
                    MBSV_ST(F + sl[j]) = n + 1;
                }
            } else
            for (int j = 0; j < ne; ++j) {
                const int slot = P.aln_esp_slot[rec.z + j];
                if (slot < 0) continue;
                const int n = MBSV_ST(F + slot);
                dcov = dcov + (Tx[n + 1] - Tx[n]);
                MBSV_ST(F + slot) = n + 1;
            }
            Etot.add(x, rec.x); Ltot.add(x, rec.y);
            touched |= 1u << x;
        } else {
            derr = derr + logperr;
            ++nerr;
        }
    };
    // ---- connected components (CC only) ----------------------------------------------------------------
    // State after the counts: per conn-comp cluster {stamp, current multisets E x (len, values[nroots]),
    // saved multisets (same shape)}, then per-subproblem scratch {leaf flags, root flags}.
    const int cc_base = F + C * E;
    const bool has_cc = CC && alive && P.sub_cc_scratch[sub] >= 0;
    auto cc_nroots = [&](int cl) -> int { return P.cluster_root_ptr[c0 + cl + 1] - P.cluster_root_ptr[c0 + cl]; };
    auto cc_lpc = [&](int cl) -> double {
        const int nr = cc_nroots(cl);
        const int b0 = cc_base + P.cluster_cc_off[c0 + cl] + 1;
        double lpc = 0.0;
        for (int x = 0; x < E; ++x) {
            const int b = b0 + x * (1 + nr);
            const int len = MBSV_ST(b);
            for (int i = 0; i < len; ++i) lpc = lpc + P.T[(size_t)x * Tstride + MBSV_ST(b + 1 + i)];
        }
        return lpc;
    };
    // setBreakpoints for one connected component from the current assignments (AM.java:264-414)
    auto cc_recompute = [&](int cl) {
        const int cg = c0 + cl;
        const int l0 = P.cluster_leaf_ptr[cg], nl = P.cluster_leaf_ptr[cg + 1] - l0;
        const int r0 = P.cluster_root_ptr[cg], nr = P.cluster_root_ptr[cg + 1] - r0;
        const int b0 = cc_base + P.cluster_cc_off[cg] + 1;
        const int sl = cc_base + P.sub_cc_scratch[sub];       // leaf flags: exp + 1 (0 = not selected), bit 8 = found
        const int sr = sl + P.sub_cc_maxleaf[sub];            // root flags: 1 = already taken
        int numselected = 0;
        for (int l = 0; l < nl; ++l) {
            int flag = 0;
            const int fg = P.leaf_owner[l0 + l];
            if (fg >= 0) {
                const int a = MBSV_ST(fg - f0);
                if (a != -1) {
                    const int4 rec = P.aln_rec[P.frag_aln_ptr[fg] + a];
                    const int ne = rec.w >> 8;
                    for (int j = 0; j < ne; ++j)
                        if (P.aln_esp_slot[rec.z + j] == -(2 + cl) && P.aln_esp_leaf[rec.z + j] == l) { flag = (rec.w & 0xff) + 1; ++numselected; }
                }
            }
            MBSV_ST(sl + l) = flag;
        }
        for (int x = 0; x < E; ++x) MBSV_ST(b0 + x * (1 + nr)) = 0;
        if (numselected == 0) return;
        for (int x = 0; x < E; ++x) {
            const int b = b0 + x * (1 + nr);
            for (int l = 0; l < nl; ++l) MBSV_ST(sl + l) &= 0xff;
            for (int r = 0; r < nr; ++r) MBSV_ST(sr + r) = 0;
            int len = 0;
            for (;;) {
                int maxcount = 0, maxind = -1;
                for (int r = 0; r < nr; ++r) {
                    if (MBSV_ST(sr + r)) continue;
                    int cur = 0;
                    for (int j = P.root_leaf_ptr[r0 + r]; j < P.root_leaf_ptr[r0 + r + 1]; ++j)
                        if (MBSV_ST(sl + P.root_leaf[j]) == x + 1) ++cur;
                    if (cur > maxcount) { maxcount = cur; maxind = r; }
                }
                if (maxind == -1) {
                    for (int l = 0; l < nl; ++l) if (MBSV_ST(sl + l) == x + 1) err = -5;   // Error at AM.java:375-376
                    break;
                }
                for (int j = P.root_leaf_ptr[r0 + maxind]; j < P.root_leaf_ptr[r0 + maxind + 1]; ++j)
                    if ((MBSV_ST(sl + P.root_leaf[j]) & 0xff) == x + 1) MBSV_ST(sl + P.root_leaf[j]) = (x + 1) | 0x100;
                // sorted insert (Collections.sort at AM.java:403)
                int pos = len;
                while (pos > 0 && MBSV_ST(b + pos) > maxcount) { MBSV_ST(b + 1 + pos) = MBSV_ST(b + pos); --pos; }
                MBSV_ST(b + 1 + pos) = maxcount;
                ++len;
                MBSV_ST(sr + maxind) = 1;
                bool all = true;
                for (int l = 0; l < nl; ++l) if (MBSV_ST(sl + l) == x + 1) all = false;
                if (all) break;
            }
            MBSV_ST(b) = len;
        }
    };
    // visit the connected components touched by fragment-alignment `al` (entries coded -(2+cluster))
    auto cc_visit = [&](int al, auto&& fn) {
        const int4 rec = P.aln_rec[al];
        const int ne = rec.w >> 8;
        for (int j = 0; j < ne; ++j) {
            const int code = P.aln_esp_slot[rec.z + j];
            if (code < -1) fn(-code - 2);
        }
    };
    auto cc_hist = [&](int cl, int which, unsigned d) {   // which: 0 current multisets, 1 saved multisets
        const int nr = cc_nroots(cl);
        const int b0 = cc_base + P.cluster_cc_off[c0 + cl] + 1 + which * E * (1 + nr);
        unsigned int* h = R.cluster_hist + P.cluster_hist_ptr[c0 + cl];
        for (int x = 0; x < E; ++x) {
            const int b = b0 + x * (1 + nr);
            const int len = MBSV_ST(b);
            for (int i = 0; i < len; ++i) atomicAdd(&h[MBSV_ST(b + 1 + i)], d);
        }
    };

    // MultiBreakSVAssignmentMatrix.java:130-246 in the reference's summation order.  Clusters are taken four at a time so
    // that the index / count / table loads of a batch are in flight together (the chain state of a large subproblem lives
    // in device memory); the additions keep the reference's order: per cluster over experiments, then cluster by cluster.
    auto full_logprob = [&]() -> double {
        double lp = 0;
        constexpr int B = 4;
        for (int i = 0; i < C; i += B) {
            int cl[B];
            double lpc[B];
            bool any_cc = false;
#pragma unroll
            for (int k = 0; k < B; ++k) {
                cl[k] = (i + k < C) ? P.supports_order_local[c0 + i + k] : -1;
                lpc[k] = 0.0;
                if (CC && cl[k] >= 0 && P.cluster_cc_off[c0 + cl[k]] >= 0) any_cc = true;
            }
            if (CC && any_cc) {
#pragma unroll
                for (int k = 0; k < B; ++k) {
                    if (cl[k] < 0) continue;
                    if (P.cluster_cc_off[c0 + cl[k]] >= 0) lpc[k] = cc_lpc(cl[k]);
                    else for (int x = 0; x < E; ++x) {
                        const int n = MBSV_ST(F + cl[k] * E + x);
                        if (n > 0) lpc[k] = lpc[k] + P.T[(size_t)x * Tstride + n];
                    }
                }
            } else {
                for (int x = 0; x < E; ++x) {
                    int n[B];
#pragma unroll
                    for (int k = 0; k < B; ++k) n[k] = cl[k] >= 0 ? MBSV_ST(F + cl[k] * E + x) : 0;
#pragma unroll
                    for (int k = 0; k < B; ++k) if (n[k] > 0) lpc[k] = lpc[k] + P.T[(size_t)x * Tstride + n[k]];
                }
            }
#pragma unroll
            for (int k = 0; k < B; ++k) if (cl[k] >= 0) lp += lpc[k];
        }
        lp += 0.0;
        lp += (double)nerr * logperr;
#pragma unroll
        for (int x = 0; x < MAXE; ++x) if (x < E) lp += lb_of(x, Etot.get(x), Ltot.get(x));
        return lp;
    };
    // difference tallies
    const int win_base = R.win_base;
    auto tally_frag = [&](int f, int a0, int from, int to, int t) {
        if (t <= win_base) return;
        const unsigned d = (unsigned)(t - win_base);
        if (from == -1) atomicAdd(&R.frag_err_count[f0 + f], d); else atomicAdd(&R.aln_count[a0 + from], d);
        if (to == -1) atomicAdd(&R.frag_err_count[f0 + f], 0u - d); else atomicAdd(&R.aln_count[a0 + to], 0u - d);
    };
    // Walk an accepted move backwards one unit step at a time, emitting the support-histogram difference
    // tallies, then re-apply it (intermediate values telescope away).
    auto retally_fragment = [&](int a0, int from, int to, int t) {
        const unsigned d = (unsigned)(t - win_base);
        if (to != -1) {
            const int4 rec = P.aln_rec[a0 + to];
            const int ne = rec.w >> 8;
            for (int j = ne - 1; j >= 0; --j) {
                const int slot = P.aln_esp_slot[rec.z + j];
                if (slot < 0) continue;
                const int n = MBSV_ST(F + slot);
                unsigned int* h = R.cluster_hist + P.slot_hist_ptr[c0 * E + slot];
                if (n - 1 > 0) atomicAdd(&h[n - 1], d);
                if (n > 0) atomicAdd(&h[n], 0u - d);
                MBSV_ST(F + slot) = n - 1;
            }
        }
        if (from != -1) {
            const int4 rec = P.aln_rec[a0 + from];
            const int ne = rec.w >> 8;
            for (int j = ne - 1; j >= 0; --j) {
                const int slot = P.aln_esp_slot[rec.z + j];
                if (slot < 0) continue;
                const int n = MBSV_ST(F + slot);
                unsigned int* h = R.cluster_hist + P.slot_hist_ptr[c0 * E + slot];
                if (n + 1 > 0) atomicAdd(&h[n + 1], d);
                if (n > 0) atomicAdd(&h[n], 0u - d);
                MBSV_ST(F + slot) = n + 1;
            }
        }
    };
    auto reapply_fragment = [&](int a0, int from, int to) {
        if (from != -1) {
            const int4 rec = P.aln_rec[a0 + from];
            const int ne = rec.w >> 8;
            for (int j = 0; j < ne; ++j) { const int slot = P.aln_esp_slot[rec.z + j]; if (slot >= 0) MBSV_ST(F + slot) -= 1; }
        }
        if (to != -1) {
            const int4 rec = P.aln_rec[a0 + to];
            const int ne = rec.w >> 8;
            for (int j = 0; j < ne; ++j) { const int slot = P.aln_esp_slot[rec.z + j]; if (slot >= 0) MBSV_ST(F + slot) += 1; }
        }
    };

    // ---------------- initializeMatrix (MultiBreakSV.java:990-1099) ----------------
    {
        const int W = alive ? R.sub_W[sub] : 0;
        for (int i = 0; i < W; ++i) MBSV_ST(i) = 0;
        // phase A: the draws of :999-1015 (serial RNG); the assignment goes straight to the state tile
        for (int f = 0; f < F; ++f) {
            int a;
            if (R.init_assign) a = R.init_assign[f0 + f];
            else {
                if (rng.nextDouble() < perr) a = -1;
                else a = rng.nextInt(P.frag_aln_ptr[f0 + f + 1] - P.frag_aln_ptr[f0 + f]);
            }
            MBSV_ST(f) = a;
            if (R.initial_assign) R.initial_assign[(size_t)chain * P.n_frag + f0 + f] = a;
        }
        // phase B: counts and per-experiment totals, eight fragments at a time so that the dependent loads
        // (assignment -> alignment record -> ESP slots) of a batch overlap; counts in a device-memory tile are
        // bumped with fire-and-forget atomics (only this lane touches its column)
        nerr = 0;
        constexpr int IB = 8;
        for (int fb = 0; fb < F; fb += IB) {
            int a[IB], al[IB];
#pragma unroll
            for (int k = 0; k < IB; ++k) a[k] = (fb + k < F) ? MBSV_ST(fb + k) : -2;
#pragma unroll
            for (int k = 0; k < IB; ++k) al[k] = a[k] >= 0 ? P.frag_aln_ptr[f0 + fb + k] + a[k] : 0;
            int4 rec[IB];
#pragma unroll
            for (int k = 0; k < IB; ++k) rec[k] = a[k] >= 0 ? P.aln_rec[al[k]] : make_int4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < IB; ++k) {
                if (a[k] == -1) ++nerr;
                if (a[k] < 0) continue;
                const int x = rec[k].w & 0xff, ne = rec[k].w >> 8;
                for (int j = 0; j < ne; ++j) {
                    const int slot = P.aln_esp_slot[rec[k].z + j];
                    if (slot < 0) continue;
                    if constexpr (SMEM || sizeof(ST_T) != 4) MBSV_ST(F + slot) += 1; else atomicAdd(&MBSV_ST(F + slot), 1);
                }
                Etot.add(x, rec[k].x); Ltot.add(x, rec[k].y);
            }
        }
#pragma unroll
        for (int x = 0; x < MAXE; ++x) if (x < E) LB.set(x, lb_of(x, Etot.get(x), Ltot.get(x)));
        if (has_cc)
            for (int cl = 0; cl < C; ++cl) if (P.cluster_cc_off[c0 + cl] >= 0) cc_recompute(cl);
    }
    double logprob = full_logprob();
    if (alive && R.initial_logprob) R.initial_logprob[sc] = logprob;

    // ---------------- iterations (MultiBreakSV.java:796-837, 1101-1182) ----------------
    const int q_base = P.sv_frag_ptr[sv0];   // only dereferenced when nsv > 0
    int iter = 0;
    bool done = false, live = false;
    // header of the pending proposal: the draws that do not depend on the state (move type; fragment or AlterSV candidate)
    // and the loads that depend only on them (alignment range and current assignment of the fragment)
    bool h_naive = true;
    int h_idx = 0, h_a0 = 0, h_n = 0, h_prev = -1;
    // Move-type coin (:1115-1118): its value cannot matter when there is no AlterSV candidate; the fragment pick
    // (:1400) cannot matter for a one-fragment problem.  The draws are still consumed.
    auto draw_header = [&]() {
        h_naive = true;
        if (nsv == 0) rng.skip<2>(); else h_naive = rng.bits53() < NAIVE_THR;
        double u = 0.0;
        if (h_naive && F == 1) rng.skip<2>(); else u = rng.nextDouble();
        if (h_naive) {
            h_idx = (int)(u * (double)F);
            h_a0 = P.frag_aln_ptr[f0 + h_idx];
            h_n = P.frag_aln_ptr[f0 + h_idx + 1] - h_a0;
            h_prev = MBSV_ST(h_idx);
        } else {
            h_idx = (int)(u * (double)nsv);
        }
    };
    for (;;) {
        // Lazy chain: with probability 1/2 stay (:1106-1113).  Lazy iterations only advance the RNG (tallies
        // are difference tallies), so each lane skips them in a tight loop; the warp-wide vote below is the
        // convergence point that brings all 32 lanes back together for the proposal code.
        if (!live) {
            if (TRACE) {
                if (!done) {
                    while (iter < N) {
                        const double r = rng.nextDouble();
                        if (!(r < 0.5)) { live = true; break; }
                        R.trace_logprob[sc * N + iter] = logprob;
                        R.trace_move_frag[sc * N + iter] = -1;
                        R.trace_move_assign[sc * N + iter] = -1;
                        ++iter;
                    }
                    done = !live;
                }
            } else if (!done) {
                lazy_skip(rng, iter, N, live, done);
            }
            if (live) draw_header();
        }
        if (__all_sync(0xffffffffu, done && !live)) break;
        // AlterSV proposals (one in ten where a subproblem has candidates) walk a list of fragments while the Naive lanes of
        // the warp wait, and with 32 chains nearly every round has one.  Chains are independent, so a lane that drew an
        // AlterSV proposal HOLDS it until enough lanes of the warp hold one (or no other lane can advance); the list
        // walk then runs once for all of them.  Only the interleaving of the chains changes, not any chain.
        bool hold = false;
        if (MBSV_SV_BATCH > 1) {
            const unsigned svm = __ballot_sync(0xffffffffu, live && !h_naive);
            if (svm) {
                const unsigned prog = __ballot_sync(0xffffffffu, (live && h_naive) || (!live && !done));
                // a quarter of the chains the warp still runs, at most MBSV_SV_BATCH: narrow warps (small batches) and the
                // tail of a run must not leave a chain waiting for company that is not there
                const int unfinished = __popc(__ballot_sync(0xffffffffu, live || !done));
                const int need = min(MBSV_SV_BATCH, max(1, unfinished >> 2));
                if (__popc(svm) < need && prog != 0) hold = live && !h_naive;
            }
        }
        if (hold || !live) continue;
        live = false;
        const int t = iter;
        const bool naive = h_naive;
        // A proposal is a list of single-fragment moves: one for Naive, the cluster's single-alignment
        // fragments for AlterSV.  Both kinds run through the same code below.
        int f = -1, a0 = 0, prev = -1, now = -1, svk = -1, q0 = 0, nmoves = 1;
        double Alogq = 0.0, Anewlogq = 0.0;
        if (naive) {
            // ---- naiveMove (:1184-1273) ----
            ++n_naive_prop;
            f = h_idx;
            a0 = h_a0;
            prev = h_prev;
            naive_pick(rng, h_n, prev, neglogF, perr_thr, logperr, R.log1mperr, P.logint, P.invint, now, Alogq, Anewlogq);
            if (now == prev) { err = -5; done = true; continue; }   // Error at :1122-1123
        } else {
            // ---- svMove (:1291-1343) ----
            cnt.add(C_SVPROP, 1);
            svk = h_idx;
            {
                const unsigned lo = (unsigned)cnt.get(C_RLO), nlo = lo + (unsigned)P.sv_numchanged[sv0 + svk];
                cnt.set(C_RLO, (int)nlo);
                if (nlo < lo) cnt.add(C_RHI, 1);
            }
            q0 = P.sv_frag_ptr[sv0 + svk];
            nmoves = P.sv_frag_ptr[sv0 + svk + 1] - q0;
            Alogq = P.sub_neglogNsv[sub];
            Anewlogq = Alogq;
        }
        // ---- evaluate the proposal ----
        double dcov = 0.0, derr = 0.0;
        unsigned touched = 0;
        bool written = false;
        double newlogprob, arg;
        bool accept = false;
        // Pass 0 applies the proposal's moves tentatively and decides; pass 1, for rejected proposals only, applies the same
        // moves backwards (count changes commute, so the order within the list does not matter).  ONE copy of the move code
        // serves both: the proposal loop is longer than the SM's 32 KB instruction cache can hold for 20 warps at different
        // places in it (ncu: ~1.7 no-instruction stall cycles per issued instruction), every instruction less counts.
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            for (int q = 0; q < nmoves; ++q) {
                int g = f, ga0 = a0, from = prev, to = now;
                if (!naive) {
                    g = P.sv_frag[q0 + q] - f0;
                    ga0 = P.frag_aln_ptr[f0 + g];
                    const int cur = MBSV_ST(g);      // old value unless the CC path already wrote the new one
                    from = (pass && written) ? ((cur == -1) ? 0 : -1) : cur;
                    to = (from == -1) ? 0 : -1;
                }
                move_counts(ga0, pass ? to : from, pass ? from : to, dcov, derr, touched);
                if (pass && written) MBSV_ST(g) = from;
            }
            if (pass) break;
            if (has_cc) {
                // conn-comps derive their selected ESPs from the assignments: write them tentatively, then
                // recompute every touched component once (stamp = t + 1), saving the old multisets
                written = true;
                for (int q = 0; q < nmoves; ++q) {
                    if (naive) MBSV_ST(f) = now;
                    else { const int g = P.sv_frag[q0 + q] - f0; const int pv = MBSV_ST(g); MBSV_ST(g) = (pv == -1) ? 0 : -1; }
                }
                auto touch = [&](int cl) {
                    const int o = cc_base + P.cluster_cc_off[c0 + cl];
                    if (MBSV_ST(o) == t + 1) return;
                    MBSV_ST(o) = t + 1;
                    const int sz = E * (1 + cc_nroots(cl));
                    const double before = cc_lpc(cl);
                    for (int i = 0; i < sz; ++i) MBSV_ST(o + 1 + sz + i) = MBSV_ST(o + 1 + i);
                    cc_recompute(cl);
                    dcov = dcov + (cc_lpc(cl) - before);
                };
                for (int q = 0; q < nmoves; ++q) {
                    int ga0 = a0, from = prev, to = now;
                    if (!naive) {
                        const int g = P.sv_frag[q0 + q] - f0;
                        ga0 = P.frag_aln_ptr[f0 + g];
                        to = MBSV_ST(g);
                        from = (to == -1) ? 0 : -1;
                    }
                    if (from != -1) cc_visit(ga0 + from, touch);
                    if (to != -1) cc_visit(ga0 + to, touch);
                }
                if (err) { done = true; live = false; break; }   // set cover failed to cover: the reference throws Error
            }
            // ---- acceptance ratio (:1146) ----
            if (use_full) {
                // full re-sum in the reference's order; the assignment change only enters through nerr, counts, totals
#pragma unroll
                for (int x = 0; x < MAXE; ++x) newLB.set(x, LB.get(x));
                newlogprob = full_logprob();
                arg = (newlogprob + Alogq) - (logprob + Anewlogq);
            } else {
                double dlb = 0.0;
#pragma unroll
                for (int x = 0; x < MAXE; ++x) newLB.set(x, LB.get(x));
                // Each lane walks ITS touched experiments in ascending order (same summation order as before), so lanes
                // whose proposals touch different platforms evaluate logBinomial together instead of one platform at a time.
                unsigned mask = touched & ((1u << E) - 1u);
                while (mask) {
                    const int x = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const double nl = lb_of(x, Etot.get(x), Ltot.get(x));
                    newLB.set(x, nl);
                    dlb = dlb + (nl - LB.get(x));
                }
                const double delta = (dcov + derr) + dlb;
                newlogprob = logprob + delta;
                arg = (delta + Alogq) - Anewlogq;
            }
            const double e = INLINE_MATH ? jm_exp(arg) : exp_dev(arg);
            const double ratio = (e != e) ? e : (1.0 < e ? 1.0 : e);   // Math.min(1, e)
            const double r_acc = rng.nextDouble();
            accept = ratio > r_acc;
            if (accept) {
                logprob = newlogprob;
#pragma unroll
                for (int x = 0; x < MAXE; ++x) LB.set(x, newLB.get(x));
                cnt.add(naive ? C_NACC : C_SVACC, 1);
                if (t > win_base) {
                    // support-histogram difference tallies: walk the move backwards, then re-apply it
                    for (int q = nmoves - 1; q >= 0; --q) {
                        int ga0 = a0, from = prev, to = now;
                        if (!naive) {
                            const int g = P.sv_frag[q0 + q] - f0;
                            ga0 = P.frag_aln_ptr[f0 + g];
                            const int cur = MBSV_ST(g);     // old value unless the CC path already wrote the new one
                            from = written ? ((cur == -1) ? 0 : -1) : cur;
                            to = (from == -1) ? 0 : -1;
                        }
                        retally_fragment(ga0, from, to, t);
                    }
                    for (int q = 0; q < nmoves; ++q) {
                        int g = f, ga0 = a0, from = prev, to = now;
                        if (!naive) {
                            g = P.sv_frag[q0 + q] - f0;
                            ga0 = P.frag_aln_ptr[f0 + g];
                            const int cur = MBSV_ST(g);
                            from = written ? ((cur == -1) ? 0 : -1) : cur;
                            to = (from == -1) ? 0 : -1;
                        }
                        reapply_fragment(ga0, from, to);
                        tally_frag(g, ga0, from, to, t);
                    }
                }
                if (has_cc) {
                    // conn-comp multisets that changed: old elements leave, new ones enter; clear the stamps
                    const unsigned d = t > win_base ? (unsigned)(t - win_base) : 0u;
                    auto settle = [&](int cl) {
                        const int o = cc_base + P.cluster_cc_off[c0 + cl];
                        if (MBSV_ST(o) != t + 1) return;
                        MBSV_ST(o) = 0;
                        if (d) { cc_hist(cl, 1, d); cc_hist(cl, 0, 0u - d); }
                    };
                    for (int q = 0; q < nmoves; ++q) {
                        int ga0 = a0, from = prev, to = now;
                        if (!naive) {
                            const int g = P.sv_frag[q0 + q] - f0;
                            ga0 = P.frag_aln_ptr[f0 + g];
                            to = MBSV_ST(g);
                            from = (to == -1) ? 0 : -1;
                        }
                        if (from != -1) cc_visit(ga0 + from, settle);
                        if (to != -1) cc_visit(ga0 + to, settle);
                    }
                }
                if (!written) {
                    for (int q = 0; q < nmoves; ++q) {
                        if (naive) MBSV_ST(f) = now;
                        else {
                            const int g = P.sv_frag[q0 + q] - f0;
                            const int pv = MBSV_ST(g);
                            MBSV_ST(g) = (pv == -1) ? 0 : -1;
                        }
                    }
                }
                break;
            }
            // rejected: restore the saved multisets of the touched connected components, then revert the counts (and, on the CC
            // path, the tentative assignments) in pass 1
            if (has_cc) {
                auto restore = [&](int cl) {
                    const int o = cc_base + P.cluster_cc_off[c0 + cl];
                    if (MBSV_ST(o) != t + 1) return;
                    MBSV_ST(o) = 0;
                    const int sz = E * (1 + cc_nroots(cl));
                    for (int i = 0; i < sz; ++i) MBSV_ST(o + 1 + i) = MBSV_ST(o + 1 + sz + i);
                };
                for (int q = 0; q < nmoves; ++q) {
                    int ga0 = a0, from = prev, to = now;
                    if (!naive) {
                        const int g = P.sv_frag[q0 + q] - f0;
                        ga0 = P.frag_aln_ptr[f0 + g];
                        to = MBSV_ST(g);
                        from = (to == -1) ? 0 : -1;
                    }
                    if (from != -1) cc_visit(ga0 + from, restore);
                    if (to != -1) cc_visit(ga0 + to, restore);
                }
            }
        }
        if (err) continue;
        if (TRACE) {
            R.trace_logprob[sc * N + t] = logprob;
            R.trace_move_frag[sc * N + t] = accept ? (naive ? f0 + f : -(2 + svk)) : -1;
            R.trace_move_assign[sc * N + t] = (accept && naive) ? now : -1;
        }
        ++iter;
    }
    (void)q_base;

    // ---------------- epilogue: final tallies, state, counters ----------------
    if (!alive) return;
    if (err == 0) {
        const long long nwin = (long long)N - win_base;
        if (nwin > 0) {
            const unsigned d = (unsigned)nwin;
            // eight loads in flight per batch: the state of a large subproblem is in device memory
            constexpr int EB = 8;
            for (int fb = 0; fb < F; fb += EB) {
                int a[EB], a0[EB];
#pragma unroll
                for (int k = 0; k < EB; ++k) a[k] = (fb + k < F) ? MBSV_ST(fb + k) : -2;
#pragma unroll
                for (int k = 0; k < EB; ++k) a0[k] = a[k] >= 0 ? P.frag_aln_ptr[f0 + fb + k] : 0;
#pragma unroll
                for (int k = 0; k < EB; ++k) {
                    if (a[k] == -1) atomicAdd(&R.frag_err_count[f0 + fb + k], d);
                    else if (a[k] >= 0) atomicAdd(&R.aln_count[a0[k] + a[k]], d);
                }
            }
            const int nslot = C * E;
            for (int sb = 0; sb < nslot; sb += EB) {
                int n[EB], hp[EB];
#pragma unroll
                for (int k = 0; k < EB; ++k) {
                    const bool ok = sb + k < nslot && !(CC && P.cluster_cc_off[c0 + (sb + k) / E] >= 0);
                    n[k] = ok ? MBSV_ST(F + sb + k) : 0;
                }
#pragma unroll
                for (int k = 0; k < EB; ++k) hp[k] = n[k] > 0 ? P.slot_hist_ptr[c0 * E + sb + k] : 0;
#pragma unroll
                for (int k = 0; k < EB; ++k) if (n[k] > 0) atomicAdd(&R.cluster_hist[hp[k] + n[k]], d);
            }
            if (CC) for (int cl = 0; cl < C; ++cl) if (P.cluster_cc_off[c0 + cl] >= 0) cc_hist(cl, 0, d);
        }
    }
    if (R.final_assign)
        for (int f = 0; f < F; ++f) R.final_assign[(size_t)chain * P.n_frag + f0 + f] = MBSV_ST(f);
    if (R.final_logprob) R.final_logprob[sc] = (MODE == 0) ? logprob : full_logprob();
    const int n_naive_acc = cnt.get(C_NACC), n_sv_prop = cnt.get(C_SVPROP), n_sv_acc = cnt.get(C_SVACC);
    const long long n_reassign = (long long)n_naive_prop + (long long)(((unsigned long long)(unsigned)cnt.get(C_RHI) << 32) | (unsigned)cnt.get(C_RLO));
    if (R.counters) {
        long long* k = R.counters + sc * 5;
        k[0] = n_naive_prop; k[1] = n_naive_acc; k[2] = n_sv_prop; k[3] = n_sv_acc; k[4] = n_reassign;
    }
    if (R.rng_state) R.rng_state[sc] = rng.state();
    if (R.error) R.error[sc] = err;
    if (R.totals) warp_add_totals(R.totals, lane, n_naive_prop, n_naive_acc, n_sv_prop, n_sv_acc, n_reassign, err);
}

}  // namespace mbsv
