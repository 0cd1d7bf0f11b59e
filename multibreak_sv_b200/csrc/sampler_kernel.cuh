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

#ifndef MBSV_MIN_BLOCKS
#define MBSV_MIN_BLOCKS 5
#endif
// Launch bounds of the global-workspace class (large subproblems).  Measured on B200: 8 blocks/SM (64 registers, spills)
// is slower than 5 (96 registers) both alone (289 vs 267 ms) and next to the other classes.
// Read-only evaluation of Naive proposals in FAST mode (1) instead of tentative application + undo (0).  Measured on B200
// (5M-fragment mixed batch, 64 chains, 2000 iterations, classes run back to back; profiles/r02_readonly_eval_ab.json):
// SLOWER in every class of this kernel — shared-memory classes 28.8 -> 55.7 ms (128 rows), 60.4 -> 92.1 ms (48 rows), global
// workspace 198.8 -> 235.7 ms: the eight-entry register lists push the 96-register kernel to ~1 KB of spills per thread,
// which costs more than the undo pass saves.  Kept as a build option; the default stays do/undo.
#ifndef MBSV_READONLY_EVAL
#define MBSV_READONLY_EVAL 0
#endif
#ifndef MBSV_MIN_BLOCKS_GLOBAL
#define MBSV_MIN_BLOCKS_GLOBAL 5
#endif

// CC = the batch contains connected-component clusters (GASV localization -1): their support is the sorted
// multiset of a greedy set cover (MultiBreakSVAssignmentMatrix.java:337-409) kept per chain next to the counts.
template <int MODE, bool TRACE, bool SMEM, int MAXE, bool CC = false, bool PACKED = false>
__global__ void __launch_bounds__(128, SMEM ? MBSV_MIN_BLOCKS : MBSV_MIN_BLOCKS_GLOBAL) mbsv_chain_kernel(const __grid_constant__ DevProblem P,
                                                         const __grid_constant__ DevRun R) {
    extern __shared__ int smem_tile[];
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
    int* st = SMEM ? (smem_tile + (size_t)wic * R.rows_per_warp * 32 + lane)
                   : (R.workspace + (size_t)R.warp_row_off[warp - R.global_warp0] * st_stride + (PACKED ? 0 : lane));

#ifdef MBSV_BOUNDS_CHECK
    const int st_rows = !alive ? 0 : SMEM ? min(R.sub_W[sub], R.rows_per_warp) : R.sub_W[sub];
#endif
    const int f0 = P.sub_frag_ptr[sub], F = alive ? P.sub_frag_ptr[sub + 1] - f0 : 0;
    const int c0 = P.sub_cluster_ptr[sub], C = alive ? P.sub_cluster_ptr[sub + 1] - c0 : 0;
    const int E = P.n_exp;
    const int sv0 = P.sub_sv_ptr[sub], nsv = P.sub_sv_ptr[sub + 1] - sv0;
    const int Tstride = P.max_support + 1;
    const long long sc = (long long)sub * R.n_chains + chain;
    const int N = alive ? R.num_iters : 0;
    const double logF = P.sub_logF[sub];
    const double neglogF = -logF;
    const double perr = R.perr, logperr = R.logperr;
    const double dF = (double)F;
    const unsigned long long perr_thr = prob_threshold53(perr);
    constexpr unsigned long long NAIVE_THR = 0x1CCCCCCCCCCCCDULL;   // 0.9 * 2^53 exactly: nextDouble() < 0.9

    JavaRandom rng;
    rng.seed(R.seed + chain);

    // Memo-eligible subproblems (mbsv_run_params.memo_states) use REPLAY arithmetic; the table-driven version lives in
    // memo_kernel.cuh, a lane that lands here (chain count not a multiple of 32) re-sums the posterior per proposal.
    const int nst = (alive && R.sub_memo) ? R.sub_memo[sub] : 0;
    const bool use_full = (MODE == 0) || (nst > 0);

    long long Etot[MAXE], Ltot[MAXE];
    double LB[MAXE];
#pragma unroll
    for (int x = 0; x < MAXE; ++x) { Etot[x] = 0; Ltot[x] = 0; LB[x] = 0.0; }
    int nerr = 0;
    int err = 0;
    int n_naive_prop = 0, n_naive_acc = 0, n_sv_prop = 0, n_sv_acc = 0;   // <= num_iters each
    long long n_reassign = 0;

    auto wrapi = [&](long long t) -> long long { return R.java_int_wrap ? (long long)(int)(unsigned int)(unsigned long long)t : t; };
    auto lb_of = [&](int x, long long e, long long l) -> double {
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
            for (int j = 0; j < ne; ++j) {
                const int slot = P.aln_esp_slot[rec.z + j];
                if (slot < 0) continue;
                const int n = MBSV_ST(F + slot);
                dcov = dcov + (Tx[n - 1] - Tx[n]);
                MBSV_ST(F + slot) = n - 1;
            }
#pragma unroll
            for (int y = 0; y < MAXE; ++y) if (y == x) { Etot[y] -= rec.x; Ltot[y] -= rec.y; }
            touched |= 1u << x;
        } else {
            derr = derr - logperr;
            --nerr;
        }
        if (to != -1) {
            const int4 rec = P.aln_rec[a0 + to];
            const int x = rec.w & 0xff, ne = rec.w >> 8;
            const double* Tx = P.T + (size_t)x * Tstride;
            for (int j = 0; j < ne; ++j) {
                const int slot = P.aln_esp_slot[rec.z + j];
                if (slot < 0) continue;
                const int n = MBSV_ST(F + slot);
                dcov = dcov + (Tx[n + 1] - Tx[n]);
                MBSV_ST(F + slot) = n + 1;
            }
#pragma unroll
            for (int y = 0; y < MAXE; ++y) if (y == x) { Etot[y] += rec.x; Ltot[y] += rec.y; }
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
        for (int x = 0; x < MAXE; ++x) if (x < E) lp += lb_of(x, Etot[x], Ltot[x]);
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
                unsigned int* h = R.cluster_hist + P.cluster_hist_ptr[c0 + slot / E];
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
                unsigned int* h = R.cluster_hist + P.cluster_hist_ptr[c0 + slot / E];
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
                    if (SMEM) MBSV_ST(F + slot) += 1; else atomicAdd(&MBSV_ST(F + slot), 1);
                }
#pragma unroll
                for (int y = 0; y < MAXE; ++y) if (y == x) { Etot[y] += rec[k].x; Ltot[y] += rec[k].y; }
            }
        }
#pragma unroll
        for (int x = 0; x < MAXE; ++x) if (x < E) LB[x] = lb_of(x, Etot[x], Ltot[x]);
        if (has_cc)
            for (int cl = 0; cl < C; ++cl) if (P.cluster_cc_off[c0 + cl] >= 0) cc_recompute(cl);
    }
    double logprob = full_logprob();
    if (alive && R.initial_logprob) R.initial_logprob[sc] = logprob;

    // ---------------- iterations (MultiBreakSV.java:796-837, 1101-1182) ----------------
    const int q_base = P.sv_frag_ptr[sv0];   // only dereferenced when nsv > 0
    const double neglogNsv = P.sub_neglogNsv[sub];
    int iter = 0;
    bool done = false;
    for (;;) {
        // Lazy chain: with probability 1/2 stay (:1106-1113).  Lazy iterations only advance the RNG (tallies
        // are difference tallies), so each lane skips them in a tight loop; the warp-wide vote below is the
        // convergence point that brings all 32 lanes back together for the proposal code.
        double r = 0.0;
        bool live = false;
        if (TRACE) {
            if (!done) {
                while (iter < N) {
                    r = rng.nextDouble();
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
        if (__all_sync(0xffffffffu, done)) break;
        if (!live) continue;
        const int t = iter;
        // Move-type coin (:1115-1118): its value cannot matter when there is no AlterSV candidate; the fragment pick
        // (:1400) cannot matter for a one-fragment problem.  The draws are still consumed.
        bool naive = true;
        if (nsv == 0) rng.skip<2>(); else naive = rng.bits53() < NAIVE_THR;
        // A proposal is a list of single-fragment moves: one for Naive, the cluster's single-alignment
        // fragments for AlterSV.  Both kinds run through the same code below.
        int f = -1, a0 = 0, prev = -1, now = -1, svk = -1, q0 = 0, nmoves = 1;
        double Alogq, Anewlogq;
        double u = 0.0;
        if (naive && F == 1) rng.skip<2>(); else u = rng.nextDouble();
        if (naive) {
            // ---- naiveMove (:1184-1273) ----
            ++n_naive_prop; ++n_reassign;
            f = (int)(u * dF);
            a0 = P.frag_aln_ptr[f0 + f];
            const int n = P.frag_aln_ptr[f0 + f + 1] - a0;
            prev = MBSV_ST(f);
            naive_pick(rng, n, prev, neglogF, perr_thr, logperr, R.log1mperr, P.logint, P.invint, now, Alogq, Anewlogq);
            if (now == prev) { err = -5; done = true; continue; }   // Error at :1122-1123
        } else {
            // ---- svMove (:1291-1343) ----
            ++n_sv_prop;
            svk = (int)(u * (double)nsv);
            n_reassign += P.sv_numchanged[sv0 + svk];
            q0 = P.sv_frag_ptr[sv0 + svk];
            nmoves = P.sv_frag_ptr[sv0 + svk + 1] - q0;
            Alogq = neglogNsv;
            Anewlogq = neglogNsv;
        }
        // ---- evaluate the proposal ----
        double dcov = 0.0, derr = 0.0;
        unsigned touched = 0;
        bool written = false;
        double newLB[MAXE];
        double newlogprob, arg;
        // Read-only evaluation (FAST, Naive, at most four ESPs per alignment, no connected component): the Delta is computed
        // from the counts as they are; nothing is written unless the move is accepted, so a rejected proposal (four in five)
        // costs no state writes and no undo pass.  ro_sl = counted slots of the alignment left (0..3) and entered (4..7),
        // ro_cn = the count each entry sees: the stored count plus the earlier entries of this proposal on the same slot.
        bool ro = false;
        int ro_sl[8], ro_cn[8];
        long long roE[MAXE], roL[MAXE];
        if (MBSV_READONLY_EVAL && MODE == 1 && !TRACE && naive && !use_full && !has_cc) {
            int4 recF = make_int4(0, 0, 0, 0), recT = make_int4(0, 0, 0, 0);
            if (prev != -1) recF = P.aln_rec[a0 + prev];
            if (now != -1) recT = P.aln_rec[a0 + now];
            const int neF = recF.w >> 8, neT = recT.w >> 8;
            if (neF <= 4 && neT <= 4) {
                ro = true;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ro_sl[j] = j < neF ? P.aln_esp_slot[recF.z + j] : -1;
                    ro_sl[4 + j] = j < neT ? P.aln_esp_slot[recT.z + j] : -1;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) ro_cn[i] = ro_sl[i] >= 0 ? MBSV_ST(F + ro_sl[i]) : 0;
#pragma unroll
                for (int i = 1; i < 8; ++i)
#pragma unroll
                    for (int i2 = 0; i2 < i; ++i2)
                        if (ro_sl[i] >= 0 && ro_sl[i2] == ro_sl[i]) ro_cn[i] += i2 < 4 ? -1 : 1;
#pragma unroll
                for (int x = 0; x < MAXE; ++x) { roE[x] = Etot[x]; roL[x] = Ltot[x]; }
                if (prev != -1) {
                    const int x = recF.w & 0xff;
                    const double* Tx = P.T + (size_t)x * Tstride;
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (ro_sl[j] >= 0) dcov = dcov + (Tx[ro_cn[j] - 1] - Tx[ro_cn[j]]);
#pragma unroll
                    for (int y = 0; y < MAXE; ++y) if (y == x) { roE[y] -= recF.x; roL[y] -= recF.y; }
                    touched |= 1u << x;
                } else derr = derr - logperr;
                if (now != -1) {
                    const int x = recT.w & 0xff;
                    const double* Tx = P.T + (size_t)x * Tstride;
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (ro_sl[4 + j] >= 0) dcov = dcov + (Tx[ro_cn[4 + j] + 1] - Tx[ro_cn[4 + j]]);
#pragma unroll
                    for (int y = 0; y < MAXE; ++y) if (y == x) { roE[y] += recT.x; roL[y] += recT.y; }
                    touched |= 1u << x;
                } else derr = derr + logperr;
                double dlb = 0.0;
#pragma unroll
                for (int x = 0; x < MAXE; ++x) newLB[x] = LB[x];
                unsigned mask = touched & ((1u << E) - 1u);
                while (mask) {
                    const int x = __ffs(mask) - 1;
                    mask &= mask - 1;
                    long long ex = 0, lx = 0;
                    double old = 0.0;
#pragma unroll
                    for (int y = 0; y < MAXE; ++y) if (y == x) { ex = roE[y]; lx = roL[y]; old = LB[y]; }
                    const double nl = lb_of(x, ex, lx);
#pragma unroll
                    for (int y = 0; y < MAXE; ++y) if (y == x) newLB[y] = nl;
                    dlb = dlb + (nl - old);
                }
                const double delta = (dcov + derr) + dlb;
                newlogprob = logprob + delta;
                arg = (delta + Alogq) - Anewlogq;
            }
        }
        if (!ro) {
        // ---- tentative application ----
        for (int q = 0; q < nmoves; ++q) {
            int ga0 = a0, from = prev, to = now;
            if (!naive) {
                const int g = P.sv_frag[q0 + q] - f0;
                ga0 = P.frag_aln_ptr[f0 + g];
                from = MBSV_ST(g);
                to = (from == -1) ? 0 : -1;
            }
            move_counts(ga0, from, to, dcov, derr, touched);
        }
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
            if (err) { done = true; continue; }   // set cover failed to cover: the reference throws Error
        }
        // ---- acceptance ratio (:1146) ----
        if (use_full) {
            // full re-sum in the reference's order; the assignment change only enters through nerr, counts, totals
#pragma unroll
            for (int x = 0; x < MAXE; ++x) newLB[x] = LB[x];
            newlogprob = full_logprob();
            arg = (newlogprob + Alogq) - (logprob + Anewlogq);
        } else {
            double dlb = 0.0;
#pragma unroll
            for (int x = 0; x < MAXE; ++x) newLB[x] = LB[x];
            // Each lane walks ITS touched experiments in ascending order (same summation order as before), so lanes
            // whose proposals touch different platforms evaluate logBinomial together instead of one platform at a time.
            unsigned mask = touched & ((1u << E) - 1u);
            while (mask) {
                const int x = __ffs(mask) - 1;
                mask &= mask - 1;
                long long ex = 0, lx = 0;
                double old = 0.0;
#pragma unroll
                for (int y = 0; y < MAXE; ++y) if (y == x) { ex = Etot[y]; lx = Ltot[y]; old = LB[y]; }
                const double nl = lb_of(x, ex, lx);
#pragma unroll
                for (int y = 0; y < MAXE; ++y) if (y == x) newLB[y] = nl;
                dlb = dlb + (nl - old);
            }
            const double delta = (dcov + derr) + dlb;
            newlogprob = logprob + delta;
            arg = (delta + Alogq) - Anewlogq;
        }
        }
        const double e = exp_dev(arg);
        const double ratio = (e != e) ? e : (1.0 < e ? 1.0 : e);   // Math.min(1, e)
        r = rng.nextDouble();
        const bool accept = ratio > r;
        if (accept && ro) {
            // apply the accepted move: counts (in entry order, so a slot hit twice ends at its final value), totals,
            // difference tallies (value cn leaves: +d, value cn +- 1 enters: -d), assignment
            logprob = newlogprob;
            ++n_naive_acc;
#pragma unroll
            for (int x = 0; x < MAXE; ++x) { LB[x] = newLB[x]; Etot[x] = roE[x]; Ltot[x] = roL[x]; }
            nerr += (now == -1 ? 1 : 0) - (prev == -1 ? 1 : 0);
            const unsigned d = t > win_base ? (unsigned)(t - win_base) : 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (ro_sl[i] < 0) continue;
                const int cn2 = ro_cn[i] + (i < 4 ? -1 : 1);
                MBSV_ST(F + ro_sl[i]) = cn2;
                if (d) {
                    unsigned int* h = R.cluster_hist + P.cluster_hist_ptr[c0 + ro_sl[i] / E];
                    if (ro_cn[i] > 0) atomicAdd(&h[ro_cn[i]], d);
                    if (cn2 > 0) atomicAdd(&h[cn2], 0u - d);
                }
            }
            tally_frag(f, a0, prev, now, t);
            MBSV_ST(f) = now;
        } else if (ro) {
            // rejected: nothing was written
        } else if (accept) {
            logprob = newlogprob;
#pragma unroll
            for (int x = 0; x < MAXE; ++x) LB[x] = newLB[x];
            if (naive) ++n_naive_acc; else ++n_sv_acc;
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
        } else {
            // revert the counts (and, on the CC path, the tentative assignments and the saved multisets)
            double d1 = 0.0, d2 = 0.0;
            unsigned tt = 0;
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
            for (int q = nmoves - 1; q >= 0; --q) {
                int g = f, ga0 = a0, from = prev, to = now;
                if (!naive) {
                    g = P.sv_frag[q0 + q] - f0;
                    ga0 = P.frag_aln_ptr[f0 + g];
                    const int cur = MBSV_ST(g);
                    from = written ? ((cur == -1) ? 0 : -1) : cur;
                    to = (from == -1) ? 0 : -1;
                }
                move_counts(ga0, to, from, d1, d2, tt);
                if (written) MBSV_ST(g) = from;
            }
        }
        if (TRACE) {
            R.trace_logprob[sc * N + iter] = logprob;
            R.trace_move_frag[sc * N + iter] = accept ? (naive ? f0 + f : -(2 + svk)) : -1;
            R.trace_move_assign[sc * N + iter] = (accept && naive) ? now : -1;
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
                for (int k = 0; k < EB; ++k) hp[k] = n[k] > 0 ? P.cluster_hist_ptr[c0 + (sb + k) / E] : 0;
#pragma unroll
                for (int k = 0; k < EB; ++k) if (n[k] > 0) atomicAdd(&R.cluster_hist[hp[k] + n[k]], d);
            }
            if (CC) for (int cl = 0; cl < C; ++cl) if (P.cluster_cc_off[c0 + cl] >= 0) cc_hist(cl, 0, d);
        }
    }
    if (R.final_assign)
        for (int f = 0; f < F; ++f) R.final_assign[(size_t)chain * P.n_frag + f0 + f] = MBSV_ST(f);
    if (R.final_logprob) R.final_logprob[sc] = (MODE == 0) ? logprob : full_logprob();
    if (R.counters) {
        long long* k = R.counters + sc * 5;
        k[0] = n_naive_prop; k[1] = n_naive_acc; k[2] = n_sv_prop; k[3] = n_sv_acc; k[4] = n_reassign;
    }
    if (R.rng_state) R.rng_state[sc] = rng.state();
    if (R.error) R.error[sc] = err;
    if (R.totals) warp_add_totals(R.totals, lane, n_naive_prop, n_naive_acc, n_sv_prop, n_sv_acc, n_reassign, err);
}

}  // namespace mbsv
