// Memoised sampler kernel: subproblems whose joint state space prod(n_f + 1) is small (mbsv_run_params.
// memo_states) and that contain no connected component.
//
// The reference itself memoises the log-posterior of every state it has visited, keyed by the state string
// (MultiBreakSV.java:58,1133-1137; the value is the same as a recomputation).  Here every warp holds
// 32 chains of ONE subproblem (the planner only routes chain counts that are multiples of 32 here), builds the table
// log-posterior[state] once in shared memory — computeLogProb in the reference's summation order
// (MultiBreakSVAssignmentMatrix.java:130-246), the entries spread over the lanes — and then a proposal costs the RNG draws,
// two table reads, one exp and the acceptance test.  Counts are only maintained for the support tallies and follow
// accepted moves; nothing is applied tentatively, so there is nothing to revert.
//
// Arithmetic: (LP[new] + Alogq) - (LP[old] + Anewlogq), i.e. exactly MultiBreakSV.java:1146 — both modes
// (REPLAY_EXACT and FAST) are served by this kernel with identical bits.
#pragma once
#include "sampler_common.cuh"

namespace mbsv {

#ifndef MBSV_MEMO_MIN_BLOCKS
#define MBSV_MEMO_MIN_BLOCKS 8
#endif

// Fills the device-memory tables: one thread per joint state of every subproblem whose table does not fit the
// shared-memory classes.  state_ptr[k] is the first global state of the k-th such subproblem (ascending), so a thread
// finds its (subproblem, state) by binary search; the counts of the state live in a [slot][thread] shared-memory
// scratch (one column per thread, bank-conflict free).
__global__ void __launch_bounds__(128) mbsv_memo_build_kernel(const __grid_constant__ DevProblem P,
                                                              const __grid_constant__ DevRun R, const int n_tab,
                                                              const long long* __restrict__ state_ptr,
                                                              const int* __restrict__ tab_sub, const int scratch_rows) {
    extern __shared__ int smem_tile[];
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= state_ptr[n_tab]) return;
    int lo = 0, hi = n_tab - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (state_ptr[mid] <= g) lo = mid; else hi = mid - 1; }
    const int sub = tab_sub[lo];
    const int idx = (int)(g - state_ptr[lo]);
    int* cnt = smem_tile + threadIdx.x;                 // cnt[slot * 128]
    const int f0 = P.sub_frag_ptr[sub], F = P.sub_frag_ptr[sub + 1] - f0;
    const int c0 = P.sub_cluster_ptr[sub], C = P.sub_cluster_ptr[sub + 1] - c0;
    const int E = P.n_exp, Tstride = P.max_support + 1;
    if (C * E > scratch_rows) return;                   // cannot happen: the host sizes the scratch
    for (int i = 0; i < C * E; ++i) cnt[i * 128] = 0;
    long long Et[MBSV_MAX_EXP], Lt[MBSV_MAX_EXP];
    for (int x = 0; x < MBSV_MAX_EXP; ++x) { Et[x] = 0; Lt[x] = 0; }
    int nerr = 0, rem = idx;
    for (int f = 0; f < F; ++f) {
        const int a0 = P.frag_aln_ptr[f0 + f];
        const int n1 = P.frag_aln_ptr[f0 + f + 1] - a0 + 1;
        const int a = rem % n1 - 1;
        rem /= n1;
        if (a == -1) { ++nerr; continue; }
        const int4 rec = P.aln_rec[a0 + a];
        const int x = rec.w & 0xff, ne = rec.w >> 8;
        for (int j = 0; j < ne; ++j) {
            const int slot = P.aln_esp_slot[rec.z + j];
            if (slot >= 0) cnt[slot * 128] += 1;
        }
        Et[x] += rec.x; Lt[x] += rec.y;
    }
    double lp = 0;
    for (int i = 0; i < C; ++i) {
        const int cl = P.supports_order_local[c0 + i];
        double lpc = 0.0;
        for (int x = 0; x < E; ++x) {
            const int n = cnt[(cl * E + x) * 128];
            if (n > 0) lpc = lpc + P.T[(size_t)x * Tstride + n];
        }
        lp += lpc;
    }
    lp += 0.0;
    lp += (double)nerr * R.logperr;
    for (int x = 0; x < E; ++x) {
        long long e = Et[x], l = Lt[x];
        if (R.java_int_wrap) { e = (long long)(int)(unsigned)(unsigned long long)e; l = (long long)(int)(unsigned)(unsigned long long)l; }
        lp += log_binomial_dev(e, l, P.exp_pseq[x], P.exp_logp[x], P.exp_log1mp[x], P.log2pi, P.logint, P.maxn);
    }
    R.memo_global[R.sub_tab_off[sub] + idx] = lp;
}

template <bool TRACE, bool MIXED>
__global__ void __launch_bounds__(128, MBSV_MEMO_MIN_BLOCKS) mbsv_memo_kernel(const __grid_constant__ DevProblem P,
                                                                              const __grid_constant__ DevRun R) {
    extern __shared__ int smem_tile[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int warp = R.warp_begin + blockIdx.x * 4 + wic;
    if (warp >= R.warp_end) return;
    const long long item_raw = R.item_base + (long long)(warp - R.warp_base) * R.lanes_per_warp + lane;
    const bool alive = lane < R.lanes_per_warp && item_raw < R.n_items;
    // idle lanes mirror the warp's first item, so that every per-subproblem quantity stays warp-uniform
    const long long item = alive ? item_raw : R.item_base + (long long)(warp - R.warp_base) * R.lanes_per_warp;
    // warp-uniform unless R.memo_mixed (then lanes_per_warp need not divide n_chains and nothing below relies on it)
    const int sub = R.sub_order[item / R.n_chains];
    constexpr bool mixed = MIXED;     // (DevRun.memo_mixed says the same; compile-time here so full warps pay nothing)
    const int chain = (int)(item % R.n_chains);
    int* st = smem_tile + (size_t)wic * R.rows_per_warp * 32 + lane;
    constexpr int st_stride = 32;

#ifdef MBSV_BOUNDS_CHECK
    const int st_rows = min(R.sub_W[sub] + (MIXED ? P.sub_frag_ptr[sub + 1] - P.sub_frag_ptr[sub] : 0), R.rows_per_warp);
#endif
    const int f0 = P.sub_frag_ptr[sub], F = P.sub_frag_ptr[sub + 1] - f0;
    const int c0 = P.sub_cluster_ptr[sub], C = P.sub_cluster_ptr[sub + 1] - c0;
    const int E = P.n_exp;
    const int sv0 = P.sub_sv_ptr[sub], nsv = P.sub_sv_ptr[sub + 1] - sv0;
    const int Tstride = P.max_support + 1;
    const long long sc = (long long)sub * R.n_chains + chain;
    const int N = alive ? R.num_iters : 0;
    const int W = R.sub_W[sub];
    const int nst = R.sub_memo[sub];
    // after the state rows: fragment strides, then the table — in shared memory when it fits the size classes,
    // otherwise in a global buffer (one region per subproblem; both warps of a subproblem fill it with the same
    // values and each reads only after its own complete build)
    // (mixed warps: the strides sit in the lane's own column, rows W .. W+F-1, and every table is in the global buffer)
    int* strides = mixed ? st + (size_t)W * 32 : smem_tile + ((size_t)wic * R.rows_per_warp + W) * 32;
    constexpr int sstr = MIXED ? 32 : 1;
    const long long tab_off = R.sub_tab_off[sub];
    double* tab = tab_off >= 0 ? R.memo_global + tab_off : reinterpret_cast<double*>(strides + ((F + 1) & ~1));
    const double logperr = R.logperr;
    const int win_base = R.win_base;
    const double neglogF = -P.sub_logF[sub];
    // One-fragment subproblems without AlterSV candidates (the bulk of a power-law batch): the whole acceptance
    // ratio min(1, exp(..)) of every (previous, proposed) pair is tabulated too, so a proposal needs no exp.
    const bool f1 = (F == 1) && (nsv == 0) && (tab_off < 0);
    // Small subproblems of several fragments without AlterSV candidates tabulate the ratios of their single-fragment moves the
    // same way (every proposal of theirs is one): rtab[state][state'] for the pairs that differ in exactly one fragment.
    const bool rt = !f1 && !MIXED && tab_off < 0 && nsv == 0 && nst <= MEMO_RT_MAX_STATES &&
                    W + (((F + 1) & ~1) + 2 * nst + 2 * nst * nst + 31) / 32 <= MEMO_SMEM_TABLE_ROWS;
    double* rtab = tab + nst;
#ifdef MBSV_BOUNDS_CHECK
    {   // strides, table and ratio table must end inside this warp's rows of the tile
        const int* end = mixed ? smem_tile : tab_off >= 0 ? strides + ((F + 1) & ~1) : reinterpret_cast<const int*>(rtab + ((f1 || rt) ? nst * nst : 0));
        if ((mixed && tab_off < 0) || end > smem_tile + (size_t)(wic + 1) * R.rows_per_warp * 32) {
            printf("memo tables of subproblem %d overrun the %d-row tile\n", sub, R.rows_per_warp);
            __trap();
        }
    }
#endif

    // ---------------- table: log-posterior of every joint state (fragment 0 = least significant digit) ----------------
    {
        const int nalive = __popc(__ballot_sync(0xffffffffu, alive));
        if (alive) {
            // tables in device memory were filled by mbsv_memo_build_kernel before this launch
            for (int idx = lane; idx < (tab_off >= 0 ? 0 : nst); idx += nalive) {
                for (int i = F; i < F + C * E; ++i) MBSV_ST(i) = 0;
                long long Et[MBSV_MAX_EXP], Lt[MBSV_MAX_EXP];
                for (int x = 0; x < MBSV_MAX_EXP; ++x) { Et[x] = 0; Lt[x] = 0; }
                int nerr = 0, rem = idx;
                for (int f = 0; f < F; ++f) {
                    const int a0 = P.frag_aln_ptr[f0 + f];
                    const int n1 = P.frag_aln_ptr[f0 + f + 1] - a0 + 1;
                    const int a = rem % n1 - 1;
                    rem /= n1;
                    if (a == -1) { ++nerr; continue; }
                    const int4 rec = P.aln_rec[a0 + a];
                    const int x = rec.w & 0xff, ne = rec.w >> 8;
                    for (int j = 0; j < ne; ++j) {
                        const int slot = P.aln_esp_slot[rec.z + j];
                        if (slot >= 0) MBSV_ST(F + slot) += 1;
                    }
                    Et[x] += rec.x; Lt[x] += rec.y;
                }
                double lp = 0;
                for (int i = 0; i < C; ++i) {
                    const int cl = P.supports_order_local[c0 + i];
                    double lpc = 0.0;
                    for (int x = 0; x < E; ++x) {
                        const int n = MBSV_ST(F + cl * E + x);
                        if (n > 0) lpc = lpc + P.T[(size_t)x * Tstride + n];
                    }
                    lp += lpc;
                }
                lp += 0.0;
                lp += (double)nerr * logperr;
                for (int x = 0; x < E; ++x) {
                    long long e = Et[x], l = Lt[x];
                    if (R.java_int_wrap) { e = (long long)(int)(unsigned)(unsigned long long)e; l = (long long)(int)(unsigned)(unsigned long long)l; }
                    lp += log_binomial_dev(e, l, P.exp_pseq[x], P.exp_logp[x], P.exp_log1mp[x], P.log2pi, P.logint, P.maxn);
                }
                tab[idx] = lp;
            }
            if (lane == 0 || mixed) {
                int sprod = 1;
                for (int f = 0; f < F; ++f) { strides[f * sstr] = sprod; sprod *= P.frag_aln_ptr[f0 + f + 1] - P.frag_aln_ptr[f0 + f] + 1; }
            }
        }
        __syncwarp();
        if (f1 && alive) {
            const int n = nst - 1;                      // alignments of the only fragment
            for (int idx = lane; idx < nst * nst; idx += nalive) {
                const int pv = idx / nst - 1, nw = idx % nst - 1;
                double ratio = 0.0;
                if (pv != nw) {
                    double Alogq, Anewlogq;
                    naive_logq(n, pv, nw, neglogF, logperr, R.log1mperr, P.logint, Alogq, Anewlogq);
                    const double e = exp_dev((tab[nw + 1] + Alogq) - (tab[pv + 1] + Anewlogq));   // :1146
                    ratio = (e != e) ? e : (1.0 < e ? 1.0 : e);
                }
                rtab[idx] = ratio;
            }
        }
        if (rt && alive) {
            for (int idx = lane; idx < nst * nst; idx += nalive) {
                const int s1 = idx / nst, s2 = idx % nst;
                int nd = 0, n = 0, pv = 0, nw = 0;
                for (int f = 0; f < F; ++f) {
                    const int nf = P.frag_aln_ptr[f0 + f + 1] - P.frag_aln_ptr[f0 + f];
                    const int d1 = (s1 / strides[f]) % (nf + 1), d2 = (s2 / strides[f]) % (nf + 1);
                    if (d1 != d2) { ++nd; n = nf; pv = d1 - 1; nw = d2 - 1; }
                }
                double ratio = 0.0;
                if (nd == 1) {
                    double Alogq, Anewlogq;
                    naive_logq(n, pv, nw, neglogF, logperr, R.log1mperr, P.logint, Alogq, Anewlogq);
                    const double e = exp_dev((tab[s2] + Alogq) - (tab[s1] + Anewlogq));   // :1146
                    ratio = (e != e) ? e : (1.0 < e ? 1.0 : e);
                }
                rtab[idx] = ratio;
            }
        }
        __syncwarp();
    }

    // ---------------- initializeMatrix (MultiBreakSV.java:990-1099) ----------------
    JavaRandom rng;
    rng.seed(R.seed + chain);
    int sidx = 0;
    if (alive) {
        for (int i = 0; i < W; ++i) MBSV_ST(i) = 0;
        for (int f = 0; f < F; ++f) {
            const int a0 = P.frag_aln_ptr[f0 + f];
            const int n = P.frag_aln_ptr[f0 + f + 1] - a0;
            int a;
            if (R.init_assign) a = R.init_assign[f0 + f];
            else {
                if (rng.nextDouble() < R.perr) a = -1;
                else a = rng.nextInt(n);
            }
            MBSV_ST(f) = a;
            sidx += (a + 1) * strides[f * sstr];
            if (R.initial_assign) R.initial_assign[(size_t)chain * P.n_frag + f0 + f] = a;
        }
    }
    double logprob = alive ? tab[sidx] : 0.0;
    if (alive && R.initial_logprob) R.initial_logprob[sc] = logprob;

    // The selected-ESP counts only feed the support tallies, so they are rebuilt from the assignments when a chain first
    // needs them (its first accepted move inside the long-burn-in window, or the epilogue) and maintained from then on.
    bool counts_live = false;
    auto rebuild_counts = [&]() {
        for (int i = F; i < F + C * E; ++i) MBSV_ST(i) = 0;
        for (int f = 0; f < F; ++f) {
            const int a = MBSV_ST(f);
            if (a == -1) continue;
            const int4 rec = P.aln_rec[P.frag_aln_ptr[f0 + f] + a];
            const int ne = rec.w >> 8;
            for (int j = 0; j < ne; ++j) {
                const int slot = P.aln_esp_slot[rec.z + j];
                if (slot >= 0) MBSV_ST(F + slot) += 1;
            }
        }
        counts_live = true;
    };
    // Counts follow an accepted single-fragment move; inside the long-burn-in window every unit step also emits
    // the support-histogram difference tallies (value n leaves: +d, value n' enters: -d).
    auto apply_move = [&](int g, int a0, int from, int to, unsigned d) {
        if (from != -1) {
            const int4 rec = P.aln_rec[a0 + from];
            const int ne = rec.w >> 8;
            for (int j = 0; j < ne; ++j) {
                const int slot = P.aln_esp_slot[rec.z + j];
                if (slot < 0) continue;
                const int n = MBSV_ST(F + slot);
                if (d) {
                    unsigned int* h = R.cluster_hist + P.slot_hist_ptr[c0 * E + slot];
                    if (n > 0) atomicAdd(&h[n], d);
                    if (n - 1 > 0) atomicAdd(&h[n - 1], 0u - d);
                }
                MBSV_ST(F + slot) = n - 1;
            }
        }
        if (to != -1) {
            const int4 rec = P.aln_rec[a0 + to];
            const int ne = rec.w >> 8;
            for (int j = 0; j < ne; ++j) {
                const int slot = P.aln_esp_slot[rec.z + j];
                if (slot < 0) continue;
                const int n = MBSV_ST(F + slot);
                if (d) {
                    unsigned int* h = R.cluster_hist + P.slot_hist_ptr[c0 * E + slot];
                    if (n > 0) atomicAdd(&h[n], d);
                    atomicAdd(&h[n + 1], 0u - d);
                }
                MBSV_ST(F + slot) = n + 1;
            }
        }
        if (d) {
            if (from == -1) atomicAdd(&R.frag_err_count[f0 + g], d); else atomicAdd(&R.aln_count[a0 + from], d);
            if (to == -1) atomicAdd(&R.frag_err_count[f0 + g], 0u - d); else atomicAdd(&R.aln_count[a0 + to], 0u - d);
        }
        MBSV_ST(g) = to;
    };

    // ---------------- iterations (MultiBreakSV.java:796-837, 1101-1182) ----------------
    const double neglogNsv = P.sub_neglogNsv[sub];
    const unsigned long long perr_thr = prob_threshold53(R.perr);
    constexpr unsigned long long NAIVE_THR = 0x1CCCCCCCCCCCCDULL;   // 0.9 * 2^53: nextDouble() < 0.9
    int n_naive_prop = 0, n_naive_acc = 0, n_sv_prop = 0, n_sv_acc = 0;
    long long n_sv_reassign = 0;
    int err = 0;
    int iter = 0;
    bool done = false;
    for (;;) {
        bool live = false;
        if (TRACE) {
            if (!done) {
                while (iter < N) {
                    const double r0 = rng.nextDouble();
                    if (!(r0 < 0.5)) { live = true; break; }
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
        bool naive = true;
        int f = 0, a0 = 0, prev = -1, now = -1, svk = -1, q0 = 0, nmoves = 1;
        int newidx = sidx;
        double newlogprob, ratio;
        if (f1) {
            // move-type coin and fragment pick are consumed but cannot matter (no AlterSV candidate, one fragment)
            rng.skip<4>();
            ++n_naive_prop;
            a0 = P.frag_aln_ptr[f0];
            prev = sidx - 1;
            now = naive_draw(rng, nst - 1, prev, perr_thr, P.invint);
            if (now == prev) { err = -5; done = true; continue; }   // Error at :1122-1123
            newidx = now + 1;
            newlogprob = tab[newidx];
            ratio = rtab[sidx * nst + newidx];
        } else if (rt) {
            // no AlterSV candidate (the move-type coin cannot matter), several fragments; the ratio comes from the table
            rng.skip<2>();
            const double u = rng.nextDouble();
            ++n_naive_prop;
            f = (int)(u * (double)F);
            a0 = P.frag_aln_ptr[f0 + f];
            prev = MBSV_ST(f);
            now = naive_draw(rng, P.frag_aln_ptr[f0 + f + 1] - a0, prev, perr_thr, P.invint);
            if (now == prev) { err = -5; done = true; continue; }   // Error at :1122-1123
            newidx += (now - prev) * strides[f * sstr];
            newlogprob = tab[newidx];
            ratio = rtab[sidx * nst + newidx];
        } else {
            if (nsv == 0) rng.skip<2>(); else naive = rng.bits53() < NAIVE_THR;
            double u = 0.0;
            if (naive && F == 1) rng.skip<2>(); else u = rng.nextDouble();
            double Alogq, Anewlogq;
            if (naive) {
                ++n_naive_prop;
                f = (int)(u * (double)F);
                a0 = P.frag_aln_ptr[f0 + f];
                const int n = P.frag_aln_ptr[f0 + f + 1] - a0;
                prev = MBSV_ST(f);
                naive_pick(rng, n, prev, neglogF, perr_thr, logperr, R.log1mperr, P.logint, P.invint, now, Alogq, Anewlogq);
                if (now == prev) { err = -5; done = true; continue; }   // Error at :1122-1123
                newidx += (now - prev) * strides[f * sstr];
            } else {
                ++n_sv_prop;
                svk = (int)(u * (double)nsv);
                n_sv_reassign += P.sv_numchanged[sv0 + svk];
                q0 = P.sv_frag_ptr[sv0 + svk];
                nmoves = P.sv_frag_ptr[sv0 + svk + 1] - q0;
                Alogq = neglogNsv;
                Anewlogq = neglogNsv;
                for (int q = 0; q < nmoves; ++q) {
                    const int g = P.sv_frag[q0 + q] - f0;
                    const int from = MBSV_ST(g);
                    newidx += ((from == -1) ? 1 : -1) * strides[g * sstr];
                }
            }
            newlogprob = tab[newidx];
            const double arg = (newlogprob + Alogq) - (logprob + Anewlogq);      // :1146
            const double e = exp_dev(arg);
            ratio = (e != e) ? e : (1.0 < e ? 1.0 : e);                          // Math.min(1, e)
        }
        double r;
        r = rng.nextDouble();
        const bool accept = ratio > r;
        if (accept) {
            logprob = newlogprob;
            sidx = newidx;
            const unsigned d = t > win_base ? (unsigned)(t - win_base) : 0u;
            if (naive) ++n_naive_acc; else ++n_sv_acc;
            if (d == 0u && !counts_live) {
                // before the window only the assignments move
                if (naive) MBSV_ST(f) = now;
                else for (int q = 0; q < nmoves; ++q) {
                    const int g = P.sv_frag[q0 + q] - f0;
                    MBSV_ST(g) = (MBSV_ST(g) == -1) ? 0 : -1;
                }
            } else {
                if (!counts_live) rebuild_counts();
                if (naive) apply_move(f, a0, prev, now, d);
                else for (int q = 0; q < nmoves; ++q) {
                    const int g = P.sv_frag[q0 + q] - f0;
                    const int from = MBSV_ST(g);
                    apply_move(g, P.frag_aln_ptr[f0 + g], from, (from == -1) ? 0 : -1, d);
                }
            }
        }
        if (TRACE) {
            R.trace_logprob[sc * N + iter] = logprob;
            R.trace_move_frag[sc * N + iter] = accept ? (naive ? f0 + f : -(2 + svk)) : -1;
            R.trace_move_assign[sc * N + iter] = (accept && naive) ? now : -1;
        }
        ++iter;
    }

    // ---------------- epilogue ----------------
    if (!alive) return;
    if (err == 0) {
        const long long nwin = (long long)N - win_base;
        if (nwin > 0) {
            const unsigned d = (unsigned)nwin;
            if (!counts_live) rebuild_counts();
            for (int f = 0; f < F; ++f) {
                const int a = MBSV_ST(f);
                if (a == -1) atomicAdd(&R.frag_err_count[f0 + f], d);
                else atomicAdd(&R.aln_count[P.frag_aln_ptr[f0 + f] + a], d);
            }
            for (int cl = 0; cl < C; ++cl) {
                unsigned int* h = R.cluster_hist + P.cluster_hist_ptr[c0 + cl];
                for (int x = 0; x < E; ++x) {
                    const int n = MBSV_ST(F + cl * E + x);
                    if (n > 0) atomicAdd(&h[n], d);
                }
            }
        }
    }
    if (R.final_assign)
        for (int f = 0; f < F; ++f) R.final_assign[(size_t)chain * P.n_frag + f0 + f] = MBSV_ST(f);
    if (R.final_logprob) R.final_logprob[sc] = logprob;
    const long long n_reassign = (long long)n_naive_prop + n_sv_reassign;
    if (R.counters) {
        long long* k = R.counters + sc * 5;
        k[0] = n_naive_prop; k[1] = n_naive_acc; k[2] = n_sv_prop; k[3] = n_sv_acc; k[4] = n_reassign;
    }
    if (R.rng_state) R.rng_state[sc] = rng.state();
    if (R.error) R.error[sc] = err;
    if (R.totals) warp_add_totals(R.totals, lane, n_naive_prop, n_naive_acc, n_sv_prop, n_sv_acc, n_reassign, err);
}

}  // namespace mbsv
