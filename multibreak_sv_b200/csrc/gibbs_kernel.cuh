// Heat-bath (Gibbs) sampler, SURVEY.md section 8 row f-4: NOT the reference's chain.  The reference is Metropolis-Hastings
// (MultiBreakSV.java:1101-1182) and only that replays its trajectory; this kernel samples the SAME posterior
// (MultiBreakSVAssignmentMatrix.computeLogProb, :130-246) with exact conditionals, for the statistical tier: random-scan
// Gibbs, one fragment per iteration redrawn from P(assignment | everything else).  No proposal is ever rejected, so it
// mixes faster per iteration than the lazy MH chain (half of whose iterations do nothing and ~83 % of whose proposals are
// rejected on the 5M config); it does not reproduce quirk Q1 of naiveMove (the asymmetric proposal), which is why the
// two samplers agree on the posterior only where the reference's own chain is unbiased.
//
// Mapping: one GROUP of G lanes (G = 4, 8 or 32, a template parameter chosen per size class) per (subproblem, chain).
// The chain's state (assignments, selected-ESP counts) lives in the group's shared memory; the candidate alignments of
// the chosen fragment are spread over the group's lanes: every lane evaluates the log-posterior change of "its" option, a
// group max / sum normalises, and the categorical draw is an inclusive __shfl_up prefix scan followed by a __ballot for
// the first lane whose cumulative weight passes u.  All lanes of a group carry the same RNG state and the same
// per-experiment totals (uniform registers), so nothing needs broadcasting; the groups of a warp run independent chains
// and only ever synchronise within their own lane mask.  Small subproblems (two or three candidates) use G = 4.
// A subproblem whose state does not fit the shared-memory classes runs with GLOBAL = true: one chain per warp (G = 32), the
// state a contiguous column of the global workspace (as kern_packed.cu does for the MH chain), only the option weights in
// shared memory; same arithmetic, same draws, bit-identical results.
// Connected-component clusters (GASV localization -1): their support is the multiset of a greedy set cover over the assigned
// leaves (MultiBreakSVAssignmentMatrix.java:337-409), a function of the assignments, not of counts.  Lane 0 of the group
// evaluates it: for every option of the redrawn fragment, each component the option's alignment touches is recomputed with the
// fragment unassigned and with the option, the difference of the two log terms joins the option's weight; after the draw the
// touched components are recomputed for the new state and their multisets' difference tallies are emitted, as the MH kernel
// does (sampler_kernel.cuh).  State layout of the components = the MH kernel's (stamp, current multisets, saved multisets,
// set-cover scratch behind the counts).
#pragma once
#include "sampler_common.cuh"

namespace mbsv {

template <int G> __device__ __forceinline__ double group_max(unsigned mask, double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) { const double w = __shfl_xor_sync(mask, v, o); v = w > v ? w : v; }
    return v;
}
template <int G> __device__ __forceinline__ double group_sum(unsigned mask, double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

template <int G, bool GLOBAL = false>
__global__ void __launch_bounds__(128) mbsv_gibbs_kernel(const __grid_constant__ DevProblem P, const __grid_constant__ DevRun R) {
    extern __shared__ int smem_tile[];
    constexpr int GPW = 32 / G;                                    // groups (chains) per warp
    const int wlane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int lane = wlane % G, grp = wlane / G;                   // `lane` = position inside the group from here on
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
    const int warp = R.warp_begin + blockIdx.x * 4 + wic;
    if (warp >= R.warp_end) return;
    const long long item = R.item_base + (long long)(warp - R.warp_base) * GPW + grp;
    if (item >= R.n_items) return;                                 // (the other groups never name these lanes in a mask)
    const int sub = R.sub_order[item / R.n_chains];
    const int chain = (int)(item % R.n_chains);
    // [F] assignments, [C*E] counts, then the option weights; GLOBAL: the state in the workspace, the weights in shared memory
    int* st = GLOBAL ? R.workspace + R.warp_row_off[warp - R.global_warp0] : smem_tile + ((size_t)wic * GPW + grp) * R.rows_per_warp;
    const int f0 = P.sub_frag_ptr[sub], F = P.sub_frag_ptr[sub + 1] - f0;
    const int c0 = P.sub_cluster_ptr[sub], C = P.sub_cluster_ptr[sub + 1] - c0;
    const int E = P.n_exp, Tstride = P.max_support + 1;
    double* wopt = GLOBAL ? reinterpret_cast<double*>(smem_tile + (size_t)wic * R.rows_per_warp)
                          : reinterpret_cast<double*>(st + ((R.sub_W[sub] + 1) & ~1));
    const long long sc = (long long)sub * R.n_chains + chain;
    const int N = R.num_iters, win_base = R.win_base;
    const double logperr = R.logperr;

    JavaRandom rng;
    rng.seed(R.seed + chain);
    long long Et[MBSV_MAX_EXP], Lt[MBSV_MAX_EXP];
    double LB[MBSV_MAX_EXP];
#pragma unroll
    for (int x = 0; x < MBSV_MAX_EXP; ++x) { Et[x] = 0; Lt[x] = 0; LB[x] = 0.0; }
    int nerr = 0;
    auto lb_of = [&](int x, long long e, long long l) {
        return log_binomial_dev(e, l, P.exp_pseq[x], P.exp_logp[x], P.exp_log1mp[x], P.log2pi, P.logint, P.maxn);
    };
    auto bump_totals = [&](int x, long long de, long long dl) {
#pragma unroll
        for (int y = 0; y < MBSV_MAX_EXP; ++y) if (y == x) { Et[y] += de; Lt[y] += dl; }
    };
    auto get_E = [&](int x) { long long v = 0;
#pragma unroll
        for (int y = 0; y < MBSV_MAX_EXP; ++y) if (y == x) v = Et[y]; return v; };
    auto get_L = [&](int x) { long long v = 0;
#pragma unroll
        for (int y = 0; y < MBSV_MAX_EXP; ++y) if (y == x) v = Lt[y]; return v; };
    auto get_LB = [&](int x) { double v = 0;
#pragma unroll
        for (int y = 0; y < MBSV_MAX_EXP; ++y) if (y == x) v = LB[y]; return v; };
    auto set_LB = [&](int x, double v) {
#pragma unroll
        for (int y = 0; y < MBSV_MAX_EXP; ++y) if (y == x) LB[y] = v; };
    // lane 0 moves the counts of one alignment by `sgn`; inside the long-burn-in window (d != 0) every unit step emits the
    // support-histogram difference tallies, exactly as the MH kernels do (value n leaves: +d, value n' enters: -d)
    auto move_counts = [&](int al, int sgn, unsigned d) {
        const int4 rec = P.aln_rec[al];
        const int ne = rec.w >> 8;
        for (int j = 0; j < ne; ++j) {
            const int slot = P.aln_esp_slot[rec.z + j];
            if (slot < 0) continue;
            const int n = st[F + slot];
            if (d) {
                unsigned int* h = R.cluster_hist + P.slot_hist_ptr[c0 * E + slot];
                if (n > 0) atomicAdd(&h[n], d);
                if (n + sgn > 0) atomicAdd(&h[n + sgn], 0u - d);
            }
            st[F + slot] = n + sgn;
        }
    };
    // ---- connected components (lane 0 only; see the header) ----
    const bool has_cc = P.sub_cc_scratch != nullptr && P.sub_cc_scratch[sub] >= 0;
    const int cc_base = F + C * E;
    int err = 0;
    auto cc_nroots = [&](int cl) -> int { return P.cluster_root_ptr[c0 + cl + 1] - P.cluster_root_ptr[c0 + cl]; };
    auto cc_lpc = [&](int cl) -> double {
        const int nr = cc_nroots(cl);
        const int b0 = cc_base + P.cluster_cc_off[c0 + cl] + 1;
        double lpc = 0.0;
        for (int x = 0; x < E; ++x) {
            const int b = b0 + x * (1 + nr);
            const int len = st[b];
            for (int i = 0; i < len; ++i) lpc = lpc + P.T[(size_t)x * Tstride + st[b + 1 + i]];
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
                const int a = st[fg - f0];
                if (a != -1) {
                    const int4 rec = P.aln_rec[P.frag_aln_ptr[fg] + a];
                    const int ne = rec.w >> 8;
                    for (int j = 0; j < ne; ++j)
                        if (P.aln_esp_slot[rec.z + j] == -(2 + cl) && P.aln_esp_leaf[rec.z + j] == l) { flag = (rec.w & 0xff) + 1; ++numselected; }
                }
            }
            st[sl + l] = flag;
        }
        for (int x = 0; x < E; ++x) st[b0 + x * (1 + nr)] = 0;
        if (numselected == 0) return;
        for (int x = 0; x < E; ++x) {
            const int b = b0 + x * (1 + nr);
            for (int l = 0; l < nl; ++l) st[sl + l] &= 0xff;
            for (int r = 0; r < nr; ++r) st[sr + r] = 0;
            int len = 0;
            for (;;) {
                int maxcount = 0, maxind = -1;
                for (int r = 0; r < nr; ++r) {
                    if (st[sr + r]) continue;
                    int cur = 0;
                    for (int j = P.root_leaf_ptr[r0 + r]; j < P.root_leaf_ptr[r0 + r + 1]; ++j)
                        if (st[sl + P.root_leaf[j]] == x + 1) ++cur;
                    if (cur > maxcount) { maxcount = cur; maxind = r; }
                }
                if (maxind == -1) {
                    for (int l = 0; l < nl; ++l) if (st[sl + l] == x + 1) err = -5;   // Error at AM.java:375-376
                    break;
                }
                for (int j = P.root_leaf_ptr[r0 + maxind]; j < P.root_leaf_ptr[r0 + maxind + 1]; ++j)
                    if ((st[sl + P.root_leaf[j]] & 0xff) == x + 1) st[sl + P.root_leaf[j]] = (x + 1) | 0x100;
                int pos = len;                                   // sorted insert (Collections.sort at AM.java:403)
                while (pos > 0 && st[b + pos] > maxcount) { st[b + 1 + pos] = st[b + pos]; --pos; }
                st[b + 1 + pos] = maxcount;
                ++len;
                st[sr + maxind] = 1;
                bool all = true;
                for (int l = 0; l < nl; ++l) if (st[sl + l] == x + 1) all = false;
                if (all) break;
            }
            st[b] = len;
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
            const int len = st[b];
            for (int i = 0; i < len; ++i) atomicAdd(&h[st[b + 1 + i]], d);
        }
    };
    // the log terms of every connected component (lane 0's value is the group's: broadcast by the caller)
    auto cc_total = [&]() -> double {
        double lp = 0.0;
        if (has_cc && lane == 0)
            for (int cl = 0; cl < C; ++cl) if (P.cluster_cc_off[c0 + cl] >= 0) lp += cc_lpc(cl);
        return __shfl_sync(gmask, lp, 0, G);
    };

    auto full_logprob = [&]() {
        double lp = 0.0;
        for (int i = lane; i < C * E; i += G) {
            const int n = st[F + i];
            if (n > 0) lp += P.T[(size_t)(i % E) * Tstride + n];
        }
        lp = group_sum<G>(gmask, lp);
        if (has_cc) lp += cc_total();
        lp += (double)nerr * logperr;
        for (int x = 0; x < E; ++x) lp += get_LB(x);
        return lp;
    };

    // ---------------- initial state: the reference's draws (MultiBreakSV.java:999-1015) or the supplied one ----------------
    for (int i = lane; i < R.sub_W[sub]; i += G) st[i] = 0;
    __syncwarp(gmask);
    for (int f = 0; f < F; ++f) {
        const int a0 = P.frag_aln_ptr[f0 + f];
        const int n = P.frag_aln_ptr[f0 + f + 1] - a0;
        int a;
        if (R.init_assign) a = R.init_assign[f0 + f];
        else if (rng.nextDouble() < R.perr) a = -1;
        else a = rng.nextInt(n);
        if (a == -1) ++nerr;
        else {
            const int4 rec = P.aln_rec[a0 + a];
            bump_totals(rec.w & 0xff, rec.x, rec.y);
            if (lane == 0) move_counts(a0 + a, +1, 0u);
        }
        if (lane == 0) {
            st[f] = a;
            if (R.initial_assign) R.initial_assign[(size_t)chain * P.n_frag + f0 + f] = a;
        }
    }
    __syncwarp(gmask);
    if (has_cc) {
        if (lane == 0) for (int cl = 0; cl < C; ++cl) if (P.cluster_cc_off[c0 + cl] >= 0) { st[cc_base + P.cluster_cc_off[c0 + cl]] = 0; cc_recompute(cl); }
        __syncwarp(gmask);
    }
    for (int x = 0; x < E; ++x) set_LB(x, lb_of(x, get_E(x), get_L(x)));
    if (R.initial_logprob) { const double lp = full_logprob(); if (lane == 0) R.initial_logprob[sc] = lp; }

    // ---------------- iterations: redraw one fragment from its conditional ----------------
    long long n_changed = 0;
    for (int t = 0; t < N; ++t) {
        const int f = F > 1 ? rng.nextInt(F) : 0;
        const int a0 = P.frag_aln_ptr[f0 + f];
        const int n = P.frag_aln_ptr[f0 + f + 1] - a0;
        const int cur = st[f];
        // state R = everything else: take the current assignment out
        int xcur = -1;
        if (cur != -1) {
            const int4 rec = P.aln_rec[a0 + cur];
            xcur = rec.w & 0xff;
            bump_totals(xcur, -(long long)rec.x, -(long long)rec.y);
            if (lane == 0) move_counts(a0 + cur, -1, 0u);
        } else --nerr;
        __syncwarp(gmask);
        const double lb_cur_removed = xcur >= 0 ? lb_of(xcur, get_E(xcur), get_L(xcur)) : 0.0;
        if (has_cc) {
            // lane 0: wopt[o] = sum over the components option o touches of (log term with the option - log term with the
            // fragment unassigned); the multisets as they were before this iteration are saved once per touched component
            if (lane == 0) {
                auto save = [&](int cl) {
                    const int o = cc_base + P.cluster_cc_off[c0 + cl];
                    if (st[o] == t + 1) return;
                    st[o] = t + 1;
                    const int sz = E * (1 + cc_nroots(cl));
                    for (int i = 0; i < sz; ++i) st[o + 1 + sz + i] = st[o + 1 + i];
                };
                if (cur != -1) cc_visit(a0 + cur, save);
                wopt[0] = 0.0;
                for (int o = 1; o <= n; ++o) {
                    double dcc = 0.0;
                    const int4 rec = P.aln_rec[a0 + o - 1];
                    const int ne = rec.w >> 8;
                    for (int j = 0; j < ne; ++j) {
                        const int code = P.aln_esp_slot[rec.z + j];
                        if (code >= -1) continue;
                        bool seen = false;                       // an alignment may hold several leaves of one component
                        for (int i = 0; i < j; ++i) if (P.aln_esp_slot[rec.z + i] == code) seen = true;
                        if (seen) continue;
                        const int cl = -code - 2;
                        save(cl);
                        st[f] = -1;
                        cc_recompute(cl);
                        const double base = cc_lpc(cl);
                        st[f] = o - 1;
                        cc_recompute(cl);
                        dcc += cc_lpc(cl) - base;
                    }
                    wopt[o] = dcc;
                }
                st[f] = cur;
            }
            __syncwarp(gmask);
        }
        // log P(option | rest) up to a constant: option 0 = "unmapped", option k = alignment k-1
        const int nopt = n + 1;
        double m = -1.7976931348623157e308;
        for (int o = lane; o < nopt; o += G) {
            double dl;
            if (o == 0) dl = logperr;
            else {
                const int4 rec = P.aln_rec[a0 + o - 1];
                const int x = rec.w & 0xff, ne = rec.w >> 8;
                double dcov = 0.0;
                for (int j = 0; j < ne; ++j) {
                    const int slot = P.aln_esp_slot[rec.z + j];
                    if (slot < 0) continue;
                    int nn = st[F + slot];
                    for (int i = 0; i < j; ++i) if (P.aln_esp_slot[rec.z + i] == slot) ++nn;     // the same cluster twice in one alignment
                    dcov += P.T[(size_t)x * Tstride + nn + 1] - (nn > 0 ? P.T[(size_t)x * Tstride + nn] : 0.0);
                }
                const double base = x == xcur ? lb_cur_removed : get_LB(x);
                dl = dcov + (lb_of(x, get_E(x) + rec.x, get_L(x) + rec.y) - base);
            }
            if (has_cc) dl += wopt[o];
            wopt[o] = dl;
            m = dl > m ? dl : m;
        }
        m = group_max<G>(gmask, m);
        __syncwarp(gmask);
        double s = 0.0;
        for (int o = lane; o < nopt; o += G) { const double w = exp_dev(wopt[o] - m); wopt[o] = w; s += w; }
        s = group_sum<G>(gmask, s);
        __syncwarp(gmask);
        const double u = rng.nextDouble() * s;
        // categorical draw: inclusive prefix scan of the weights across lanes, ballot for the first lane past u
        int pick = nopt - 1;
        double run = 0.0;
        for (int o0 = 0; o0 < nopt; o0 += G) {
            const int o = o0 + lane;
            double c = o < nopt ? wopt[o] : 0.0;
#pragma unroll
            for (int k = 1; k < G; k <<= 1) { const double v = __shfl_up_sync(gmask, c, k, G); if (lane >= k) c += v; }
            c += run;
            const unsigned hit = __ballot_sync(gmask, o < nopt && c > u) >> (grp * G);
            if (hit) { pick = o0 + __ffs(hit) - 1; break; }
            run = __shfl_sync(gmask, c, G - 1, G);
        }
        __syncwarp(gmask);
        const int now = pick - 1;
        // put the drawn assignment in; tallies only when the state changed
        if (xcur >= 0) set_LB(xcur, lb_cur_removed);
        const unsigned d = (now != cur && t > win_base) ? (unsigned)(t - win_base) : 0u;
        if (lane == 0) {
            if (now == cur) { if (cur != -1) move_counts(a0 + cur, +1, 0u); }
            else {
                if (cur != -1) { move_counts(a0 + cur, +1, 0u); move_counts(a0 + cur, -1, d); }      // leave with tallies
                if (now != -1) move_counts(a0 + now, +1, d);
                if (d) {
                    if (cur == -1) atomicAdd(&R.frag_err_count[f0 + f], d); else atomicAdd(&R.aln_count[a0 + cur], d);
                    if (now == -1) atomicAdd(&R.frag_err_count[f0 + f], 0u - d); else atomicAdd(&R.aln_count[a0 + now], 0u - d);
                }
                st[f] = now;
            }
            if (has_cc) {
                // every component an alignment of this fragment touches: the multisets of the state after the draw, then
                // (inside the window) old elements leave, new ones enter; the stamps are cleared
                for (int o = 0; o < n; ++o) cc_visit(a0 + o, [&](int cl) {
                    const int b = cc_base + P.cluster_cc_off[c0 + cl];
                    if (st[b] != t + 1) return;
                    st[b] = 0;
                    cc_recompute(cl);
                    if (d) { cc_hist(cl, 1, d); cc_hist(cl, 0, 0u - d); }
                });
            }
        }
        if (has_cc && __any_sync(gmask, err != 0)) break;      // set cover failed to cover: the reference throws Error
        if (now != -1) {
            const int4 rec = P.aln_rec[a0 + now];
            const int x = rec.w & 0xff;
            bump_totals(x, rec.x, rec.y);
            set_LB(x, lb_of(x, get_E(x), get_L(x)));
        } else ++nerr;
        if (now != cur) ++n_changed;
        __syncwarp(gmask);
    }

    // ---------------- epilogue: the final state stays until the end of the window ----------------
    const long long nwin = (long long)N - win_base;
    const bool failed = has_cc && __any_sync(gmask, err != 0);
    if (nwin > 0 && !failed) {
        const unsigned d = (unsigned)nwin;
        for (int f = lane; f < F; f += G) {
            const int a = st[f];
            if (a == -1) atomicAdd(&R.frag_err_count[f0 + f], d);
            else atomicAdd(&R.aln_count[P.frag_aln_ptr[f0 + f] + a], d);
        }
        for (int i = lane; i < C * E; i += G) {
            const int nn = st[F + i];
            if (nn > 0) atomicAdd(&(R.cluster_hist + P.slot_hist_ptr[c0 * E + i])[nn], d);
        }
        if (has_cc && lane == 0) for (int cl = 0; cl < C; ++cl) if (P.cluster_cc_off[c0 + cl] >= 0) cc_hist(cl, 0, d);
    }
    if (R.final_assign) for (int f = lane; f < F; f += G) R.final_assign[(size_t)chain * P.n_frag + f0 + f] = st[f];
    const double lpf = full_logprob();
    if (lane == 0) {
        if (R.final_logprob) R.final_logprob[sc] = lpf;
        if (R.counters) { long long* k = R.counters + sc * 5; k[0] = N; k[1] = n_changed; k[2] = 0; k[3] = 0; k[4] = N; }
        if (R.rng_state) R.rng_state[sc] = rng.state();
        if (R.error) R.error[sc] = err;
        if (R.totals) {
            atomicAdd(&R.totals[0], (unsigned long long)N); atomicAdd(&R.totals[1], (unsigned long long)n_changed);
            atomicAdd(&R.totals[4], (unsigned long long)N);
            if (err) atomicCAS(&R.totals[5], 0ULL, (unsigned long long)(long long)err);
        }
    }
}

}  // namespace mbsv
