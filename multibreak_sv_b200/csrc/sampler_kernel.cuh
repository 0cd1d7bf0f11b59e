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

#include "jmath.cuh"

namespace mbsv {

struct DevProblem {
    int n_sub, n_frag, n_aln, n_cluster, n_exp, max_support, maxn;
    const int* sub_frag_ptr;
    const int* sub_cluster_ptr;
    const int* sub_sv_ptr;
    const double* sub_logF;          // log(F_s)  (Math.log(fragments.size()), MultiBreakSV.java:1219-1271)
    const int* frag_aln_ptr;
    const int4* aln_rec;             // {numerrors, len, first ESP entry, exp | nesp << 8}
    const int* aln_esp_slot;         // per alignment-ESP entry: local count slot cluster*E + exp, or -1
    const int* supports_order_local; // per subproblem: local cluster indices in A.supports.keySet() order
    const int* cluster_hist_ptr;
    const int* sv_frag_ptr;
    const int* sv_frag;
    const int* sv_numchanged;
    const double* T;                 // logprobcov [E][max_support+1]
    const double* logint;            // log(n), n = 0..maxn
    const double* exp_pseq;
    const double* exp_logp;
    const double* exp_log1mp;
    double log2pi;                   // Math.log(2)+Math.log(Math.PI)
};

struct DevRun {
    int n_chains, num_iters, burnin, win_base, java_int_wrap;
    double perr, logperr, log1mperr;
    long long seed;
    const int* init_assign;
    // plan: item g = warp*32 + lane is chain (g % n_chains) of subproblem sub_order[g / n_chains]
    const int* sub_order;            // subproblems by ascending state size
    const int* sub_W;                // state rows per chain of each subproblem
    long long n_items;               // n_sub * n_chains
    const long long* warp_row_off;   // [warp - global_warp0] row offset into workspace (global class)
    int* workspace;
    int warp_begin, warp_end, rows_per_warp, global_warp0;
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
};

// MultiBreakSV.java:1661-1695 (logBinomial, logbinomialcoefficient, logNormal)
__device__ __noinline__ double log_binomial_dev(long long k, long long n, double p, double logp, double log1mp,
                                                double log2pi) {
    double dn = (double)n;
    if (dn * p <= 5 || dn * (1 - p) <= 5) {
        double t = 0;
        long long r = k, m = n - k;
        if (r < m) r = m;
        for (long long i = n, j = 1; i > r; --i, ++j) t = t + jm_log((double)i) - jm_log((double)j);
        return t + (double)k * logp + (double)(n - k) * log1mp;
    }
    double mu = dn * p;
    double sig = dn * p * (1 - p);
    double d = (double)k - mu;
    return -(d * d) / (2 * sig) - (log2pi + jm_log(sig)) / 2;
}

__device__ __noinline__ double exp_dev(double x) { return jm_exp(x); }

#define MBSV_ST(i) st[(size_t)(i) << 5]

template <int MODE, bool TRACE, bool SMEM, int MAXE>
__global__ void __launch_bounds__(128) mbsv_chain_kernel(const __grid_constant__ DevProblem P,
                                                         const __grid_constant__ DevRun R) {
    extern __shared__ int smem_tile[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int warp = R.warp_begin + blockIdx.x * 4 + wic;
    if (warp >= R.warp_end) return;
    const long long item = (long long)warp * 32 + lane;
    if (item >= R.n_items) return;
    const int sub = R.sub_order[item / R.n_chains];
    const int chain = (int)(item % R.n_chains);
    int* st = SMEM ? (smem_tile + (size_t)wic * R.rows_per_warp * 32 + lane)
                   : (R.workspace + (size_t)R.warp_row_off[warp - R.global_warp0] * 32 + lane);

    const int f0 = P.sub_frag_ptr[sub], F = P.sub_frag_ptr[sub + 1] - f0;
    const int c0 = P.sub_cluster_ptr[sub], C = P.sub_cluster_ptr[sub + 1] - c0;
    const int E = P.n_exp;
    const int sv0 = P.sub_sv_ptr[sub], nsv = P.sub_sv_ptr[sub + 1] - sv0;
    const int Tstride = P.max_support + 1;
    const long long sc = (long long)sub * R.n_chains + chain;
    const int N = R.num_iters;
    const double logF = P.sub_logF[sub];
    const double neglogF = -logF;
    const double perr = R.perr, logperr = R.logperr;
    const double dF = (double)F;

    JavaRandom rng;
    rng.seed(R.seed + chain);

    long long Etot[MAXE], Ltot[MAXE];
    double LB[MAXE];
#pragma unroll
    for (int x = 0; x < MAXE; ++x) { Etot[x] = 0; Ltot[x] = 0; LB[x] = 0.0; }
    int nerr = 0;
    int err = 0;
    long long n_naive_prop = 0, n_naive_acc = 0, n_sv_prop = 0, n_sv_acc = 0, n_reassign = 0;

    auto wrapi = [&](long long t) -> long long { return R.java_int_wrap ? (long long)(int)(unsigned int)(unsigned long long)t : t; };
    auto lb_of = [&](int x, long long e, long long l) -> double {
        return log_binomial_dev(wrapi(e), wrapi(l), P.exp_pseq[x], P.exp_logp[x], P.exp_log1mp[x], P.log2pi);
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
    // MultiBreakSVAssignmentMatrix.java:130-246 in the reference's summation order.
    auto full_logprob = [&]() -> double {
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
        const int W = F + C * E;
        for (int i = 0; i < W; ++i) MBSV_ST(i) = 0;
        double dc = 0.0, de = 0.0;
        unsigned tch = 0;
        nerr = F;   // move_counts(-1 -> a) decrements
        for (int f = 0; f < F; ++f) {
            const int a0 = P.frag_aln_ptr[f0 + f];
            const int n = P.frag_aln_ptr[f0 + f + 1] - a0;
            int a;
            if (R.init_assign) a = R.init_assign[f0 + f];
            else {
                if (rng.nextDouble() < perr) a = -1;
                else a = rng.nextInt(n);
            }
            MBSV_ST(f) = a;
            if (a != -1) move_counts(a0, -1, a, dc, de, tch);
            if (R.initial_assign) R.initial_assign[(size_t)chain * P.n_frag + f0 + f] = a;
        }
#pragma unroll
        for (int x = 0; x < MAXE; ++x) if (x < E) LB[x] = lb_of(x, Etot[x], Ltot[x]);
    }
    double logprob = full_logprob();
    if (R.initial_logprob) R.initial_logprob[sc] = logprob;

    // ---------------- iterations (MultiBreakSV.java:796-837, 1101-1182) ----------------
    int iter = 0;
    while (iter < N && err == 0) {
        // lazy chain: with probability 1/2 stay (:1106-1113)
        double r = rng.nextDouble();
        if (r < 0.5) {
            if (TRACE) {
                R.trace_logprob[sc * N + iter] = logprob;
                R.trace_move_frag[sc * N + iter] = -1;
                R.trace_move_assign[sc * N + iter] = -1;
            }
            ++iter;
            continue;
        }
        const int t = iter;
        r = rng.nextDouble();
        const bool naive = (r < 0.9) || (nsv == 0);
        int f = -1, a0 = 0, prev = -1, now = -1, svk = -1;
        double Alogq, Anewlogq;
        if (naive) {
            // ---- naiveMove (:1184-1273) ----
            ++n_naive_prop; ++n_reassign;
            f = (int)(rng.nextDouble() * dF);
            a0 = P.frag_aln_ptr[f0 + f];
            const int n = P.frag_aln_ptr[f0 + f + 1] - a0;
            prev = MBSV_ST(f);
            r = rng.nextDouble();
            if (prev == -1) {
                int thisassign = 0;
                const double inc = (double)1 / n;
                double sum = inc;
                while (thisassign < n - 1 && r > sum) { ++thisassign; sum += inc; }
                now = thisassign;
                if (n == 1) { Alogq = neglogF + 0.0; Anewlogq = neglogF + 0.0; }
                else { Alogq = neglogF + logperr; Anewlogq = neglogF - P.logint[n]; }
            } else if (r < perr || n == 1) {
                now = -1;
                if (n == 1) { Alogq = neglogF + 0.0; Anewlogq = neglogF + 0.0; }
                else { Alogq = neglogF - P.logint[n]; Anewlogq = neglogF + logperr; }
            } else {
                int thisassign = 0;
                r = rng.nextDouble();
                const double inc = (double)1 / (n - 1);
                double sum = inc;
                while (thisassign < n - 1 && (thisassign == prev || r > sum)) {
                    ++thisassign;
                    if (thisassign != prev) sum += inc;
                }
                now = thisassign;
                Anewlogq = neglogF + R.log1mperr - P.logint[n - 1];
                Alogq = Anewlogq;
            }
            if (now == prev) { err = -5; break; }   // Error at :1122-1123
        } else {
            // ---- svMove (:1291-1343) ----
            ++n_sv_prop;
            svk = (int)(rng.nextDouble() * (double)nsv);
            n_reassign += P.sv_numchanged[sv0 + svk];
            const double l = jm_log((double)nsv);
            Alogq = -l;
            Anewlogq = -l;
        }
        // ---- tentative application ----
        double dcov = 0.0, derr = 0.0;
        unsigned touched = 0;
        long long E0[MAXE], L0[MAXE];
#pragma unroll
        for (int x = 0; x < MAXE; ++x) { E0[x] = Etot[x]; L0[x] = Ltot[x]; }
        const int nerr0 = nerr;
        if (naive) move_counts(a0, prev, now, dcov, derr, touched);
        else {
            for (int q = P.sv_frag_ptr[sv0 + svk]; q < P.sv_frag_ptr[sv0 + svk + 1]; ++q) {
                const int g = P.sv_frag[q] - f0;
                const int pv = MBSV_ST(g);
                move_counts(P.frag_aln_ptr[f0 + g], pv, pv == -1 ? 0 : -1, dcov, derr, touched);
            }
        }
        // ---- acceptance ratio (:1146) ----
        double newLB[MAXE];
        double newlogprob, arg;
        if (MODE == 0) {
            // full re-sum in the reference's order; the assignment change only enters through nerr, counts, totals
#pragma unroll
            for (int x = 0; x < MAXE; ++x) newLB[x] = LB[x];
            newlogprob = full_logprob();
            arg = (newlogprob + Alogq) - (logprob + Anewlogq);
        } else {
            double dlb = 0.0;
#pragma unroll
            for (int x = 0; x < MAXE; ++x) {
                newLB[x] = LB[x];
                if (x < E && ((touched >> x) & 1u)) {
                    newLB[x] = lb_of(x, Etot[x], Ltot[x]);
                    dlb = dlb + (newLB[x] - LB[x]);
                }
            }
            const double delta = (dcov + derr) + dlb;
            newlogprob = logprob + delta;
            arg = (delta + Alogq) - Anewlogq;
        }
        const double e = exp_dev(arg);
        const double ratio = (e != e) ? e : (1.0 < e ? 1.0 : e);   // Math.min(1, e)
        r = rng.nextDouble();
        const bool accept = ratio > r;
        if (accept) {
            logprob = newlogprob;
#pragma unroll
            for (int x = 0; x < MAXE; ++x) LB[x] = newLB[x];
            if (naive) {
                ++n_naive_acc;
                MBSV_ST(f) = now;
                if (t > win_base) {
                    tally_frag(f, a0, prev, now, t);
                    retally_fragment(a0, prev, now, t);
                    reapply_fragment(a0, prev, now);
                }
            } else {
                ++n_sv_acc;
                const int q0 = P.sv_frag_ptr[sv0 + svk], q1 = P.sv_frag_ptr[sv0 + svk + 1];
                if (t > win_base) {
                    for (int q = q1 - 1; q >= q0; --q) {
                        const int g = P.sv_frag[q] - f0;
                        const int pv = MBSV_ST(g);
                        retally_fragment(P.frag_aln_ptr[f0 + g], pv, pv == -1 ? 0 : -1, t);
                    }
                    for (int q = q0; q < q1; ++q) {
                        const int g = P.sv_frag[q] - f0;
                        const int pv = MBSV_ST(g);
                        const int ga0 = P.frag_aln_ptr[f0 + g];
                        reapply_fragment(ga0, pv, pv == -1 ? 0 : -1);
                        tally_frag(g, ga0, pv, pv == -1 ? 0 : -1, t);
                    }
                }
                for (int q = q0; q < q1; ++q) {
                    const int g = P.sv_frag[q] - f0;
                    const int pv = MBSV_ST(g);
                    MBSV_ST(g) = (pv == -1) ? 0 : -1;
                }
            }
        } else {
            // revert
            double d1 = 0.0, d2 = 0.0;
            unsigned tt = 0;
            if (naive) move_counts(a0, now, prev, d1, d2, tt);
            else {
                for (int q = P.sv_frag_ptr[sv0 + svk + 1] - 1; q >= P.sv_frag_ptr[sv0 + svk]; --q) {
                    const int g = P.sv_frag[q] - f0;
                    const int pv = MBSV_ST(g);   // assignment was not written yet: pv is still the old value
                    move_counts(P.frag_aln_ptr[f0 + g], pv == -1 ? 0 : -1, pv, d1, d2, tt);
                }
            }
#pragma unroll
            for (int x = 0; x < MAXE; ++x) { Etot[x] = E0[x]; Ltot[x] = L0[x]; }
            nerr = nerr0;
        }
        if (TRACE) {
            R.trace_logprob[sc * N + iter] = logprob;
            R.trace_move_frag[sc * N + iter] = accept ? (naive ? f0 + f : -(2 + svk)) : -1;
            R.trace_move_assign[sc * N + iter] = (accept && naive) ? now : -1;
        }
        ++iter;
    }

    // ---------------- epilogue: final tallies, state, counters ----------------
    if (err == 0) {
        const long long nwin = (long long)N - win_base;
        if (nwin > 0) {
            const unsigned d = (unsigned)nwin;
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
    if (R.final_logprob) R.final_logprob[sc] = (MODE == 0) ? logprob : full_logprob();
    if (R.counters) {
        long long* k = R.counters + sc * 5;
        k[0] = n_naive_prop; k[1] = n_naive_acc; k[2] = n_sv_prop; k[3] = n_sv_acc; k[4] = n_reassign;
    }
    if (R.rng_state) R.rng_state[sc] = rng.state();
    if (R.error) R.error[sc] = err;
}

}  // namespace mbsv
