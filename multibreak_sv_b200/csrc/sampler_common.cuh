// Device helpers shared by the sampler kernels (sampler_kernel.cuh: generic MH kernel; memo_kernel.cuh: memoised
// kernel): Java-compatible math wrappers, the lazy-iteration look-ahead, the Naive proposal draw, counter reduction.
#pragma once
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#include "jmath.cuh"
#include "sampler_types.cuh"

namespace mbsv {

#define MBSV_MAX_EXP 4

// logtab[i] = jm_log(i) for 0 < i <= ntab: the exact branch walks Math.log(i) - Math.log(j) over small integers.
#ifdef MBSV_INLINE_MATH
#define MBSV_MATH_INLINE __forceinline__
#else
#define MBSV_MATH_INLINE __noinline__
#endif
static __device__ __forceinline__ double log_binomial_impl(long long k, long long n, double p, double logp, double log1mp,
                                                double log2pi, const double* __restrict__ logtab, int ntab) {
    double dn = (double)n;
    if (dn * p <= 5 || dn * (1 - p) <= 5) {
        double t = 0;
        long long r = k, m = n - k;
        if (r < m) r = m;
        if (n <= ntab && n >= 1) {
            const int ri = (int)(r < 0 ? 0 : r);
            for (int i = (int)n, j = 1; i > ri; --i, ++j) t = t + logtab[i] - logtab[j];
        } else {
            for (long long i = n, j = 1; i > r; --i, ++j) t = t + jm_log((double)i) - jm_log((double)j);
        }
        return t + (double)k * logp + (double)(n - k) * log1mp;
    }
    double mu = dn * p;
    double sig = dn * p * (1 - p);
    double d = (double)k - mu;
    return -(d * d) / (2 * sig) - (log2pi + jm_log(sig)) / 2;
}

static __device__ MBSV_MATH_INLINE double log_binomial_dev(long long k, long long n, double p, double logp, double log1mp,
                                                double log2pi, const double* __restrict__ logtab, int ntab) {
    return log_binomial_impl(k, n, p, logp, log1mp, log2pi, logtab, ntab);
}

static __device__ MBSV_MATH_INLINE double exp_dev(double x) { return jm_exp(x); }

// Row i of this chain's column of the [rows][st_stride] state tile (32 in shared memory: one bank per lane; the global
// workspace packs the rows of narrow warps, st_stride = chains per warp).  The checked build (make check) traps on a row outside
// the subproblem's rows instead of corrupting a neighbour's state; `st_rows` is defined next to `st` in each kernel.
#ifdef MBSV_BOUNDS_CHECK
template <typename T>
static __device__ __noinline__ T& mbsv_st_checked(T* st, long long i, int rows, int line, int stride) {
    if (i < 0 || i >= rows) {
        printf("MBSV_ST: row %lld outside [0,%d) at line %d (block %d thread %d)\n", i, rows, line, blockIdx.x, threadIdx.x);
        __trap();
    }
    return st[i * (long long)stride];
}
#define MBSV_ST(i) mbsv_st_checked(st, (i), st_rows, __LINE__, st_stride)
#else
#define MBSV_ST(i) st[(long long)(int)(i) * (long long)st_stride]      // one IMAD.WIDE
#endif


// Lazy chain (MultiBreakSV.java:1106-1113): with probability 1/2 an iteration only consumes one nextDouble().
// Branch-free look-ahead over the next LAZY_J iterations: iteration j is lazy iff its first LCG step (2j+1 steps
// from the current state) has bit 47 clear (nextDouble() < 0.5 <=> next(26) < 2^25).  All candidates are
// independent multiply-adds of the current state by jump-ahead constants.  On return either `live` (the chain
// sits on a non-lazy iteration `iter`, its lazy-coin draw consumed) or the look-ahead window was all lazy
// (`iter` advanced, call again) or the chain is `done`.
#ifndef MBSV_LAZY_J
#define MBSV_LAZY_J 5
#endif
__device__ __forceinline__ void lazy_skip(JavaRandom& rng, int& iter, const int N, bool& live, bool& done) {
    constexpr int LAZY_J = MBSV_LAZY_J;   // 2^-J of the rounds find no live iteration; more candidates cost every round.
                                        // Measured on the 5M config (ms per 2000 iterations): J=3 1071, 4 1032, 5 1011, 6 1011, 7 1018
    const uint64_t s0 = rng.s;
    uint64_t cand[LAZY_J];
    unsigned bits = 0;
#pragma unroll
    for (int j = 0; j < LAZY_J; ++j) {
        cand[j] = s0 * JavaRandom::jump_mul(2 * j + 1) + JavaRandom::jump_add(2 * j + 1);
        bits |= ((unsigned)(cand[j] >> 47) & 1u) << j;
    }
    const int jstar = bits ? (__ffs(bits) - 1) : LAZY_J;
    const int remaining = N - iter;
    const int lim = remaining < LAZY_J ? remaining : LAZY_J;
    if (jstar < lim) {
        uint64_t sel = cand[0];
#pragma unroll
        for (int j = 1; j < LAZY_J; ++j) if (jstar == j) sel = cand[j];
        rng.s = sel * JavaRandom::MULT + JavaRandom::ADD;   // second LCG step of that nextDouble()
        iter += jstar;
        live = true;
    } else if (lim == LAZY_J) {
        rng.s = s0 * JavaRandom::jump_mul(2 * LAZY_J) + JavaRandom::jump_add(2 * LAZY_J);
        iter += LAZY_J;
        done = iter >= N;
    } else {
        for (int j = 0; j < lim; ++j) rng.s = rng.s * JavaRandom::MULT2 + JavaRandom::ADD2;
        iter += lim;
        done = true;
    }
}

// naiveMove (MultiBreakSV.java:1198-1272): new assignment of a fragment with n alignments currently at `prev`, and the
// proposal log-probabilities of the reverse (Alogq) and forward (Anewlogq) moves.  Includes quirk Q1: the scan over
// "the other n-1" is biased when prev == 0.  The cumulative sums are accumulated exactly as the reference does.
// ceil(p * 2^53) as an integer threshold: nextDouble() < p  <=>  bits53() < threshold  (p * 2^53 is exact)
__device__ __forceinline__ unsigned long long prob_threshold53(double p) {
    const double x = p * 9007199254740992.0;
    if (!(x > 0.0)) return 0ULL;
    if (x >= 9007199254740992.0) return 1ULL << 53;
    return __double2ull_ru(x);
}

// which alignment naiveMove draws (MultiBreakSV.java:1198-1256); -1 = "unmapped".  Consumes one or two nextDouble()s
// exactly as the reference does.
__device__ __forceinline__ int naive_draw(JavaRandom& rng, const int n, const int prev, const unsigned long long perr_thr,
                                          const double* __restrict__ invint) {
    // The reference has two scans (:1205-1209 from "unmapped": thresholds k/n, no index skipped; :1251-1256 to another
    // alignment: thresholds k/(n-1), index `prev` skipped).  With prev == -1 the second loop never skips, so ONE loop with
    // m = n or n-1 reproduces both, and lanes taking different branches stay converged.
    const unsigned long long m53 = rng.bits53();
    const bool from_err = prev == -1;
    const bool to_err = !from_err && (m53 < perr_thr || n == 1);
    double r = JavaRandom::to_double(m53);
    if (!from_err && !to_err) r = rng.nextDouble();          // the second draw of :1248
    const double inc = invint[from_err ? n : n - 1];       // (double)1 / m, tabulated with the same IEEE division
    double sum = inc;
    int thisassign = 0;
    while (!to_err && thisassign < n - 1 && (thisassign == prev || r > sum)) {
        ++thisassign;
        if (thisassign != prev) sum += inc;
    }
    return to_err ? -1 : thisassign;
}

// proposal log-probabilities of the reverse (Alogq) and forward (Anewlogq) Naive moves (:1218-1271)
__device__ __forceinline__ void naive_logq(const int n, const int prev, const int now, const double neglogF,
                                           const double logperr, const double log1mperr,
                                           const double* __restrict__ logint, double& Alogq, double& Anewlogq) {
    if (n == 1 && (prev == -1 || now == -1)) { Alogq = neglogF + 0.0; Anewlogq = neglogF + 0.0; }
    else if (prev == -1) { Alogq = neglogF + logperr; Anewlogq = neglogF - logint[n]; }
    else if (now == -1) { Alogq = neglogF - logint[n]; Anewlogq = neglogF + logperr; }
    else { Anewlogq = neglogF + log1mperr - logint[n - 1]; Alogq = Anewlogq; }
}

__device__ __forceinline__ void naive_pick(JavaRandom& rng, const int n, const int prev, const double neglogF,
                                           const unsigned long long perr_thr, const double logperr, const double log1mperr,
                                           const double* __restrict__ logint, const double* __restrict__ invint, int& now,
                                           double& Alogq, double& Anewlogq) {
    now = naive_draw(rng, n, prev, perr_thr, invint);
    naive_logq(n, prev, now, neglogF, logperr, log1mperr, logint, Alogq, Anewlogq);
}

// One atomic per counter per converged group of lanes.
__device__ __forceinline__ void warp_add_totals(unsigned long long* totals, int lane, long long v0, long long v1,
                                                long long v2, long long v3, long long v4, int err) {
    const unsigned m = __activemask();
    long long v[5] = {v0, v1, v2, v3, v4};
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        long long x = v[i];
        for (int o = 16; o > 0; o >>= 1) {
            const long long y = __shfl_xor_sync(m, x, o);
            const int src = lane ^ o;
            if ((m >> src) & 1u) x += y;
        }
        if (lane == (__ffs(m) - 1)) atomicAdd(&totals[i], (unsigned long long)x);
    }
    if (err) atomicCAS(&totals[5], 0ULL, (unsigned long long)(long long)err);
}

}  // namespace mbsv
