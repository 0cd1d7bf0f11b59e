// Java-compatible scalar math for the sampler, usable from host and device code.
//
//   jm_log / jm_exp  — the fdlibm 5.3 algorithms (__ieee754_log / __ieee754_exp) that Java's StrictMath
//                      specifies; the reference calls Math.log / Math.exp at MultiBreakSV.java:1146,
//                      1219-1271,1342-1343,1645-1693 and MultiBreakSVAssignmentMatrix.java:223.
//   JavaRandom       — java.util.Random's 48-bit LCG (MultiBreakSV.java:78 and its call sites).
//
// This translation unit must be compiled WITHOUT floating-point contraction (nvcc -fmad=false, host
// -ffp-contract=off): every operation below is an individually rounded IEEE-754 double operation, so a
// seeded chain takes the same decisions on the device as on any IEEE host.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define JM_HD __host__ __device__ __forceinline__
#define JM_CX __host__ __device__ constexpr
#else
#define JM_HD inline
#define JM_CX constexpr
#endif

namespace mbsv {

JM_HD int32_t jm_hi(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    uint64_t u; std::memcpy(&u, &x, 8); return (int32_t)(u >> 32);
#endif
}
JM_HD uint32_t jm_lo(double x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(x);
#else
    uint64_t u; std::memcpy(&u, &x, 8); return (uint32_t)u;
#endif
}
JM_HD double jm_make(int32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, (int)lo);
#else
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | lo; double d; std::memcpy(&d, &u, 8); return d;
#endif
}
JM_HD double jm_bits(uint64_t u) { return jm_make((int32_t)(u >> 32), (uint32_t)u); }

// fdlibm e_log.c.  The two polynomial-correction forms (selected by `i > 0`) and the k == 0 special cases are
// evaluated through one straight-line path and a select; both rewrites are bit-identical to fdlibm's
// branches (negation commutes with IEEE subtraction/division, and adding dk*ln2 terms with dk == +0.0 is
// exact), they only remove warp divergence.  The rare paths (zero, negative, subnormal, inf/NaN,
// |x-1| < 2^-20) keep their branches.
JM_HD double jm_log(double x) {
    const double ln2_hi = jm_bits(0x3fe62e42fee00000ULL), ln2_lo = jm_bits(0x3dea39ef35793c76ULL);
    const double two54 = jm_bits(0x4350000000000000ULL);
    const double Lg1 = jm_bits(0x3FE5555555555593ULL), Lg2 = jm_bits(0x3FD999999997FA04ULL),
                 Lg3 = jm_bits(0x3FD2492494229359ULL), Lg4 = jm_bits(0x3FCC71C51D8E78AFULL),
                 Lg5 = jm_bits(0x3FC7466496CB03DEULL), Lg6 = jm_bits(0x3FC39A09D078C69FULL),
                 Lg7 = jm_bits(0x3FC2F112DF3E5244ULL);
    int32_t hx = jm_hi(x);
    uint32_t lx = jm_lo(x);
    int32_t k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -two54 / 0.0;
        if (hx < 0) return (x - x) / 0.0;
        k -= 54; x *= two54; hx = jm_hi(x); lx = jm_lo(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    x = jm_make(hx | (i ^ 0x3ff00000), lx);
    k += (i >> 20);
    const double f = x - 1.0;
    const double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return dk * ln2_hi + dk * ln2_lo;
        }
        const double R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    const double s = f / (2.0 + f);
    const double z = s * s;
    i = hx - 0x6147a;
    const double w = z * z;
    const int32_t j = 0x6b851 - hx;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dkhi = dk * ln2_hi, dklo = dk * ln2_lo;
    const double A = dkhi - ((hfsq - (s * (hfsq + R) + dklo)) - f);   // i > 0
    const double B = dkhi - ((s * (f - R) - dklo) - f);               // i <= 0
    return i > 0 ? A : B;
}

// fdlibm e_exp.c.  The three argument-reduction cases (|x| <= 0.5 ln2, < 1.5 ln2, general) run through the
// general formula with k = 0, +-1, k: bit-identical (k*ln2HI is exact for those k, 0 - q == -q, and
// (x*c)/(c-2) == -((x*c)/(2-c))), and free of divergence.  Overflow/underflow/NaN and |x| < 2^-28 keep
// their (rare) branches.
JM_HD double jm_exp(double x) {
    const double huge = 1.0e+300;
    const double twom1000 = jm_bits(0x0170000000000000ULL);
    const double o_threshold = jm_bits(0x40862E42FEFA39EFULL);
    const double u_threshold = jm_bits(0xc0874910D52D3051ULL);
    const double ln2HI = jm_bits(0x3fe62e42fee00000ULL), ln2LO = jm_bits(0x3dea39ef35793c76ULL);
    const double invln2 = jm_bits(0x3ff71547652b82feULL);
    const double P1 = jm_bits(0x3FC555555555553EULL), P2 = jm_bits(0xBF66C16C16BEBD93ULL),
                 P3 = jm_bits(0x3F11566AAF25DE2CULL), P4 = jm_bits(0xBEBBBD41C5D26BF1ULL),
                 P5 = jm_bits(0x3E66376972BEA4D0ULL);
    uint32_t hx = (uint32_t)jm_hi(x);
    const int32_t xsb = (int32_t)((hx >> 31) & 1);
    hx &= 0x7fffffff;
    if (hx >= 0x40862E42) {
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | jm_lo(x)) != 0) return x + x;
            return (xsb == 0) ? x : 0.0;
        }
        if (x > o_threshold) return huge * huge;
        if (x < u_threshold) return twom1000 * twom1000;
    }
    if (hx < 0x3e300000) {
        if (huge + x > 1.0) return 1.0 + x;
    }
    int32_t k = 0;
    if (hx > 0x3fd62e42) k = (int32_t)(invln2 * x + (xsb ? -0.5 : 0.5));
    const double tk = (double)k;
    const double hi = x - tk * ln2HI;
    const double lo = tk * ln2LO;
    x = hi - lo;
    const double t = x * x;
    const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) return jm_make(jm_hi(y) + (k << 20), jm_lo(y));
    y = jm_make(jm_hi(y) + ((k + 1000) << 20), jm_lo(y));
    return y * twom1000;
}

// java.util.Random.  The state is kept unmasked in 64 bits: multiplication and addition modulo 2^64 keep
// the low 48 bits exact, and every extraction masks.
struct JavaRandom {
    uint64_t s;
    static constexpr uint64_t MULT = 0x5DEECE66DULL, ADD = 0xBULL, MASK = (1ULL << 48) - 1;
    // two steps at once: s2 = MULT^2 * s + (MULT*ADD + ADD)
    static constexpr uint64_t MULT2 = (MULT * MULT) & MASK, ADD2 = (MULT * ADD + ADD) & MASK;
    // jump-ahead: the state m steps after s is jump_mul(m) * s + jump_add(m)  (mod 2^48)
    static JM_CX uint64_t jump_mul(int m) { uint64_t a = 1; for (int i = 0; i < m; ++i) a = (a * MULT) & MASK; return a; }
    static JM_CX uint64_t jump_add(int m) { uint64_t c = 0; for (int i = 0; i < m; ++i) c = (c * MULT + ADD) & MASK; return c; }
    JM_HD void seed(int64_t sd) { s = ((uint64_t)sd ^ MULT) & MASK; }
    JM_HD uint32_t next(int bits) {
        s = s * MULT + ADD;
        return (uint32_t)((s & MASK) >> (48 - bits));
    }
    // advance by m steps without producing values (draws whose value cannot influence the chain)
    template <int M_STEPS> JM_HD void skip() { s = s * jump_mul(M_STEPS) + jump_add(M_STEPS); }
    // the 53-bit integer of nextDouble(): nextDouble() == bits53() * 2^-53, so `nextDouble() < p` is the integer
    // comparison bits53() < ceil(p * 2^53)
    JM_HD uint64_t bits53() {
        uint64_t s1 = s * MULT + ADD;
        uint64_t s2 = s * MULT2 + ADD2;
        s = s2;
        return (((s1 & MASK) >> 22) << 27) + ((s2 & MASK) >> 21);
    }
    static JM_HD double to_double(uint64_t m53) {
#if defined(__CUDA_ARCH__)
        const double da = __hiloint2double(0x43300000, (int)(m53 >> 27)) - 4503599627370496.0;
        const double db = __hiloint2double(0x43300000, (int)(m53 & 0x7FFFFFFu)) - 4503599627370496.0;
        return da * 0x1.0p-26 + db * 0x1.0p-53;
#else
        return (double)(int64_t)m53 * 0x1.0p-53;
#endif
    }
    JM_HD double nextDouble() {
        uint64_t s1 = s * MULT + ADD;      // independent of s2: both multiplies issue back to back
        uint64_t s2 = s * MULT2 + ADD2;
        s = s2;
        uint64_t a = (s1 & MASK) >> 22;    // next(26)
        uint64_t b = (s2 & MASK) >> 21;    // next(27)
#if defined(__CUDA_ARCH__)
        // ((a << 27) + b) * 2^-53 without the 64-bit int->double conversion: 2^52 + v is built from bits,
        // every step below is exact, so the value equals the host expression.
        const double da = __hiloint2double(0x43300000, (int)a) - 4503599627370496.0;
        const double db = __hiloint2double(0x43300000, (int)b) - 4503599627370496.0;
        return da * 0x1.0p-26 + db * 0x1.0p-53;
#else
        return (double)(int64_t)((a << 27) + b) * 0x1.0p-53;
#endif
    }
    JM_HD int32_t nextInt(int32_t bound) {
        int32_t r = (int32_t)next(31);
        int32_t m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        int32_t u = r;
        for (;;) {
            r = u % bound;
            if ((int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m) < 0) u = (int32_t)next(31); else break;
        }
        return r;
    }
    JM_HD uint64_t state() const { return s & MASK; }
};

}  // namespace mbsv
