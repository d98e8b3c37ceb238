/*
 * aps_math.h -- deterministic scalar arithmetic shared by the CUDA kernels and the CPU oracle.
 *
 * Everything in this header is written with IEEE-754 binary64 add / mul / div / sqrt and
 * *explicit* fused multiply-adds only, so that the same source gives bit-identical results
 * when compiled by gcc for x86-64 (-ffp-contract=off) and by nvcc for sm_100a (-fmad=false).
 * No libm / libdevice transcendental is called anywhere on the hot path.
 *
 * What it replaces in the reference (TuringLang/AdvancedPS.jl v0.7.2, paths relative to
 * /root/reference):
 *   - src/rng.jl:2,9-31      TracedRNG around Random123.Philox2x  -> aps_philox2x64 (10 rounds)
 *   - Random.randn / rand(Normal) reached from src/pgas.jl:60-68  -> aps_normal_pair (Box-Muller)
 *   - StatsFuns.logsumexp / softmax (src/container.jl:95,109)     -> aps_exp / aps_log building blocks
 * The third-party Julia arithmetic (ziggurat randn, LogExpFunctions) is NOT vendored in the
 * reference tree, so bit-level parity with Julia's streams is unpinned; Philox2x64-10 itself is
 * pinned against Random123's published known-answer vectors (tests/test_oracle_math.py).
 */
#ifndef APS_MATH_H
#define APS_MATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define APS_HD __host__ __device__ __forceinline__
#else
#define APS_HD static inline
#endif

/* ------------------------------------------------------------------ bit casts / fma */
APS_HD double aps_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

APS_HD uint64_t aps_d2bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
#endif
}

APS_HD double aps_bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}

APS_HD double aps_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}

/* 64x64 -> 128 multiply */
APS_HD void aps_mul64(uint64_t a, uint64_t b, uint64_t *hi, uint64_t *lo) {
#if defined(__CUDA_ARCH__)
    /* nvcc's own 64-bit high / low products share their partial products and use carry-in forms
     * (IMAD.WIDE.U32.X): 11 instructions per Philox round against 16 for a hand-written pass over
     * the four 32x32 products (cuobjdump, round 2) */
    *hi = __umul64hi(a, b);
    *lo = a * b;
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    *hi = (uint64_t)(p >> 64);
    *lo = (uint64_t)p;
#endif
}

/* 2^k for k in [-1022, 1023] */
APS_HD double aps_pow2i(int k) { return aps_bits2d((uint64_t)(k + 1023) << 52); }

/* ------------------------------------------------------------------ Philox2x64-10
 * Salmon et al. 2011; constants as in Random123 (multiplier M, Weyl key bump W).
 * ctr = (c0, c1), one 64-bit key; returns two 64-bit words.                                  */
#define APS_PHILOX_M 0xD2B74407B1CE6E93ULL
#define APS_PHILOX_W 0x9E3779B97F4A7C15ULL

APS_HD void aps_philox2x64(uint64_t c0, uint64_t c1, uint64_t key, uint64_t *o0, uint64_t *o1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint64_t hi, lo;
        aps_mul64(APS_PHILOX_M, c0, &hi, &lo);
        c0 = hi ^ key ^ c1;
        c1 = lo;
        key += APS_PHILOX_W;
    }
    *o0 = c0;
    *o1 = c1;
}

/* Counter layout used by every draw of a sweep (replaces the per-particle key tree of
 * src/rng.jl:38-42 + src/container.jl:126-159,202-215 by position/time-derived counters):
 *   c0 = global index (slot pair for state draws, child for resampler draws),
 *   c1 = step << 16 | domain << 8 | block.                                                    */
#define APS_DOM_STATE 0u    /* particle state draws (prior / transition)            */
#define APS_DOM_RESAMPLE 1u /* container stream: resampler uniforms                 */
#define APS_DOM_PGAS 2u     /* container stream: PGAS ancestor draw                 */
#define APS_DOM_PICK 3u     /* container stream: final trajectory pick              */
#define APS_DOM_DATA 4u     /* synthetic data simulation (bench / fixtures)         */

APS_HD uint64_t aps_ctr1(uint64_t step, uint32_t domain, uint32_t block) {
    return (step << 16) | ((uint64_t)domain << 8) | (uint64_t)block;
}

/* 53-bit integer uniform in [0, 2^53) */
APS_HD uint64_t aps_u53(uint64_t w) { return w >> 11; }
/* double in [0,1), multiple of 2^-53 */
APS_HD double aps_u01(uint64_t w) { return (double)(int64_t)(w >> 11) * 0x1.0p-53; }
/* double in (0,1): (k + 1/2) 2^-52, k = top 52 bits -- exactly representable */
APS_HD double aps_u01_open(uint64_t w) {
    return ((double)(int64_t)(w >> 12) + 0.5) * 0x1.0p-52;
}

/* ------------------------------------------------------------------ exp
 * k = round(x / ln 2), r = x - k ln2 (two-part), degree-13 Taylor in r (|r| <= 0.3466,
 * truncation < 5e-18), result scaled by 2^k in two steps so the subnormal range rounds once.  */
APS_HD double aps_exp(double x) {
    if (!(x == x)) return x;                 /* NaN */
    if (x > 709.782712893384) return aps_bits2d(0x7FF0000000000000ULL);
    if (x < -745.2) return 0.0;
    const double shifter = 0x1.8p52;
    double kd = aps_fma(x, 1.4426950408889634, shifter) - shifter;
    int k = (int)kd;
    double r = aps_fma(-kd, 0x1.62e42fee00000p-1, x);
    r = aps_fma(-kd, 0x1.a39ef35793c76p-33, r);
    double p = 1.6059043836821613e-10;
    p = aps_fma(p, r, 2.08767569878681e-09);
    p = aps_fma(p, r, 2.505210838544172e-08);
    p = aps_fma(p, r, 2.755731922398589e-07);
    p = aps_fma(p, r, 2.7557319223985893e-06);
    p = aps_fma(p, r, 2.48015873015873e-05);
    p = aps_fma(p, r, 0.0001984126984126984);
    p = aps_fma(p, r, 0.001388888888888889);
    p = aps_fma(p, r, 0.008333333333333333);
    p = aps_fma(p, r, 0.041666666666666664);
    p = aps_fma(p, r, 0.16666666666666666);
    p = aps_fma(p, r, 0.5);
    p = aps_fma(p, r, 1.0);
    p = aps_fma(p, r, 1.0);
    int k1 = k >> 1;
    int k2 = k - k1;
    return (p * aps_pow2i(k1)) * aps_pow2i(k2);
}

/* ------------------------------------------------------------------ log
 * x = 2^k m, m in [sqrt(1/2), sqrt(2)); f = m - 1; s = f / (2 + f); log(1+f) = 2 atanh(s)
 * with the classical 7-term minimax polynomial in s^2 (Remez coefficients as published with
 * the Sun/FreeBSD msun e_log.c algorithm).                                                    */
/* core: b = bits of a positive, finite, NORMAL double; k0 = exponent adjustment already applied */
APS_HD double aps_log_core(uint64_t b, int k) {
    k += (int)(b >> 52) - 1023;
    uint64_t mant = b & 0x000FFFFFFFFFFFFFULL;
    double m = aps_bits2d(mant | 0x3FF0000000000000ULL); /* [1,2) */
    if (mant > 0x6A09E667F3BCCULL) {                     /* m > sqrt(2) */
        m = m * 0.5;
        k += 1;
    }
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double R = 1.479819860511658591e-01;
    R = aps_fma(R, z, 1.531383769920937332e-01);
    R = aps_fma(R, z, 1.818357216161805012e-01);
    R = aps_fma(R, z, 2.222219843214978396e-01);
    R = aps_fma(R, z, 2.857142874366239149e-01);
    R = aps_fma(R, z, 3.999999999940941908e-01);
    R = aps_fma(R, z, 6.666666666666735130e-01);
    R = R * z;
    double hfsq = 0.5 * f * f;
    double dk = (double)k;
    /* log(x) = k ln2_hi - ((hfsq - (s (hfsq + R) + k ln2_lo)) - f) */
    double t = aps_fma(s, hfsq + R, dk * 0x1.a39ef35793c76p-33);
    return aps_fma(dk, 0x1.62e42fee00000p-1, -((hfsq - t) - f));
}
APS_HD double aps_log(double x) {
    if (!(x == x)) return x;
    if (x < 0.0) return aps_bits2d(0x7FF8000000000000ULL);
    if (x == 0.0) return aps_bits2d(0xFFF0000000000000ULL);
    uint64_t b = aps_d2bits(x);
    if (b >= 0x7FF0000000000000ULL) return x; /* +inf */
    int k = 0;
    if (b < 0x0010000000000000ULL) { /* subnormal */
        x = x * 0x1.0p54;
        b = aps_d2bits(x);
        k = -54;
    }
    return aps_log_core(b, k);
}
/* same value as aps_log(x) for a positive, finite, normal x (no special cases to test) */
APS_HD double aps_log_normal(double x) { return aps_log_core(aps_d2bits(x), 0); }

/* ------------------------------------------------------------------ sin(pi t), cos(pi t), t in [0, 2]
 * n = round(2t), r = t - n/2 in [-1/4, 1/4] (exact), Taylor in (pi r) with pre-multiplied
 * coefficients pi^j / j!, then quadrant rotation.                                             */
APS_HD void aps_sincospi(double t, double *sp, double *cp) {
    const double shifter = 0x1.8p52;
    double nd = (aps_fma(t, 2.0, shifter)) - shifter; /* round-to-nearest-even of 2t */
    int n = (int)nd;
    double r = aps_fma(nd, -0.5, t);
    double r2 = r * r;
    double s = 7.952054001475513e-07;
    s = aps_fma(s, r2, -2.1915353447830217e-05);
    s = aps_fma(s, r2, 0.00046630280576761255);
    s = aps_fma(s, r2, -0.0073704309457143504);
    s = aps_fma(s, r2, 0.08214588661112823);
    s = aps_fma(s, r2, -0.5992645293207921);
    s = aps_fma(s, r2, 2.5501640398773455);
    s = aps_fma(s, r2, -5.16771278004997);
    s = aps_fma(s, r2, 3.141592653589793);
    s = s * r;
    double c = -1.3878952462213771e-07;
    c = aps_fma(c, r2, 4.303069587032947e-06);
    c = aps_fma(c, r2, -0.0001046381049248457);
    c = aps_fma(c, r2, 0.0019295743094039231);
    c = aps_fma(c, r2, -0.02580689139001406);
    c = aps_fma(c, r2, 0.2353306303588932);
    c = aps_fma(c, r2, -1.3352627688545895);
    c = aps_fma(c, r2, 4.0587121264167685);
    c = aps_fma(c, r2, -4.934802200544679);
    c = aps_fma(c, r2, 1.0);
    /* quadrant rotation without branches: odd n swaps the two, the signs follow n (exact) */
    const double a = (n & 1) ? c : s, b = (n & 1) ? s : c;
    *sp = (n & 2) ? -a : a;
    *cp = ((n + 1) & 2) ? -b : b;
}

/* ------------------------------------------------------------------ standard normals
 * Box-Muller on one Philox block: u1 in (0,1) from word 0, u2 in [0,1) from word 1.
 * z0 = rho cos(2 pi u2), z1 = rho sin(2 pi u2), rho = sqrt(-2 log u1).                         */
APS_HD void aps_normal_pair(uint64_t w0, uint64_t w1, double *z0, double *z1) {
    double u1 = aps_u01_open(w0);
    double t = 2.0 * aps_u01(w1);
    double rho = aps_sqrt(-2.0 * aps_log_normal(u1)); /* u1 in [2^-53, 1): positive and normal */
    double s, c;
    aps_sincospi(t, &s, &c);
    *z0 = rho * c;
    *z1 = rho * s;
}

/* ------------------------------------------------------------------ weight quantisation
 * Canonical ("canon") weights are exact integers: q = floor(exp(logw - M) * 2^S) with
 * S = 62 - ceil(log2 N), so that sums over <= N particles fit in 62 bits and every prefix
 * sum is associative -> identical results for any tile size, launch order or GPU count.       */
APS_HD int aps_ceil_log2(uint64_t n) {
    int l = 0;
    while (((uint64_t)1 << l) < n) ++l;
    return l;
}
APS_HD int aps_weight_shift(uint64_t n) {
    int l = aps_ceil_log2(n < 2 ? 2 : n);
    return 62 - l;
}
/* shift h applied before squaring for the ESS sum so that sum (q>>h)^2 < 2^63 */
APS_HD int aps_ess_shift(uint64_t n) {
    int l = aps_ceil_log2(n < 2 ? 2 : n);
    int s = 62 - l;
    int keep = (63 - l) / 2;
    return s > keep ? s - keep : 0;
}
/* e in [0,1] -> integer weight; NaN -> 0 (caller flags the error separately) */
APS_HD uint64_t aps_quantise(double e, int shift) {
    if (!(e > 0.0)) return 0;
    double v = e * aps_pow2i(shift);
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double2ull_rz(v);
#else
    return (uint64_t)v;
#endif
}

/* order-preserving encoding of a double into uint64 (for atomicMax reductions) */
APS_HD uint64_t aps_encode_ordered(double x) {
    uint64_t b = aps_d2bits(x);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
APS_HD double aps_decode_ordered(uint64_t e) {
    uint64_t b = (e & 0x8000000000000000ULL) ? (e & 0x7FFFFFFFFFFFFFFFULL) : ~e;
    return aps_bits2d(b);
}

#endif /* APS_MATH_H */
