/*
 * aps_oracle.cpp -- CPU oracle for the SMC / particle-MCMC hot path. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (advancedps.jl_b200/) never does.
 *
 * It is a single-threaded, sequential restatement of the reference's algorithm
 * (TuringLang/AdvancedPS.jl v0.7.2, paths relative to /root/reference), function by function:
 *   sweep                      src/container.jl:316-363  (T+1 rounds of resample -> logZ0 -> reweight -> logZ1)
 *   resample_propagate (ESS)   src/container.jl:233-251, src/resampling.jl:193-204
 *   resample_propagate (core)  src/container.jl:171-231  (index draw, histogram, children in parent order,
 *                                                        reference in the last slot, log-weights reset)
 *   reweight / advance         src/container.jl:259-302, src/pgas.jl:53-89
 *   logZ / getweights / ESS    src/container.jl:95-119
 *   resample_*                 src/resampling.jl:11-21,31-35,53-81,98-131,149-183
 *   PGAS update_ref            src/pgas.jl:26-46,113-128 (incl. its c-2 / c-1 index convention)
 *   final pick                 src/container.jl:33-36, src/smc.jl:127
 * The reference deep-copies whole trajectories on fork (src/pgas.jl:99-104); the oracle stores
 * per-step states and ancestor indices instead, which is observationally the same.
 *
 * Two arithmetic modes:
 *   SEQ   - the reference's floating-point order: fp64 softmax, sequentially rounded running sum
 *           `v += n*w[j]`, `u += 1.0` (src/resampling.jl:157-179), ess = 1/sum(w^2).
 *   CANON - exact-integer weights q = floor(exp(logw - max) 2^S) (aps_math.h); all sums are
 *           integers, thresholds are compared in exact rational arithmetic. This is the mode the
 *           CUDA path reproduces bit-for-bit; SEQ-vs-CANON differences are measured in tests.
 *
 * PARITY PINNING. The reference cannot run here (no Julia). The oracle is pinned on every
 * RNG-independent known answer the reference's tests hold for this path (test/container.jl:45-68,
 * 86-99,110-119; test/smc.jl:104; test/pgas.jl:83-87; test/resampling.jl:12-15) and on closed-form
 * Kalman log-likelihoods. RNG bit-streams are NOT pinned to Julia's: the reference draws through
 * Random123.jl / Random.randn / MersenneTwister-based key splitting (src/rng.jl:38-42), none of
 * which is vendored; here every draw is Philox2x64-10 (pinned to Random123's published KATs) at
 * position/time-derived counters, with Box-Muller normals (one block per pair of slots).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <algorithm>
#include <thread>

#include "../include/aps_b200.h"

typedef unsigned __int128 u128;

/* Threads. The reference sweep is serial (src/container.jl:194,264) and so is this oracle by
 * default. orc_set_threads(n > 1) lets the particle-parallel loops of the CANON mode (advance,
 * quantise, integer sums, maxima) run on n threads: every one of them is either independent per
 * particle or an exact integer / max reduction over per-thread partials, so results are
 * bit-identical for any n. It exists only to give bench.py an all-cores CPU figure next to the
 * single-thread one; the SEQ mode (sequential fp64 sums, the reference's order) and the
 * resampling walks stay serial. (std::thread: the image has no OpenMP runtime.)                 */
static int g_threads = 1;
extern "C" void orc_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
extern "C" int orc_max_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}
/* f(lo, hi, k): chunk k of [0, n) */
template <class F>
static void pfor(int64_t n, F f) {
    const int nt = (g_threads > 1 && n >= 4096) ? g_threads : 1;
    if (nt == 1) {
        f((int64_t)0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    const int64_t per = (n + nt - 1) / nt;
    for (int k = 0; k < nt; ++k) {
        const int64_t lo = (int64_t)k * per, hi = std::min(n, lo + per);
        if (lo < hi) th.emplace_back([=] { f(lo, hi, k); });
    }
    for (auto &t : th) t.join();
}

enum { ORC_SEQ = 0, ORC_CANON = 1 };

extern "C" {

/* ------------------------------------------------------------------ math wrappers (for tests) */
void orc_philox2x64(uint64_t c0, uint64_t c1, uint64_t key, uint64_t *out) {
    aps_philox2x64(c0, c1, key, &out[0], &out[1]);
}
double orc_exp(double x) { return aps_exp(x); }
double orc_log(double x) { return aps_log(x); }
void orc_sincospi(double t, double *s, double *c) { aps_sincospi(t, s, c); }
void orc_normal_pair(uint64_t w0, uint64_t w1, double *z) { aps_normal_pair(w0, w1, &z[0], &z[1]); }
double orc_u01(uint64_t w) { return aps_u01(w); }
int orc_weight_shift(int64_t n) { return aps_weight_shift((uint64_t)n); }
int orc_ess_shift(int64_t n) { return aps_ess_shift((uint64_t)n); }

/* ------------------------------------------------------------------ weights (src/container.jl:95-119) */
static double max_of(const double *x, int64_t n) {
    double part[256];
    int nanp[256];
    for (int k = 0; k < 256; ++k) { part[k] = -INFINITY; nanp[k] = 0; }
    pfor(n, [&](int64_t lo, int64_t hi, int k) {
        double m = -INFINITY;
        int nan = 0;
        for (int64_t i = lo; i < hi; ++i) {
            if (x[i] != x[i]) nan = 1;
            else if (x[i] > m) m = x[i];
        }
        part[k] = m;
        nanp[k] = nan;
    });
    double m = -INFINITY;
    for (int k = 0; k < 256; ++k) {
        if (nanp[k]) return NAN;
        if (part[k] > m) m = part[k];
    }
    return m;
}

/* canonical integer weights of a log-weight vector; returns 0, or 2 if not normalisable */
int orc_quantise_logw(const double *logw, int64_t n, uint64_t *q, double *max_out, uint64_t *total_out) {
    double m = max_of(logw, n);
    *max_out = m;
    if (n <= 0) return 1;
    if (m != m || m == -INFINITY || m == INFINITY) return 2;
    int S = aps_weight_shift((uint64_t)n);
    uint64_t Q = 0;
    for (int64_t i = 0; i < n; ++i) {
        q[i] = aps_quantise(aps_exp(logw[i] - m), S);
        Q += q[i];
    }
    *total_out = Q;
    return Q == 0 ? 2 : 0;
}

/* canonical integer weights of a SHARD: maximum and particle count are the global ones */
int orc_quantise_shard(const double *logw, int64_t n_local, double global_max, int64_t n_global, uint64_t *q,
                       uint64_t *total_out) {
    int S = aps_weight_shift((uint64_t)n_global);
    uint64_t Q = 0;
    for (int64_t i = 0; i < n_local; ++i) {
        q[i] = aps_quantise(aps_exp(logw[i] - global_max), S);
        Q += q[i];
    }
    *total_out = Q;
    return 0;
}

/* canonical integer weights of a (not necessarily normalised) non-negative weight vector */
int orc_quantise_w(const double *w, int64_t n, uint64_t *q, uint64_t *total_out) {
    if (n <= 0) return 1;
    double m = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (!(w[i] >= 0.0)) return 2; /* negative or NaN */
        if (w[i] > m) m = w[i];
    }
    if (!(m > 0.0) || m == INFINITY) return 2;
    int S = aps_weight_shift((uint64_t)n);
    uint64_t Q = 0;
    for (int64_t i = 0; i < n; ++i) {
        q[i] = aps_quantise(w[i] / m, S);
        Q += q[i];
    }
    *total_out = Q;
    return Q == 0 ? 2 : 0;
}

static double canon_logsum(uint64_t Q, int S) { return aps_log((double)Q * aps_pow2i(-S)); }

static double canon_ess(const uint64_t *q, int64_t n) {
    int h = aps_ess_shift((uint64_t)n);
    uint64_t p1[256] = {0}, p2[256] = {0};
    pfor(n, [&](int64_t lo, int64_t hi, int k) {
        uint64_t a = 0, b = 0;
        for (int64_t i = lo; i < hi; ++i) {
            uint64_t t = q[i] >> h;
            a += t;
            b += t * t;
        }
        p1[k] = a;
        p2[k] = b;
    });
    uint64_t s1 = 0, s2 = 0;
    for (int k = 0; k < 256; ++k) { s1 += p1[k]; s2 += p2[k]; }
    return ((double)s1 * (double)s1) / (double)s2;
}

int orc_logsumexp(const double *logw, int64_t n, int mode, double *out) {
    if (n <= 0) return 1;
    double m = max_of(logw, n);
    if (m != m) return 2;
    if (m == -INFINITY) { *out = -INFINITY; return 0; }
    if (mode == ORC_SEQ) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += aps_exp(logw[i] - m);
        *out = m + aps_log(s);
        return 0;
    }
    std::vector<uint64_t> q((size_t)n);
    uint64_t Q;
    double mm;
    int rc = orc_quantise_logw(logw, n, q.data(), &mm, &Q);
    if (rc) return rc;
    *out = m + canon_logsum(Q, aps_weight_shift((uint64_t)n));
    return 0;
}

int orc_softmax(const double *logw, int64_t n, int mode, double *w) {
    if (n <= 0) return 1;
    double m = max_of(logw, n);
    if (m != m || m == -INFINITY) return 2;
    if (mode == ORC_SEQ) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) { w[i] = aps_exp(logw[i] - m); s += w[i]; }
        double inv = 1.0 / s;
        for (int64_t i = 0; i < n; ++i) w[i] *= inv;
        return 0;
    }
    std::vector<uint64_t> q((size_t)n);
    uint64_t Q;
    double mm;
    int rc = orc_quantise_logw(logw, n, q.data(), &mm, &Q);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) w[i] = (double)q[i] / (double)Q;
    return 0;
}

int orc_ess(const double *logw, int64_t n, int mode, double *out) {
    if (n <= 0) return 1;
    if (mode == ORC_SEQ) {
        std::vector<double> w((size_t)n);
        int rc = orc_softmax(logw, n, ORC_SEQ, w.data());
        if (rc) return rc;
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += w[i] * w[i];
        *out = 1.0 / s;
        return 0;
    }
    std::vector<uint64_t> q((size_t)n);
    uint64_t Q;
    double mm;
    int rc = orc_quantise_logw(logw, n, q.data(), &mm, &Q);
    if (rc) return rc;
    *out = canon_ess(q.data(), n);
    return 0;
}

/* ------------------------------------------------------------------ SEQ resamplers: literal restatements
 * of src/resampling.jl on fp64 weights with the uniforms passed in. Output indices are 1-based. */

/* src/resampling.jl:11-21 */
int64_t orc_randcat_seq(const double *p, int64_t n, double r) {
    double cp = p[0];
    int64_t s = 1;
    while (cp <= r && s < n) {
        s += 1;
        cp += p[s - 1];
    }
    return s;
}

/* src/resampling.jl:149-183; returns 0 ok, 1 empty, 2 "sample could not be selected" */
int orc_resample_systematic_seq(const double *w, int64_t m, int64_t n, double u0, int64_t *out) {
    if (m <= 0) return 1;
    double v = (double)n * w[0];
    double u = u0;
    int64_t sample = 1;
    for (int64_t i = 0; i < n; ++i) {
        while (v < u) {
            sample += 1;
            if (sample > m) return 2;
            v += (double)n * w[sample - 1];
        }
        out[i] = sample;
        u += 1.0;
    }
    return 0;
}

/* src/resampling.jl:98-131; us[i] is the fresh rand() of iteration i */
int orc_resample_stratified_seq(const double *w, int64_t m, int64_t n, const double *us, int64_t *out) {
    if (m <= 0) return 1;
    double v = (double)n * w[0];
    int64_t sample = 1;
    for (int64_t i = 0; i < n; ++i) {
        double u = (double)i + us[i]; /* i - 1 + rand(rng) with 1-based i */
        while (v < u) {
            sample += 1;
            if (sample > m) return 2;
            v += (double)n * w[sample - 1];
        }
        out[i] = sample;
    }
    return 0;
}

/* src/resampling.jl:31-35: n i.i.d. categorical draws. Upstream draws through Distributions'
 * alias table (third-party, not in the reference tree); restated as inverse-CDF draws with the
 * randcat walk of :11-21, which has the same law. Unsorted, like the reference.                */
int orc_resample_multinomial_seq(const double *w, int64_t m, int64_t n, const double *us, int64_t *out) {
    if (m <= 0) return 1;
    std::vector<double> cdf((size_t)m);
    double c = 0.0;
    for (int64_t j = 0; j < m; ++j) { c += w[j]; cdf[(size_t)j] = c; }
    for (int64_t i = 0; i < n; ++i) {
        /* first j with cdf[j] > u  == randcat's `while cp <= r && s < n` */
        int64_t j = std::upper_bound(cdf.begin(), cdf.end(), us[i]) - cdf.begin();
        if (j >= m) j = m - 1;
        out[i] = j + 1;
    }
    return 0;
}

/* src/resampling.jl:53-81, documented intent (:43-51): floor(n w_j) copies of j in order, then
 * the remaining slots i.i.d. from the normalised residuals. (Upstream's residual branch calls an
 * unimported `rand!`) us has one uniform per residual slot.              */
int orc_resample_residual_seq(const double *w, int64_t m, int64_t n, const double *us, int64_t *out) {
    if (m <= 0) return 1;
    std::vector<double> res((size_t)m);
    int64_t i = 0;
    for (int64_t j = 0; j < m; ++j) {
        double x = (double)n * w[j];
        int64_t fl = (int64_t)floor(x);
        for (int64_t k = 0; k < fl && i < n; ++k) out[i++] = j + 1;
        res[(size_t)j] = x - (double)fl;
    }
    if (i < n) {
        double s = 0.0;
        for (int64_t j = 0; j < m; ++j) s += res[(size_t)j];
        for (int64_t j = 0; j < m; ++j) res[(size_t)j] /= s;
        int64_t r = 0;
        for (; i < n; ++i, ++r) out[i] = orc_randcat_seq(res.data(), m, us[r]);
    }
    return 0;
}

/* ------------------------------------------------------------------ CANON resamplers on integer weights.
 * Same walks as above with exact arithmetic: parent j is selected for child i iff
 *   C_j * n >= i * Q + R_i,  C_j = q_1 + ... + q_j,  Q = C_m,  R_i = ceil(U_i Q / 2^53),
 * i.e. NOT (v < u) of src/resampling.jl:165 evaluated without rounding. Uniforms come from
 * Philox2x64-10: U_i = top 53 bits of word 0 of block (i, ctr1(step, DOM_RESAMPLE, 0)).          */
static uint64_t draw_u53(uint64_t key, uint64_t idx, uint64_t step, uint32_t dom) {
    uint64_t w0, w1;
    aps_philox2x64(idx, aps_ctr1(step, dom, 0), key, &w0, &w1);
    return aps_u53(w0);
}
/* i.i.d. draws (multinomial, residual): draw i is word (i & 1) of block i >> 1 -- one Philox
 * block serves two draws */
static uint64_t draw_iid_u53(uint64_t key, uint64_t i, uint64_t step, uint32_t dom) {
    uint64_t w[2];
    aps_philox2x64(i >> 1, aps_ctr1(step, dom, 0), key, &w[0], &w[1]);
    return aps_u53(w[i & 1]);
}
static uint64_t ceil_uq(uint64_t U, uint64_t Q) { /* ceil(U Q / 2^53) */
    u128 p = (u128)U * Q + (((u128)1 << 53) - 1);
    return (uint64_t)(p >> 53);
}
static uint64_t floor_uq(uint64_t U, uint64_t Q) { /* floor(U Q / 2^53) in [0, Q) */
    return (uint64_t)(((u128)U * Q) >> 53);
}

int orc_resample_systematic_canon(const uint64_t *q, int64_t m, int64_t n, uint64_t key, uint64_t step,
                                  int64_t *out) {
    if (m <= 0) return 1;
    uint64_t Q = 0;
    for (int64_t j = 0; j < m; ++j) Q += q[j];
    if (Q == 0) return 2;
    uint64_t R = ceil_uq(draw_u53(key, 0, step, APS_DOM_RESAMPLE), Q);
    uint64_t C = q[0];
    int64_t sample = 1;
    for (int64_t i = 0; i < n; ++i) {
        u128 thr = (u128)(uint64_t)i * Q + R;
        while ((u128)C * (uint64_t)n < thr) {
            sample += 1;
            if (sample > m) return 2;
            C += q[sample - 1];
        }
        out[i] = sample;
    }
    return 0;
}

int orc_resample_stratified_canon(const uint64_t *q, int64_t m, int64_t n, uint64_t key, uint64_t step,
                                  int64_t *out) {
    if (m <= 0) return 1;
    uint64_t Q = 0;
    for (int64_t j = 0; j < m; ++j) Q += q[j];
    if (Q == 0) return 2;
    uint64_t C = q[0];
    int64_t sample = 1;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t R = ceil_uq(draw_u53(key, (uint64_t)i, step, APS_DOM_RESAMPLE), Q);
        u128 thr = (u128)(uint64_t)i * Q + R;
        while ((u128)C * (uint64_t)n < thr) {
            sample += 1;
            if (sample > m) return 2;
            C += q[sample - 1];
        }
        out[i] = sample;
    }
    return 0;
}

/* categorical draw on integer weights: first j with C_j > tau, tau = floor(U Q / 2^53) */
static int64_t canon_categorical(const std::vector<uint64_t> &cum, uint64_t U) {
    uint64_t Q = cum.back();
    uint64_t tau = floor_uq(U, Q);
    int64_t j = std::upper_bound(cum.begin(), cum.end(), tau) - cum.begin();
    if (j >= (int64_t)cum.size()) j = (int64_t)cum.size() - 1;
    return j; /* 0-based */
}

int orc_resample_multinomial_canon(const uint64_t *q, int64_t m, int64_t n, uint64_t key, uint64_t step,
                                   int64_t *out) {
    if (m <= 0) return 1;
    std::vector<uint64_t> cum((size_t)m);
    uint64_t C = 0;
    for (int64_t j = 0; j < m; ++j) { C += q[j]; cum[(size_t)j] = C; }
    if (C == 0) return 2;
    for (int64_t i = 0; i < n; ++i)
        out[i] = canon_categorical(cum, draw_iid_u53(key, (uint64_t)i, step, APS_DOM_RESAMPLE)) + 1;
    return 0;
}

/* residual: d_j = floor(n q_j / Q) copies; residual r_j = n q_j - d_j Q; remaining Rc = n - sum d_j
 * slots i.i.d. with integer weights r_j >> ceil_log2(Rc + 1) (so their sum fits 62 bits).        */
int orc_resample_residual_canon(const uint64_t *q, int64_t m, int64_t n, uint64_t key, uint64_t step,
                                int64_t *out) {
    if (m <= 0) return 1;
    uint64_t Q = 0;
    for (int64_t j = 0; j < m; ++j) Q += q[j];
    if (Q == 0) return 2;
    std::vector<uint64_t> res((size_t)m);
    int64_t i = 0;
    for (int64_t j = 0; j < m; ++j) {
        u128 x = (u128)q[j] * (uint64_t)n;
        uint64_t d = (uint64_t)(x / Q);
        res[(size_t)j] = (uint64_t)(x - (u128)d * Q);
        for (uint64_t k = 0; k < d && i < n; ++k) out[i++] = j + 1;
    }
    int64_t Rc = n - i;
    if (Rc > 0) {
        int sh = aps_ceil_log2((uint64_t)Rc + 1);
        std::vector<uint64_t> cum((size_t)m);
        uint64_t C = 0;
        for (int64_t j = 0; j < m; ++j) { C += res[(size_t)j] >> sh; cum[(size_t)j] = C; }
        if (C == 0) return 2;
        for (int64_t r = 0; r < Rc; ++r, ++i)
            out[i] = canon_categorical(cum, draw_iid_u53(key, (uint64_t)r, step, APS_DOM_RESAMPLE)) + 1;
    }
    return 0;
}

/* operator-level entry: (kind, fp64 weights) -> 1-based indices, mode CANON or SEQ with Philox uniforms */
int orc_resample(int kind, int mode, const double *w, int64_t m, int64_t n, uint64_t key, uint64_t step,
                 int64_t *out) {
    if (m <= 0 || n < 0) return 1;
    if (mode == ORC_CANON) {
        std::vector<uint64_t> q((size_t)m);
        uint64_t Q;
        int rc = orc_quantise_w(w, m, q.data(), &Q);
        if (rc) return rc;
        switch (kind) {
            case APS_RESAMPLE_SYSTEMATIC: return orc_resample_systematic_canon(q.data(), m, n, key, step, out);
            case APS_RESAMPLE_STRATIFIED: return orc_resample_stratified_canon(q.data(), m, n, key, step, out);
            case APS_RESAMPLE_MULTINOMIAL: return orc_resample_multinomial_canon(q.data(), m, n, key, step, out);
            case APS_RESAMPLE_RESIDUAL: return orc_resample_residual_canon(q.data(), m, n, key, step, out);
        }
        return 1;
    }
    if (kind == APS_RESAMPLE_SYSTEMATIC)
        return orc_resample_systematic_seq(w, m, n, aps_u01(draw_u53(key, 0, step, APS_DOM_RESAMPLE) << 11), out);
    std::vector<double> us((size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i)
        us[(size_t)i] = (double)(int64_t)(kind == APS_RESAMPLE_STRATIFIED ? draw_u53(key, (uint64_t)i, step, APS_DOM_RESAMPLE)
                                                                          : draw_iid_u53(key, (uint64_t)i, step, APS_DOM_RESAMPLE)) * 0x1.0p-53;
    switch (kind) {
        case APS_RESAMPLE_STRATIFIED: return orc_resample_stratified_seq(w, m, n, us.data(), out);
        case APS_RESAMPLE_MULTINOMIAL: return orc_resample_multinomial_seq(w, m, n, us.data(), out);
        case APS_RESAMPLE_RESIDUAL: return orc_resample_residual_seq(w, m, n, us.data(), out);
    }
    return 1;
}

/* randcat on fp64 weights; canon = integer categorical */
int orc_randcat(const double *w, int64_t n, int mode, uint64_t key, uint64_t step, int64_t *out) {
    if (n <= 0) return 1;
    uint64_t U = draw_u53(key, 0, step, APS_DOM_RESAMPLE);
    if (mode == ORC_SEQ) {
        *out = orc_randcat_seq(w, n, (double)(int64_t)U * 0x1.0p-53);
        return 0;
    }
    std::vector<uint64_t> q((size_t)n), cum((size_t)n);
    uint64_t Q;
    int rc = orc_quantise_w(w, n, q.data(), &Q);
    if (rc) return rc;
    uint64_t C = 0;
    for (int64_t j = 0; j < n; ++j) { C += q[j]; cum[(size_t)j] = C; }
    *out = canon_categorical(cum, U) + 1;
    return 0;
}

} /* extern "C" */

/* ================================================================== the sweep */
namespace {

struct Sweep {
    aps_config cfg;
    aps_model_dev md;
    int64_t N, T;
    int d, dy, mode;
    uint64_t key;
    const double *Y;
    const double *ref; /* T x d or NULL */
    std::vector<double> logw;
    std::vector<uint64_t> q;
    std::vector<double> w; /* SEQ: normalised fp64 weights */
    double M;
    uint64_t Q;
    int S;
    /* outputs */
    double *x_hist;     /* T x N x d */
    int32_t *anc_hist;  /* (T+1) x N : slab s-1 holds ancestors used to build set s, s = 1..T+1 */
    double *logz, *ess;
    uint8_t *resampled;
    int err;
};

template <int D, int DY, int OBS>
void advance_all(Sweep &sw, int64_t t) { /* reweight!: src/container.jl:259-302 -> advance!: src/pgas.jl:53-89 */
    const int64_t N = sw.N;
    const bool hasref = sw.ref != nullptr;
    const double *y = sw.Y + (size_t)(t - 1) * sw.dy;
    double *xt = sw.x_hist + (size_t)(t - 1) * N * D;
    const double *xp_all = t > 1 ? sw.x_hist + (size_t)(t - 2) * N * D : nullptr;
    const int32_t *anc = sw.anc_hist + (size_t)(t - 1) * N;
    pfor(N, [&](int64_t lo_i, int64_t hi_i, int) {
    for (int64_t i = lo_i; i < hi_i; ++i) {
        double x[D];
        if (hasref && i == N - 1) {
            for (int k = 0; k < D; ++k) x[k] = sw.ref[(size_t)(t - 1) * D + k]; /* pgas.jl:69-72 */
        } else {
            double zz[2 * D];
            aps_pair_normals<D>(sw.key, (uint64_t)(i >> 1), (uint64_t)t, zz);
            const double *z = zz + (i & 1) * D;
            if (t == 1) {
                aps_prior_draw<D>(&sw.md, z, x);
            } else {
                const double *xp = xp_all + (size_t)anc[i] * D;
                aps_trans_draw<D>(&sw.md, xp, z, x);
            }
        }
        for (int k = 0; k < D; ++k) xt[(size_t)i * D + k] = x[k];
        sw.logw[(size_t)i] += aps_obs_logpdf<D, DY, OBS>(&sw.md, x, y); /* increase_logweight!, container.jl:279 */
    }
    });
}

typedef void (*advance_fn)(Sweep &, int64_t);
template <int D, int OBS>
advance_fn pick_adv_dy(int dy) {
    switch (dy) {
        case 1: return advance_all<D, 1, OBS>;
        case 2: return advance_all<D, 2, OBS>;
        case 3: return advance_all<D, 3, OBS>;
        default: return advance_all<D, 4, OBS>;
    }
}
template <int OBS>
advance_fn pick_adv(int d, int dy) {
    switch (d) {
        case 1: return pick_adv_dy<1, OBS>(dy);
        case 2: return pick_adv_dy<2, OBS>(dy);
        case 3: return pick_adv_dy<3, OBS>(dy);
        default: return pick_adv_dy<4, OBS>(dy);
    }
}

template <int D>
double trans_lp(const aps_model_dev *md, const double *xp, const double *xn) {
    return aps_trans_logpdf<D>(md, xp, xn);
}

/* refresh M, q, Q (CANON) or w (SEQ) from logw; returns ESS */
double refresh_weights(Sweep &sw) {
    const int64_t N = sw.N;
    double m = max_of(sw.logw.data(), N);
    sw.M = m;
    if (m != m || m == -INFINITY || m == INFINITY) { sw.err = 2; return NAN; }
    if (sw.mode == ORC_CANON) {
        uint64_t part[256] = {0};
        pfor(N, [&](int64_t lo, int64_t hi, int k) {
            uint64_t a = 0;
            for (int64_t i = lo; i < hi; ++i) {
                sw.q[(size_t)i] = aps_quantise(aps_exp(sw.logw[(size_t)i] - m), sw.S);
                a += sw.q[(size_t)i];
            }
            part[k] = a;
        });
        uint64_t Q = 0;
        for (int k = 0; k < 256; ++k) Q += part[k];
        sw.Q = Q;
        if (Q == 0) { sw.err = 2; return NAN; }
        return canon_ess(sw.q.data(), N);
    }
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) { sw.w[(size_t)i] = aps_exp(sw.logw[(size_t)i] - m); s += sw.w[(size_t)i]; }
    double inv = 1.0 / s, s2 = 0.0;
    for (int64_t i = 0; i < N; ++i) { sw.w[(size_t)i] *= inv; s2 += sw.w[(size_t)i] * sw.w[(size_t)i]; }
    return 1.0 / s2;
}

double current_logZ(Sweep &sw) { /* logZ(pc), src/container.jl:109 */
    const int64_t N = sw.N;
    double m = max_of(sw.logw.data(), N);
    if (m != m || m == -INFINITY) { sw.err = 2; return NAN; }
    if (sw.mode == ORC_CANON) {
        uint64_t part[256] = {0};
        pfor(N, [&](int64_t lo, int64_t hi, int k) {
            uint64_t a = 0;
            for (int64_t i = lo; i < hi; ++i) a += aps_quantise(aps_exp(sw.logw[(size_t)i] - m), sw.S);
            part[k] = a;
        });
        uint64_t Q = 0;
        for (int k = 0; k < 256; ++k) Q += part[k];
        return m + canon_logsum(Q, sw.S);
    }
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) s += aps_exp(sw.logw[(size_t)i] - m);
    return m + aps_log(s);
}

/* one categorical draw over lw (PGAS ancestor / final pick) */
int64_t categorical_logw(Sweep &sw, const double *lw, int64_t n, uint64_t U) {
    double m = max_of(lw, n);
    if (m != m || m == -INFINITY) { sw.err = 2; return 0; }
    if (sw.mode == ORC_CANON) {
        std::vector<uint64_t> cum((size_t)n);
        uint64_t C = 0;
        for (int64_t i = 0; i < n; ++i) { C += aps_quantise(aps_exp(lw[i] - m), sw.S); cum[(size_t)i] = C; }
        if (C == 0) { sw.err = 2; return 0; }
        return canon_categorical(cum, U);
    }
    std::vector<double> p((size_t)n);
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) { p[(size_t)i] = aps_exp(lw[i] - m); s += p[(size_t)i]; }
    double inv = 1.0 / s;
    for (int64_t i = 0; i < n; ++i) p[(size_t)i] *= inv;
    return orc_randcat_seq(p.data(), n, (double)(int64_t)U * 0x1.0p-53) - 1;
}

/* resample_propagate! core (src/container.jl:171-231) at decision point s: fills ancestors of set s+1 */
void resample_core(Sweep &sw, int64_t s) {
    const int64_t N = sw.N;
    const bool hasref = sw.ref != nullptr;
    const int64_t n = hasref ? N - 1 : N; /* :181 */
    int32_t *anc_next = sw.anc_hist + (size_t)s * N;
    std::vector<int64_t> indx((size_t)(n > 0 ? n : 1));
    int rc = 0;
    const int kind = sw.cfg.resampler;
    if (sw.mode == ORC_CANON) {
        switch (kind) {
            case APS_RESAMPLE_SYSTEMATIC: rc = orc_resample_systematic_canon(sw.q.data(), N, n, sw.key, (uint64_t)s, indx.data()); break;
            case APS_RESAMPLE_STRATIFIED: rc = orc_resample_stratified_canon(sw.q.data(), N, n, sw.key, (uint64_t)s, indx.data()); break;
            case APS_RESAMPLE_MULTINOMIAL: rc = orc_resample_multinomial_canon(sw.q.data(), N, n, sw.key, (uint64_t)s, indx.data()); break;
            default: rc = orc_resample_residual_canon(sw.q.data(), N, n, sw.key, (uint64_t)s, indx.data()); break;
        }
    } else {
        if (kind == APS_RESAMPLE_SYSTEMATIC) {
            double u0 = (double)(int64_t)draw_u53(sw.key, 0, (uint64_t)s, APS_DOM_RESAMPLE) * 0x1.0p-53;
            rc = orc_resample_systematic_seq(sw.w.data(), N, n, u0, indx.data());
        } else {
            std::vector<double> us((size_t)(n > 0 ? n : 1));
            for (int64_t i = 0; i < n; ++i)
                us[(size_t)i] = (double)(int64_t)(kind == APS_RESAMPLE_STRATIFIED
                                                      ? draw_u53(sw.key, (uint64_t)i, (uint64_t)s, APS_DOM_RESAMPLE)
                                                      : draw_iid_u53(sw.key, (uint64_t)i, (uint64_t)s, APS_DOM_RESAMPLE)) * 0x1.0p-53;
            if (kind == APS_RESAMPLE_STRATIFIED) rc = orc_resample_stratified_seq(sw.w.data(), N, n, us.data(), indx.data());
            else if (kind == APS_RESAMPLE_MULTINOMIAL) rc = orc_resample_multinomial_seq(sw.w.data(), N, n, us.data(), indx.data());
            else rc = orc_resample_residual_seq(sw.w.data(), N, n, us.data(), indx.data());
        }
    }
    if (rc) { sw.err = rc; return; }
    /* count children (:185-188), emit grouped by parent in increasing parent order (:194-217) */
    std::vector<int32_t> nchild((size_t)N, 0);
    for (int64_t i = 0; i < n; ++i) nchild[(size_t)(indx[(size_t)i] - 1)] += 1;
    int64_t j = 0;
    for (int64_t i = 0; i < N; ++i)
        for (int32_t k = 0; k < nchild[(size_t)i]; ++k) anc_next[j++] = (int32_t)i;
    if (hasref) {
        /* update_ref! (src/pgas.jl:113-128) then children[n] = ref (:219-224) */
        int32_t a = (int32_t)(N - 1);
        const int64_t c = s + 1; /* reference particle's step counter */
        if (sw.cfg.sampler == APS_PGAS && c > 2 && c <= sw.T) {
            const int D = sw.d;
            const double *xref = sw.ref + (size_t)(c - 2) * D;             /* X_ref[c-1]           */
            const double *xpp = sw.x_hist + (size_t)(c - 3) * N * D;        /* states of time c-2   */
            const int32_t *anc_cur = sw.anc_hist + (size_t)(c - 2) * N;     /* ancestors of set c-1 */
            std::vector<double> lw((size_t)N);
            for (int64_t i = 0; i < N; ++i) {
                const double *xp = xpp + (size_t)anc_cur[i] * D;           /* X_i[c-2]             */
                double lp;
                switch (D) {
                    case 1: lp = trans_lp<1>(&sw.md, xp, xref); break;
                    case 2: lp = trans_lp<2>(&sw.md, xp, xref); break;
                    case 3: lp = trans_lp<3>(&sw.md, xp, xref); break;
                    default: lp = trans_lp<4>(&sw.md, xp, xref); break;
                }
                lw[(size_t)i] = lp + sw.logw[(size_t)i];                    /* pgas.jl:43 */
            }
            uint64_t U = draw_u53(sw.key, 0, (uint64_t)s, APS_DOM_PGAS);
            a = (int32_t)categorical_logw(sw, lw.data(), N, U);
        }
        anc_next[N - 1] = a;
    }
    std::fill(sw.logw.begin(), sw.logw.end(), 0.0); /* reset_logweights!, :228 */
}

} /* namespace */

extern "C" {

/* Full sweep. Outputs (any may be NULL except logevidence):
 *   x_hist T*N*d ([t-1][i][k]); anc_hist (T+1)*N ([s][i], ancestors of set s+1, 0-based);
 *   logz T (logZ1 after observing y_t); ess T+1, resampled T+1 (decision points s = 0..T);
 *   final_logw N; final_w N (normalised weights of the final set).                             */
int orc_sweep(const aps_config *cfg, const double *Y, uint64_t seed, const double *ref_traj, int mode,
              double *logevidence, double *x_hist, int32_t *anc_hist, double *logz, double *ess,
              uint8_t *resampled, double *final_logw, double *final_w) {
    Sweep sw;
    sw.cfg = *cfg;
    if (aps_model_prepare(&cfg->model, &sw.md)) return 1;
    sw.N = cfg->n_particles;
    sw.T = cfg->n_steps;
    sw.d = cfg->model.d;
    sw.dy = cfg->model.dy;
    sw.mode = mode;
    sw.key = seed;
    sw.Y = Y;
    sw.ref = (cfg->sampler == APS_SMC) ? nullptr : ref_traj;
    sw.err = 0;
    const int64_t N = sw.N, T = sw.T;
    if (N < 1 || T < 1) return 1;
    if (sw.ref && N < 2) return 1;
    sw.S = aps_weight_shift((uint64_t)N);
    sw.logw.assign((size_t)N, 0.0);
    sw.q.assign((size_t)N, 0);
    sw.w.assign((size_t)N, 0.0);
    std::vector<double> xbuf;
    std::vector<int32_t> abuf;
    std::vector<double> zbuf((size_t)T + 1), ebuf((size_t)T + 1);
    std::vector<uint8_t> rbuf((size_t)T + 1);
    if (!x_hist) { xbuf.resize((size_t)T * N * sw.d); x_hist = xbuf.data(); }
    if (!anc_hist) { abuf.resize((size_t)(T + 1) * N); anc_hist = abuf.data(); }
    sw.x_hist = x_hist;
    sw.anc_hist = anc_hist;
    sw.logz = logz ? logz : zbuf.data();
    sw.ess = ess ? ess : ebuf.data();
    sw.resampled = resampled ? resampled : rbuf.data();

    advance_fn adv;
    switch (cfg->model.obs_kind) {
        case APS_OBS_LINEAR_GAUSS: adv = pick_adv<APS_OBS_LINEAR_GAUSS>(sw.d, sw.dy); break;
        case APS_OBS_STOCH_VOL: adv = pick_adv<APS_OBS_STOCH_VOL>(sw.d, 1); break;
        default: adv = pick_adv<APS_OBS_CONST>(sw.d, 1); break;
    }
    const bool bare = cfg->ess_threshold != cfg->ess_threshold; /* NaN: bare resampler function */
    double logev = 0.0;
    for (int64_t s = 0; s <= T; ++s) {
        /* ---- resample_propagate! (src/container.jl:325,346) */
        double e = refresh_weights(sw);
        if (sw.err) return sw.err;
        sw.ess[s] = e;
        bool doit = bare ? true : (e <= cfg->ess_threshold * (double)N); /* :242-244 */
        sw.resampled[s] = doit ? 1 : 0;
        int32_t *anc_next = anc_hist + (size_t)s * N;
        if (doit && s >= 1) {
            resample_core(sw, s);
            if (sw.err) return sw.err;
        } else {
            /* s == 0: particles carry no state yet, resampling them only re-keys (no-op with
             * position-derived counters); else-branch: update_keys! (:247), weights kept.        */
            for (int64_t i = 0; i < N; ++i) anc_next[i] = (int32_t)i;
            if (doit) std::fill(sw.logw.begin(), sw.logw.end(), 0.0);
        }
        double logZ0 = current_logZ(sw); /* :332,350 */
        if (s == T) break;               /* reweight! finds every particle done (:288) and adds 0 */
        adv(sw, s + 1);                  /* :335,353 */
        double logZ1 = current_logZ(sw); /* :338,356 */
        if (sw.err) return sw.err;
        sw.logz[s] = logZ1;
        logev += logZ1 - logZ0; /* :341,359 */
    }
    *logevidence = logev;
    if (final_logw) memcpy(final_logw, sw.logw.data(), (size_t)N * sizeof(double));
    if (final_w) {
        int rc = orc_softmax(sw.logw.data(), N, mode, final_w);
        if (rc) return rc;
    }
    return 0;
}

/* rand(pc.rng, pc) (src/container.jl:33-36, src/smc.jl:127) on the final set + trajectory
 * extraction through the genealogy. Returns the 0-based slot; traj is T x d.                    */
int orc_pick_trajectory(const aps_config *cfg, uint64_t seed, int mode, const double *final_logw,
                        const double *x_hist, const int32_t *anc_hist, int64_t *slot_out, double *traj) {
    Sweep sw;
    sw.cfg = *cfg;
    sw.N = cfg->n_particles;
    sw.T = cfg->n_steps;
    sw.d = cfg->model.d;
    sw.mode = mode;
    sw.err = 0;
    sw.S = aps_weight_shift((uint64_t)sw.N);
    uint64_t U = draw_u53(seed, 0, (uint64_t)(sw.T + 1), APS_DOM_PICK);
    int64_t slot = categorical_logw(sw, final_logw, sw.N, U);
    if (sw.err) return sw.err;
    *slot_out = slot;
    if (traj) {
        const int64_t N = sw.N, T = sw.T;
        const int D = sw.d;
        int64_t j = anc_hist[(size_t)T * N + slot];
        for (int64_t t = T; t >= 1; --t) {
            for (int k = 0; k < D; ++k) traj[(size_t)(t - 1) * D + k] = x_hist[((size_t)(t - 1) * N + j) * D + k];
            j = anc_hist[(size_t)(t - 1) * N + j];
        }
    }
    return 0;
}

/* trajectory of one final-set slot */
int orc_trajectory(const aps_config *cfg, int64_t slot, const double *x_hist, const int32_t *anc_hist,
                   double *traj) {
    const int64_t N = cfg->n_particles, T = cfg->n_steps;
    const int D = cfg->model.d;
    int64_t j = anc_hist[(size_t)T * N + slot];
    for (int64_t t = T; t >= 1; --t) {
        for (int k = 0; k < D; ++k) traj[(size_t)(t - 1) * D + k] = x_hist[((size_t)(t - 1) * N + j) * D + k];
        j = anc_hist[(size_t)(t - 1) * N + j];
    }
    return 0;
}

/* synthetic data from the model itself (bench + fixtures): x path and y, keyed by data_key */
int orc_simulate_data(const aps_model *model, int64_t T, uint64_t data_key, double *x_out, double *y_out) {
    aps_model_dev md;
    if (aps_model_prepare(model, &md)) return 1;
    const int d = model->d, dy = model->dy;
    double x[APS_MAX_D] = {0}, xp[APS_MAX_D] = {0};
    for (int64_t t = 1; t <= T; ++t) {
        double z[APS_MAX_D + 1], e[APS_MAX_D + 1];
        for (int j = 0; j < 2; ++j) {
            uint64_t w0, w1;
            aps_philox2x64(0, aps_ctr1((uint64_t)t, APS_DOM_DATA, (uint32_t)j), data_key, &w0, &w1);
            aps_normal_pair(w0, w1, &z[2 * j], &z[2 * j + 1]);
            aps_philox2x64(1, aps_ctr1((uint64_t)t, APS_DOM_DATA, (uint32_t)j), data_key, &w0, &w1);
            aps_normal_pair(w0, w1, &e[2 * j], &e[2 * j + 1]);
        }
        for (int k = 0; k < d; ++k) {
            if (t == 1) {
                x[k] = model->mu0[k] + model->sigma0[k] * z[k];
            } else {
                double acc = model->b[k];
                for (int l = 0; l < d; ++l) acc += model->A[k * APS_MAX_D + l] * xp[l];
                x[k] = acc + model->q[k] * z[k];
            }
        }
        for (int m = 0; m < dy; ++m) {
            double y;
            if (model->obs_kind == APS_OBS_LINEAR_GAUSS) {
                double mean = 0.0;
                for (int l = 0; l < d; ++l) mean += model->H[m * APS_MAX_D + l] * x[l];
                y = mean + model->r[m] * e[m];
            } else if (model->obs_kind == APS_OBS_STOCH_VOL) {
                y = aps_exp(0.5 * x[0]) * e[m];
            } else {
                y = 0.0;
            }
            y_out[(size_t)(t - 1) * dy + m] = y;
        }
        for (int k = 0; k < d; ++k) {
            if (x_out) x_out[(size_t)(t - 1) * d + k] = x[k];
            xp[k] = x[k];
        }
    }
    return 0;
}

} /* extern "C" */
