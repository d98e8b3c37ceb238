/*
 * aps_model.h -- state-space model descriptor and its per-particle arithmetic.
 *
 * Replaces, for the recognised model families, the Julia closures the reference calls through
 * SSMProblems (src/pgas.jl:60-76: simulate(prior) / simulate(dyn, step, x) / logdensity(obs, ...))
 * and the PGAS transition density (src/pgas.jl:26-32). Families cover the models the reference
 * itself defines: linear-Gaussian (test/linear-gaussian.jl:59-94, test/pgas.jl:12-37,
 * examples/gaussian-ssm/script.jl:37-69) and stochastic volatility
 * (examples/particle-gibbs/script.jl:55-83); plus a constant-log-likelihood observation used to
 * restate the RNG-independent known answers of test/smc.jl:104 and test/container.jl:4-18.
 * Noise parameters are STANDARD DEVIATIONS, as on the AdvancedPS side of those files.
 *
 * Shared by nvcc and gcc so the per-particle results are bit-identical (see aps_math.h).
 */
#ifndef APS_MODEL_H
#define APS_MODEL_H

#include "aps_math.h"

#define APS_MAX_D 4

enum aps_obs_kind {
    APS_OBS_LINEAR_GAUSS = 0, /* y_t ~ N(H x_t, diag(r)^2)                        */
    APS_OBS_STOCH_VOL = 1,    /* y_t ~ N(0, exp(x_t / 2)^2), d = dy = 1            */
    APS_OBS_CONST = 2         /* log g(y_t | x_t) = y_t  (state-independent)       */
};

/* user-facing model: x_1 ~ N(mu0, diag(sigma0)^2); x_t = A x_{t-1} + b + diag(q) eps */
typedef struct aps_model {
    int32_t obs_kind;
    int32_t d;  /* state dimension, 1..APS_MAX_D        */
    int32_t dy; /* observation dimension, 1..APS_MAX_D  */
    int32_t reserved;
    double mu0[APS_MAX_D];
    double sigma0[APS_MAX_D];
    double A[APS_MAX_D * APS_MAX_D]; /* row-major d x d, leading dimension APS_MAX_D */
    double b[APS_MAX_D];
    double q[APS_MAX_D];
    double H[APS_MAX_D * APS_MAX_D]; /* row-major dy x d, leading dimension APS_MAX_D */
    double r[APS_MAX_D];
} aps_model;

/* model + constants derived once on the host (same code on both sides) */
typedef struct aps_model_dev {
    aps_model m;
    double inv_r[APS_MAX_D];
    double inv_q[APS_MAX_D];
    double obs_const;   /* -sum log r - dy/2 log(2 pi)  */
    double trans_const; /* -sum log q - d/2 log(2 pi)   */
} aps_model_dev;

#define APS_HALF_LOG_2PI 0.9189385332046728

static inline int aps_model_prepare(const aps_model *m, aps_model_dev *out) {
    if (m->d < 1 || m->d > APS_MAX_D || m->dy < 1 || m->dy > APS_MAX_D) return 1;
    if (m->obs_kind < 0 || m->obs_kind > 2) return 1;
    if (m->obs_kind == APS_OBS_STOCH_VOL && (m->d != 1 || m->dy != 1)) return 1;
    out->m = *m;
    double oc = 0.0, tc = 0.0;
    for (int k = 0; k < APS_MAX_D; ++k) {
        out->inv_r[k] = 0.0;
        out->inv_q[k] = 0.0;
    }
    for (int k = 0; k < m->dy; ++k) {
        if (m->obs_kind == APS_OBS_LINEAR_GAUSS) {
            if (!(m->r[k] > 0.0)) return 1;
            out->inv_r[k] = 1.0 / m->r[k];
            oc = oc - aps_log(m->r[k]);
        }
        oc = oc - APS_HALF_LOG_2PI;
    }
    for (int k = 0; k < m->d; ++k) {
        if (!(m->q[k] > 0.0) || !(m->sigma0[k] >= 0.0)) return 1;
        out->inv_q[k] = 1.0 / m->q[k];
        tc = tc - aps_log(m->q[k]) - APS_HALF_LOG_2PI;
    }
    out->obs_const = oc;
    out->trans_const = tc;
    return 0;
}

#ifdef __cplusplus /* per-particle arithmetic: C++ templates, shared by nvcc and g++ */

/* number of Philox blocks a state draw consumes */
APS_HD int aps_blocks_for_dim(int d) { return (d + 1) >> 1; }

/* State draws are generated per PAIR of adjacent slots p = i >> 1: the pair consumes Philox
 * blocks j = 0..D-1 at counter (p, ctr1(step, DOM_STATE, j)); their 2D Box-Muller normals, in
 * block order, go to slot 2p (first D) and slot 2p+1 (last D). No normal is wasted for odd D.  */
template <int D>
APS_HD void aps_pair_normals(uint64_t key, uint64_t pair, uint64_t step, double *z /* 2 D */) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < D; ++j) {
        uint64_t w0, w1;
        aps_philox2x64(pair, aps_ctr1(step, APS_DOM_STATE, (uint32_t)j), key, &w0, &w1);
        aps_normal_pair(w0, w1, &z[2 * j], &z[2 * j + 1]);
    }
}

/* the same draws in two halves, so that a kernel can put its (dependent, long-latency) parent-state
 * gather between the integer-only Philox rounds and the floating-point transform */
template <int D>
APS_HD void aps_pair_words(uint64_t key, uint64_t pair, uint64_t step, uint64_t *w /* 2 D */) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < D; ++j)
        aps_philox2x64(pair, aps_ctr1(step, APS_DOM_STATE, (uint32_t)j), key, &w[2 * j], &w[2 * j + 1]);
}
template <int D>
APS_HD void aps_words_to_normals(const uint64_t *w, double *z /* 2 D */) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < D; ++j) aps_normal_pair(w[2 * j], w[2 * j + 1], &z[2 * j], &z[2 * j + 1]);
}

/* x_1 = mu0 + sigma0 z   (src/pgas.jl:60-62: simulate(rng, prior)) */
template <int D>
APS_HD void aps_prior_draw(const aps_model_dev *md, const double *z, double *x) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < D; ++k) x[k] = aps_fma(md->m.sigma0[k], z[k], md->m.mu0[k]);
}

/* mean_k = b_k + sum_l A_kl xp_l, accumulated left to right with fma */
template <int D>
APS_HD void aps_trans_mean(const aps_model_dev *md, const double *xp, double *mean) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < D; ++k) {
        double acc = md->m.b[k];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = 0; l < D; ++l) acc = aps_fma(md->m.A[k * APS_MAX_D + l], xp[l], acc);
        mean[k] = acc;
    }
}

/* x_t = mean + q z   (src/pgas.jl:63-67: simulate(rng, dyn, step, x_prev)) */
template <int D>
APS_HD void aps_trans_draw(const aps_model_dev *md, const double *xp, const double *z, double *x) {
    double mean[D];
    aps_trans_mean<D>(md, xp, mean);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < D; ++k) x[k] = aps_fma(md->m.q[k], z[k], mean[k]);
}

/* log f(xn | xp)   (src/pgas.jl:26-32: logdensity(dyn, iter, x_prev, x)) */
template <int D>
APS_HD double aps_trans_logpdf(const aps_model_dev *md, const double *xp, const double *xn) {
    double mean[D];
    aps_trans_mean<D>(md, xp, mean);
    double acc = md->trans_const;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < D; ++k) {
        double zz = (xn[k] - mean[k]) * md->inv_q[k];
        acc = aps_fma(-0.5 * zz, zz, acc);
    }
    return acc;
}

/* log g(y | x)   (src/pgas.jl:74-76: logdensity(obs, step, x, y)); y has DY entries */
template <int D, int DY, int OBS>
APS_HD double aps_obs_logpdf(const aps_model_dev *md, const double *x, const double *y) {
    if (OBS == APS_OBS_CONST) return y[0];
    if (OBS == APS_OBS_STOCH_VOL) {
        /* N(0, sigma = exp(x/2)): -y^2 exp(-x)/2 - x/2 - log(2 pi)/2 */
        double e = aps_exp(-x[0]);
        double t = y[0] * y[0];
        return aps_fma(-0.5 * t, e, aps_fma(-0.5, x[0], -APS_HALF_LOG_2PI));
    }
    double acc = md->obs_const;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int m = 0; m < DY; ++m) {
        double mean = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = 0; l < D; ++l) mean = aps_fma(md->m.H[m * APS_MAX_D + l], x[l], mean);
        double zz = (y[m] - mean) * md->inv_r[m];
        acc = aps_fma(-0.5 * zz, zz, acc);
    }
    return acc;
}

#endif /* __cplusplus */

#endif /* APS_MODEL_H */
