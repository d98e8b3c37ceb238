"""Independent anchors for the per-particle densities of include/aps_model.h (the GPU and the oracle
compile the same header, so their equality proves nothing about the arithmetic itself). With an ESS threshold
of 0 nothing is ever resampled, so the final log-weight of particle i is sum_t log g(y_t | x_t^i) over its own
unbroken path; scipy's normal log-density of the same states and observations must agree.
Reference: src/pgas.jl:74-76 (logdensity(obs, step, x, y)), examples/particle-gibbs/script.jl:55-83 (SV),
test/linear-gaussian.jl:59-94 (LG)."""
import numpy as np
from scipy import stats

import advancedps_b200 as aps  # noqa: F401
import oracle as O
from advancedps_b200 import _abi, models


def _paths(model, N, T, Y, seed):
    cfg = _abi.make_config(model, N, T, ess_threshold=0.0)
    ro = O.sweep(cfg, Y, seed, mode=O.CANON)
    assert ro.resampled.sum() == 0
    return ro


def test_sv_observation_density_matches_scipy():
    m = models.stochastic_volatility()
    T, N = 6, 500
    _, Y = O.simulate_data(m, T, 17)
    ro = _paths(m, N, T, Y, 3)
    want = np.zeros(N)
    for t in range(T):
        x = ro.x_hist[t].reshape(N)
        want += stats.norm.logpdf(Y[t, 0], loc=0.0, scale=np.exp(0.5 * x))
    np.testing.assert_allclose(ro.final_logw, want, rtol=1e-12, atol=1e-12)


def test_lg1_observation_density_matches_scipy():
    m = models.linear_gaussian(h=1.3, r=0.25)
    T, N = 5, 400
    _, Y = O.simulate_data(m, T, 5)
    ro = _paths(m, N, T, Y, 9)
    want = np.zeros(N)
    for t in range(T):
        want += stats.norm.logpdf(Y[t, 0], loc=1.3 * ro.x_hist[t].reshape(N), scale=0.25)
    np.testing.assert_allclose(ro.final_logw, want, rtol=1e-12, atol=1e-12)


def test_lg4_observation_density_matches_scipy():
    """d = dy = 4 with a dense H and unequal observation noise (diagonal covariance)."""
    d = 4
    rng = np.random.default_rng(0)
    A = 0.5 * np.eye(d) + 0.1 * (np.ones((d, d)) - np.eye(d))
    H = rng.normal(size=(d, d))
    r = np.array([0.1, 0.2, 0.3, 0.4])
    m = models.linear_gaussian_nd(A, 0.2 * np.ones(d), 0.1 * np.ones(d), H, r, np.zeros(d), np.ones(d))
    T, N = 4, 300
    _, Y = O.simulate_data(m, T, 11)
    ro = _paths(m, N, T, Y, 2)
    want = np.zeros(N)
    for t in range(T):
        x = ro.x_hist[t].reshape(N, d)
        want += stats.norm.logpdf(Y[t][None, :], loc=x @ H.T, scale=r[None, :]).sum(axis=1)
    np.testing.assert_allclose(ro.final_logw, want, rtol=1e-11, atol=1e-11)


def test_transition_moments_match_the_model():
    """x_t | x_{t-1} ~ N(a x_{t-1} + b, q^2): with nothing resampled, slot i's path is one AR(1) draw; the
    standardised innovations must be standard normal (mean, variance, KS)."""
    a, b, q = 0.7, -0.3, 0.5
    m = models.linear_gaussian(a=a, b=b, q=q)
    T, N = 3, 200_000
    _, Y = O.simulate_data(m, T, 1)
    ro = _paths(m, N, T, Y, 4)
    x1, x2 = ro.x_hist[0].reshape(N), ro.x_hist[1].reshape(N)
    e = (x2 - (a * x1 + b)) / q
    assert abs(e.mean()) < 4 / np.sqrt(N) and abs(e.var() - 1.0) < 6 * np.sqrt(2.0 / N)
    assert stats.kstest(e, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(e, x1)[0, 1]) < 5 / np.sqrt(N)
