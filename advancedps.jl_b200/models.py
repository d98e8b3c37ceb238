"""State-space model families the device path recognises, and the BASELINE.json configurations.

Parameters are standard deviations, as on the AdvancedPS side of the reference's models
(test/linear-gaussian.jl:59-94, test/pgas.jl:12-37, examples/particle-gibbs/script.jl:55-83).
"""
import numpy as np

from . import _abi


def linear_gaussian(a=0.5, b=0.2, q=0.1, h=1.0, r=0.1, x0=0.0, sigma0=1.0):
    """1-D linear-Gaussian SSM; defaults are test/linear-gaussian.jl:32-42 (config C1/C2/C5)."""
    return _abi.make_model(_abi.OBS_LINEAR_GAUSS, 1, 1, [x0], [sigma0], [[a]], [b], [q], [[h]], [r])


def linear_gaussian_nd(A, b, q, H, r, mu0, sigma0):
    """d-dimensional linear-Gaussian SSM with diagonal noise (config C3 uses d = dy = 4)."""
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    H = np.atleast_2d(np.asarray(H, dtype=np.float64))
    d, dy = A.shape[0], H.shape[0]
    return _abi.make_model(_abi.OBS_LINEAR_GAUSS, d, dy, mu0, sigma0, A, b, q, H, r)


def lg4():
    """Config C3: d = dy = 4, A = .5 I + .1 (11' - I), b = .2, q = r = .1, H = I, prior N(0, I)."""
    d = 4
    A = 0.5 * np.eye(d) + 0.1 * (np.ones((d, d)) - np.eye(d))
    return linear_gaussian_nd(A, 0.2 * np.ones(d), 0.1 * np.ones(d), np.eye(d), 0.1 * np.ones(d),
                              np.zeros(d), np.ones(d))


def stochastic_volatility(a=0.9, q=0.5):
    """examples/particle-gibbs/script.jl:55-83: x1 ~ N(0, q), x' ~ N(a x, q), y ~ N(0, exp(x/2))."""
    return _abi.make_model(_abi.OBS_STOCH_VOL, 1, 1, [0.0], [q], [[a]], [0.0], [q])


def constant_loglik(d=1):
    """Observation log-density equal to y_t regardless of the state (known-answer tests:
    test/smc.jl:104 evidence -2 log 2; test/container.jl:4-18 LogPModel)."""
    return _abi.make_model(_abi.OBS_CONST, d, 1, np.zeros(d), np.ones(d), np.zeros((d, d)),
                           np.zeros(d), np.ones(d))


def kalman_loglik(model, Y):
    """Exact log p(y_1:T) of a linear-Gaussian model (numpy; ground truth for tests/bench)."""
    d, dy = model.d, model.dy
    A = np.array(model.A[:]).reshape(_abi.APS_MAX_D, _abi.APS_MAX_D)[:d, :d]
    H = np.array(model.H[:]).reshape(_abi.APS_MAX_D, _abi.APS_MAX_D)[:dy, :d]
    b = np.array(model.b[:d])
    Qm = np.diag(np.array(model.q[:d]) ** 2)
    Rm = np.diag(np.array(model.r[:dy]) ** 2)
    m = np.array(model.mu0[:d])
    P = np.diag(np.array(model.sigma0[:d]) ** 2)
    Y = np.asarray(Y, dtype=np.float64).reshape(-1, dy)
    ll = 0.0
    means, covs = [], []
    for t in range(Y.shape[0]):
        if t > 0:
            m = A @ m + b
            P = A @ P @ A.T + Qm
        S = H @ P @ H.T + Rm
        v = Y[t] - H @ m
        ll += -0.5 * (v @ np.linalg.solve(S, v) + np.linalg.slogdet(S)[1] + dy * np.log(2 * np.pi))
        K = P @ H.T @ np.linalg.inv(S)
        m = m + K @ v
        P = P - K @ H @ P
        means.append(m.copy())
        covs.append(P.copy())
    return ll, np.array(means), np.array(covs)
