"""Per-kernel-class device time (CUDA events around every launch, aps_sweep_profiled) for the
BASELINE config shapes: python scripts/profile_config.py [c2|c3|c4|c5m|c5r|c5s ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from advancedps_b200 import _abi, _lib, models  # noqa: E402

nan = float("nan")
CASES = {
    "c2": (models.linear_gaussian, 10**6, 100, 0, 3, nan),
    "c3": (models.lg4, 4 * 10**6, 200, 1, 3, 0.5),
    "c4": (models.stochastic_volatility, 2 * 10**6, 500, 2, 3, 1.0),
    "c5s": (models.linear_gaussian, 10**6, 100, 0, 2, nan),
    "c5r": (models.linear_gaussian, 10**6, 100, 0, 1, nan),
    "c5m": (models.linear_gaussian, 10**6, 100, 0, 0, nan),
}


def simulate(m, T, rng):
    """observations simulated from the model itself with numpy (no oracle on this path)"""
    d, dy = m.d, m.dy
    A = np.array(m.A).reshape(4, 4)[:d, :d]
    H = np.array(m.H).reshape(4, 4)[:dy, :d]
    x = np.array(m.mu0)[:d] + np.array(m.sigma0)[:d] * rng.normal(size=d)
    Y = np.zeros((T, dy))
    for t in range(T):
        if t:
            x = A @ x + np.array(m.b)[:d] + np.array(m.q)[:d] * rng.normal(size=d)
        if m.obs_kind == 0:
            Y[t] = H @ x + np.array(m.r)[:dy] * rng.normal(size=dy)
        else:
            Y[t] = np.exp(0.5 * x[0]) * rng.normal()
    return Y


rng = np.random.default_rng(0)
for name in sys.argv[1:] or ["c2", "c3", "c4"]:
    fac, N, T, smp, res, thr = CASES[name]
    T = int(os.environ.get("APS_PROF_T", T))
    m = fac()
    h = _lib.Handle(_abi.make_config(m, N, T, sampler=smp, resampler=res, ess_threshold=thr))
    h.set_observations(simulate(m, T, rng))
    h.sweep(1)
    ms_graph = h.last_sweep_ms()
    h.sweep_profiled(2)
    le, ms, n = h.sweep_profiled(3)
    _, ess, rs = h.step_stats()
    print(json.dumps({"config": name, "N": N, "T": T, "graph_ms": ms_graph, "class_ms": ms, "class_launches": n,
                      "avg_us": [1e3 * a / b if b else 0 for a, b in zip(ms, n)],
                      "ess_frac_median": float(np.median(ess[1:]) / N), "resampled": int(rs.sum())}), flush=True)
    h.close()
