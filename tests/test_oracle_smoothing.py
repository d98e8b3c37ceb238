"""PG / PGAS chains against the exact Rauch-Tung-Striebel smoother of the linear-Gaussian model --
the meaningful version of the reference's KS test (test/linear-gaussian.jl:99-111 tests 3 numbers and
mixes variances with standard deviations). CPU only, on the oracle (the GPU path is
bit-compared with it); fixed seeds, so the check is deterministic."""
import numpy as np
import pytest

import oracle as O
from advancedps_b200 import _abi, models


def rts_smoother(m, Y):
    """1-d model: smoothing means / variances from the Kalman filter moments."""
    _, fm, fc = models.kalman_loglik(m, Y)
    a, b, q = m.A[0], m.b[0], m.q[0]
    sm, sv = fm[:, 0].copy(), fc[:, 0, 0].copy()
    for t in range(len(sm) - 2, -1, -1):
        pp = a * fc[t, 0, 0] * a + q * q
        g = fc[t, 0, 0] * a / pp
        sm[t] = fm[t, 0] + g * (sm[t + 1] - (a * fm[t, 0] + b))
        sv[t] = fc[t, 0, 0] + g * (sv[t + 1] - pp) * g
    return sm, sv


@pytest.mark.parametrize("sampler,thr", [(_abi.SAMPLER_PG, 0.5), (_abi.SAMPLER_PGAS, 1.0)])
def test_particle_gibbs_chain_matches_rts_smoother(sampler, thr):
    m = models.linear_gaussian()
    T, N, iters, burn = 8, 16, 3000, 200
    _, Y = O.simulate_data(m, T, 42)
    sm, sv = rts_smoother(m, Y)
    cfg = _abi.make_config(m, N, T, sampler=sampler, ess_threshold=thr)
    ref, acc, acc2 = None, np.zeros(T), np.zeros(T)
    for k in range(iters):                       # AbstractMCMC's loop around step (src/smc.jl:101-129)
        r = O.sweep(cfg, Y, 1000 + k, ref_traj=ref)
        _, ref = O.pick_trajectory(cfg, 1000 + k, r)
        if k >= burn:
            acc += ref[:, 0]
            acc2 += ref[:, 0] ** 2
    n = iters - burn
    mean, var = acc / n, acc2 / n - (acc / n) ** 2
    assert np.max(np.abs(mean - sm) / np.sqrt(sv)) < 0.25     # observed 0.07 (PG) / 0.12 (PGAS)
    assert np.all((var / sv > 0.75) & (var / sv < 1.35))      # observed 0.92 .. 1.11
