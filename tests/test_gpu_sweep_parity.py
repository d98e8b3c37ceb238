"""GPU parity: the CUDA sweep (through the C ABI) against the oracle in CANON mode.

Bar: ancestor indices, states, log-weights and per-step logZ / ESS bit-exact; log-evidence
bit-equal (the 1e-6 relative tolerance of BASELINE.json is therefore met with margin).
"""
import numpy as np
import pytest

import advancedps_b200 as aps  # noqa: F401
import oracle as O
from advancedps_b200 import _abi, _lib, models

pytestmark = pytest.mark.gpu


def run_both(model, N, T, seed, data_key, sampler=_abi.SAMPLER_SMC, resampler=_abi.RESAMPLE_SYSTEMATIC,
             ess_threshold=float("nan"), ref=None, Y=None):
    cfg = _abi.make_config(model, N, T, sampler=sampler, resampler=resampler, ess_threshold=ess_threshold)
    if Y is None:
        _, Y = O.simulate_data(model, T, data_key)
    ro = O.sweep(cfg, Y, seed, ref_traj=ref, mode=O.CANON)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    le = h.sweep(seed, ref_traj=ref)
    return cfg, Y, ro, h, le


def assert_sweep_equal(cfg, ro, h, le):
    T = cfg.n_steps
    logz, ess, res = h.step_stats()
    assert np.array_equal(res, ro.resampled)
    for t in range(1, T + 1):
        xg = h.states(t)
        assert np.array_equal(xg, ro.x_hist[t - 1]), f"states differ at t={t}"
    for t in range(2, T + 2):
        ag = h.ancestors(t)
        bad = np.nonzero(ag != ro.anc_hist[t - 1])[0]
        assert bad.size == 0, f"{bad.size} ancestors differ at t={t}, first {bad[:5]}"
    assert np.array_equal(logz, ro.logz)
    assert np.array_equal(ess, ro.ess)
    assert le == ro.logevidence
    assert np.array_equal(h.logweights(), ro.final_logw)
    assert np.array_equal(h.weights(), ro.final_w)


@pytest.mark.parametrize("N,T", [(1, 3), (3, 5), (1000, 50), (2048, 7), (2049, 7), (100003, 12)])
def test_lg1_smc_systematic_bare(N, T):
    """configs[0] (C1: LG d=1, T=50, N=1000, SMC + resample_systematic) and ragged sizes."""
    cfg, Y, ro, h, le = run_both(models.linear_gaussian(), N, T, 1234, 0xDA7A0001)
    assert_sweep_equal(cfg, ro, h, le)


def test_lg1_c1_vs_kalman():
    m = models.linear_gaussian()
    cfg, Y, ro, h, le = run_both(m, 1000, 50, 1234, 0xDA7A0001)
    ll, _, _ = models.kalman_loglik(m, Y)
    assert abs(le - ll) < 1.5  # Monte-Carlo error of a 1000-particle filter over 50 steps


@pytest.mark.parametrize("thr", [0.5, 1.0, 0.0])
def test_lg1_ess_threshold(thr):
    cfg, Y, ro, h, le = run_both(models.linear_gaussian(), 5000, 30, 99, 0xDA7A0001, ess_threshold=thr)
    assert_sweep_equal(cfg, ro, h, le)
    if thr == 0.0:
        assert ro.resampled.sum() == 0


def test_lg1_stratified():
    cfg, Y, ro, h, le = run_both(models.linear_gaussian(), 7001, 20, 5, 0xDA7A0001,
                                 resampler=_abi.RESAMPLE_STRATIFIED)
    assert_sweep_equal(cfg, ro, h, le)


@pytest.mark.parametrize("resampler", [_abi.RESAMPLE_MULTINOMIAL, _abi.RESAMPLE_RESIDUAL])
@pytest.mark.parametrize("thr", [float("nan"), 0.5])
def test_lg1_multinomial_residual(resampler, thr):
    cfg, Y, ro, h, le = run_both(models.linear_gaussian(), 6007, 14, 21, 0xDA7A0001, resampler=resampler,
                                 ess_threshold=thr)
    assert_sweep_equal(cfg, ro, h, le)


@pytest.mark.parametrize("resampler", [_abi.RESAMPLE_MULTINOMIAL, _abi.RESAMPLE_RESIDUAL, _abi.RESAMPLE_STRATIFIED])
def test_pg_conditional_other_resamplers(resampler):
    m = models.linear_gaussian()
    N, T = 3000, 8
    cfg = _abi.make_config(m, N, T, sampler=_abi.SAMPLER_PG, resampler=resampler, ess_threshold=0.5)
    _, Y = O.simulate_data(m, T, 3)
    ref = np.linspace(0.2, 0.6, T).reshape(T, 1)
    ro = O.sweep(cfg, Y, 5, ref_traj=ref, mode=O.CANON)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    le = h.sweep(5, ref_traj=ref)
    assert_sweep_equal(cfg, ro, h, le)


def test_lg4_smc():
    cfg, Y, ro, h, le = run_both(models.lg4(), 20000, 15, 7, 0xDA7A0003)
    assert_sweep_equal(cfg, ro, h, le)
    ll, _, _ = models.kalman_loglik(models.lg4(), Y)
    assert abs(le - ll) < 5.0


def test_sv_smc():
    cfg, Y, ro, h, le = run_both(models.stochastic_volatility(), 30000, 25, 11, 0xDA7A0004, ess_threshold=0.5)
    assert_sweep_equal(cfg, ro, h, le)


def test_same_seed_same_result_and_graph_replay():
    """test/pgas.jl:99-127: same rng state => same result; also exercises CUDA-graph replay."""
    m = models.linear_gaussian()
    cfg = _abi.make_config(m, 4096, 10)
    _, Y = O.simulate_data(m, 10, 1)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    a = h.sweep(42)
    xa = h.states(10).copy()
    b = h.sweep(43)
    c = h.sweep(42)
    assert a == c and a != b
    assert np.array_equal(xa, h.states(10))


@pytest.mark.parametrize("sampler", [_abi.SAMPLER_PG, _abi.SAMPLER_PGAS])
def test_conditional_sweeps(sampler):
    """PG / PGAS: unconditional sweep, pick, then conditional sweeps (src/smc.jl:101-129)."""
    m = models.stochastic_volatility() if sampler == _abi.SAMPLER_PGAS else models.linear_gaussian()
    N, T = 3000, 12
    thr = 1.0 if sampler == _abi.SAMPLER_PGAS else 0.5  # PGAS(n) default, src/smc.jl:99
    cfg = _abi.make_config(m, N, T, sampler=sampler, ess_threshold=thr)
    _, Y = O.simulate_data(m, T, 0xDA7A0004)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    ref = None
    for seed in [1, 2, 3]:
        ro = O.sweep(cfg, Y, seed, ref_traj=ref, mode=O.CANON)
        le = h.sweep(seed, ref_traj=ref)
        assert_sweep_equal(cfg, ro, h, le)
        slot_o, traj_o = O.pick_trajectory(cfg, seed, ro, mode=O.CANON)
        slot_g, traj_g = h.pick_trajectory()
        assert slot_o == slot_g
        assert np.array_equal(traj_o, traj_g)
        assert np.array_equal(h.trajectory(slot_g), traj_g)
        if ref is not None:
            # the reference stays in the last slot (test/container.jl:91)
            for t in range(1, T + 1):
                assert np.array_equal(h.states(t)[N - 1], ref[t - 1])
        ref = traj_g
    # conditioning on the device-resident picked trajectory == passing it from the host
    ro = O.sweep(cfg, Y, 77, ref_traj=ref, mode=O.CANON)
    assert h.sweep(77, ref_on_device=True) == ro.logevidence


def test_weights_view_is_the_weights():
    cfg, Y, ro, h, le = run_both(models.linear_gaussian(), 5000, 9, 3, 2, ess_threshold=0.0)
    v = h.weights_view()
    assert not v.flags.writeable and np.array_equal(v, ro.final_w) and np.array_equal(h.weights(), v)


def test_final_states():
    m = models.linear_gaussian()
    cfg, Y, ro, h, le = run_both(m, 5000, 9, 3, 2, ess_threshold=0.5)
    T = 9
    assert np.array_equal(h.final_states(), ro.x_hist[T - 1][ro.anc_hist[T]])


def test_constant_loglik_evidence_known_answer():
    """test/smc.jl:104: a likelihood that ignores the state gives logevidence = -2 log 2."""
    m = models.constant_loglik()
    cfg = _abi.make_config(m, 100, 2)
    Y = np.full((2, 1), np.log(0.5))
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    le = h.sweep(100)
    assert abs(le - (-2 * np.log(2))) < 1e-14


def test_not_normalisable_weights_raise():
    """all log-weights -Inf -> error, as the reference's resamplers throw (src/resampling.jl:120,169)."""
    m = models.constant_loglik()
    cfg = _abi.make_config(m, 100, 2)
    h = _lib.Handle(cfg)
    h.set_observations(np.full((2, 1), -np.inf))
    with pytest.raises(_lib.ApsError) as e:
        h.sweep(1)
    assert e.value.code == _abi.ERR_WEIGHTS


@pytest.mark.parametrize("thr", [float("nan"), 0.5])
def test_smoothing_mean_matches_host_backtrace(thr):
    """N2: weighted mean trajectory from the device genealogy == the same sum formed on
    the host from the oracle's state / ancestor history (fp64 sums in a different order: 1e-12)."""
    m = models.lg4()
    N, T = 6000, 9
    cfg, Y, ro, h, le = run_both(m, N, T, 4, 0xDA7A0003, ess_threshold=thr)
    b = ro.anc_hist[T].astype(np.int64)
    want = np.zeros((T, m.d))
    for t in range(T, 0, -1):
        want[t - 1] = (ro.final_w[:, None] * ro.x_hist[t - 1][b]).sum(axis=0)
        b = ro.anc_hist[t - 1][b].astype(np.int64)
    got = h.smoothing_mean()
    assert np.allclose(got, want, rtol=1e-12, atol=1e-13)
    # and it is what averaging the materialised trajectories gives
    some = [0, 1, N // 2, N - 1]
    for i in some:
        assert np.array_equal(h.trajectory(i), O.trajectory(cfg, i, ro))
