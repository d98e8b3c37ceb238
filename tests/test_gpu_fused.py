"""The fused persistent sweep kernel (csrc/aps_fused.cuh): a single-GPU systematic / stratified sweep
of the d = 1 families as ONE cooperative launch. By default it runs where it is the faster path
(PGAS conditional sweeps, stratified resampling); APS_FUSED=1 forces it everywhere it is eligible,
which is how this module checks it against the oracle on shapes the other parity tests do not reach
(several expand passes per CTA, two-slab history, many sweeps on one handle) and against the
three-kernel path bit for bit."""
import numpy as np
import pytest

import oracle as O
from advancedps_b200 import _abi, _lib, models
from test_gpu_sweep_parity import assert_sweep_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def force_fused(monkeypatch):
    monkeypatch.setenv("APS_FUSED", "1")


def handle(m, N, T, Y, **kw):
    h = _lib.Handle(_abi.make_config(m, N, T, **kw))
    h.set_observations(Y)
    return h


def test_fused_is_the_path_that_runs():
    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, 5, 1)
    for res, want in ((_abi.RESAMPLE_SYSTEMATIC, 1), (_abi.RESAMPLE_STRATIFIED, 1)):
        h = handle(m, 5000, 5, Y, resampler=res)
        h.sweep(1)
        assert h.last_sweep_launches() == want
    h = handle(m, 5000, 5, Y, resampler=_abi.RESAMPLE_MULTINOMIAL)
    h.sweep(1)
    assert h.last_sweep_launches() > 5      # multinomial / residual: the per-step kernels


def test_default_dispatch(monkeypatch):
    """Without APS_FUSED: fused for stratified resampling and for PGAS sweeps that condition on a
    reference (5 launches per step otherwise), the three-kernel graph for the rest."""
    monkeypatch.delenv("APS_FUSED")
    lg, sv = models.linear_gaussian(), models.stochastic_volatility()
    _, Y = O.simulate_data(lg, 6, 1)
    h = handle(lg, 5000, 6, Y)
    h.sweep(1)
    assert h.last_sweep_launches() == 3 * 6 + 2
    h = handle(lg, 5000, 6, Y, resampler=_abi.RESAMPLE_STRATIFIED)
    h.sweep(1)
    assert h.last_sweep_launches() == 1
    h = handle(sv, 5000, 6, Y, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
    h.sweep(1)
    assert h.last_sweep_launches() > 1      # unconditional first sweep: nothing for ancestor sampling to do
    h.pick_trajectory()
    h.sweep(2, ref_on_device=True)
    assert h.last_sweep_launches() == 1


@pytest.mark.parametrize("N,T,res,thr", [
    (3_000_000, 3, _abi.RESAMPLE_SYSTEMATIC, float("nan")),    # ~20k slots per CTA: two expand passes
    (2_500_003, 3, _abi.RESAMPLE_STRATIFIED, 0.7),
    (64, 4, _abi.RESAMPLE_SYSTEMATIC, float("nan")),
    (65, 4, _abi.RESAMPLE_STRATIFIED, float("nan")),
    (9473, 6, _abi.RESAMPLE_SYSTEMATIC, 0.5),                   # exactly 148 x 64 + 1
])
def test_fused_sizes(N, T, res, thr):
    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, T, 0xDA7A0002)
    cfg = _abi.make_config(m, N, T, resampler=res, ess_threshold=thr)
    O.set_threads(O.max_threads())
    try:
        ro = O.sweep(cfg, Y, 31, mode=O.CANON)
    finally:
        O.set_threads(1)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    le = h.sweep(31)
    assert h.last_sweep_launches() == 1
    assert_sweep_equal(cfg, ro, h, le)


def test_fused_degenerate_weights_large():
    """One parent takes (almost) every child at N = 1e6: every CTA writes its own slots, no lists."""
    m = models.linear_gaussian(r=0.00002)
    N, T = 1_000_000, 4
    _, Y = O.simulate_data(m, T, 5)
    cfg = _abi.make_config(m, N, T)
    ro = O.sweep(cfg, Y, 3, mode=O.CANON)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    le = h.sweep(3)
    assert ro.ess[1:].min() < 50
    assert_sweep_equal(cfg, ro, h, le)


def test_fused_two_slab_history_and_many_sweeps():
    m = models.stochastic_volatility()
    for T in (6, 7):
        _, Y = O.simulate_data(m, T, 2)
        cfg = _abi.make_config(m, 40000, T, keep_history=False, ess_threshold=0.5)
        h = _lib.Handle(cfg)
        h.set_observations(Y)
        for seed in range(1, 9):
            le = h.sweep(seed)
            ro = O.sweep(_abi.make_config(m, 40000, T, ess_threshold=0.5), Y, seed, mode=O.CANON)
            assert le == ro.logevidence
            assert np.array_equal(h.weights(), ro.final_w)
            assert np.array_equal(h.ancestors(T + 1), ro.anc_hist[T])
            assert np.array_equal(h.states(T), ro.x_hist[T - 1])


@pytest.mark.parametrize("case", ["lg1-sys", "sv-strat-ess", "lg1-pg", "sv-pgas"])
def test_fused_equals_three_kernel_path(case, monkeypatch):
    mk, res, thr, smp = {
        "lg1-sys": (models.linear_gaussian, _abi.RESAMPLE_SYSTEMATIC, float("nan"), _abi.SAMPLER_SMC),
        "sv-strat-ess": (models.stochastic_volatility, _abi.RESAMPLE_STRATIFIED, 0.5, _abi.SAMPLER_SMC),
        "lg1-pg": (models.linear_gaussian, _abi.RESAMPLE_SYSTEMATIC, 0.5, _abi.SAMPLER_PG),
        "sv-pgas": (models.stochastic_volatility, _abi.RESAMPLE_SYSTEMATIC, 1.0, _abi.SAMPLER_PGAS),
    }[case]
    m = mk()
    N, T = 150_001, 9
    _, Y = O.simulate_data(m, T, 7)
    ref = np.linspace(-0.3, 0.4, T).reshape(T, 1) if smp != _abi.SAMPLER_SMC else None
    hf = handle(m, N, T, Y, resampler=res, ess_threshold=thr, sampler=smp)
    lef = hf.sweep(5, ref_traj=ref)
    assert hf.last_sweep_launches() == 1
    monkeypatch.delenv("APS_FUSED")
    monkeypatch.setenv("APS_NO_FUSED", "1")
    hk = handle(m, N, T, Y, resampler=res, ess_threshold=thr, sampler=smp)
    lek = hk.sweep(5, ref_traj=ref)
    assert hk.last_sweep_launches() > 1
    assert lef == lek
    for t in range(1, T + 1):
        assert np.array_equal(hf.states(t), hk.states(t)), f"states differ at t={t}"
    for t in range(2, T + 2):
        assert np.array_equal(hf.ancestors(t), hk.ancestors(t)), f"ancestors differ at t={t}"
    for a, b in zip(hf.step_stats(), hk.step_stats()):
        assert np.array_equal(a, b)
    assert np.array_equal(hf.weights(), hk.weights()) and np.array_equal(hf.logweights(), hk.logweights())
    assert hf.pick_trajectory()[0] == hk.pick_trajectory()[0]
