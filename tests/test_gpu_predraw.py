"""Pre-drawn state normals (k_draw_normals on a parallel branch of the sweep's graph, csrc/aps_api.cu
enqueue_sweep): by default only where it is the faster path (one wave of tiles, N >= ~5e5 per GPU);
APS_PREDRAW=1 forces it, which is how this module checks it against the oracle on small and ragged
shapes, for every sampler / resampler that runs the three-kernel path, one and two steps ahead, and
against the path that draws inside the propagate kernel bit for bit."""
import numpy as np
import pytest

import oracle as O
from advancedps_b200 import _abi, _lib, models
from test_gpu_sweep_parity import assert_sweep_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def force_predraw(monkeypatch):
    monkeypatch.setenv("APS_PREDRAW", "1")
    monkeypatch.setenv("APS_NO_FUSED", "1")   # (the fused persistent kernel has no separate draw kernel)


def both(model, N, T, seed, **kw):
    cfg = _abi.make_config(model, N, T, **kw)
    _, Y = O.simulate_data(model, T, 0xDA7A0001)
    ro = O.sweep(cfg, Y, seed, mode=O.CANON)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    le = h.sweep(seed)
    return cfg, Y, ro, h, le


@pytest.mark.parametrize("ahead", ["1", "2"])
@pytest.mark.parametrize("N,T", [(1, 3), (3, 5), (2049, 7), (100003, 12), (70000, 1), (70000, 2)])
def test_predrawn_sweep_equals_oracle(monkeypatch, ahead, N, T):
    monkeypatch.setenv("APS_DRAW_AHEAD", ahead)
    cfg, Y, ro, h, le = both(models.linear_gaussian(), N, T, 1234)
    assert h.last_sweep_launches() == 4 * T + 2          # T draw kernels on top of the 3 T + 2 of the plain graph
    assert_sweep_equal(cfg, ro, h, le)
    # replay of the same graph with another seed
    ro2 = O.sweep(cfg, Y, 77, mode=O.CANON)
    assert h.sweep(77) == ro2.logevidence


@pytest.mark.parametrize("kw", [
    dict(ess_threshold=0.5),
    dict(resampler=_abi.RESAMPLE_STRATIFIED),
    dict(resampler=_abi.RESAMPLE_MULTINOMIAL),
    dict(resampler=_abi.RESAMPLE_RESIDUAL, ess_threshold=0.5),
])
def test_predrawn_resamplers(kw):
    cfg, Y, ro, h, le = both(models.linear_gaussian(), 6007, 14, 21, **kw)
    assert_sweep_equal(cfg, ro, h, le)


@pytest.mark.parametrize("d", [2, 3])
def test_predrawn_d2_d3(d):
    A = 0.5 * np.eye(d) + 0.1 * (np.ones((d, d)) - np.eye(d))
    m = models.linear_gaussian_nd(A, 0.2 * np.ones(d), 0.1 * np.ones(d), np.eye(d), 0.1 * np.ones(d), np.zeros(d), np.ones(d))
    cfg, Y, ro, h, le = both(m, 4099, 9, 5)
    assert_sweep_equal(cfg, ro, h, le)


def test_predrawn_d4_per_slot_kernel():
    """d = 4 runs k_propagate1 (one thread per slot); its PRE variant reads the slot's four normals from the same
    pair-layout buffer. Opt-in only (APS_PREDRAW=1): at the configs[2] shape it is slower than drawing in the
    kernel, but the path must stay correct."""
    cfg, Y, ro, h, le = both(models.lg4(), 20001, 9, 7)
    assert h.last_sweep_launches() == 4 * 9 + 2
    assert_sweep_equal(cfg, ro, h, le)


def test_predrawn_sv_pgas_conditional():
    """Conditional PGAS sweeps (reference trajectory in the last slot) on the three-kernel path."""
    m = models.stochastic_volatility()
    T, N = 15, 8192
    cfg = _abi.make_config(m, N, T, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
    _, Y = O.simulate_data(m, T, 3)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    ro = O.sweep(cfg, Y, 11, mode=O.CANON)
    assert h.sweep(11) == ro.logevidence
    slot, traj = h.pick_trajectory()
    ro2 = O.sweep(cfg, Y, 12, ref_traj=traj, mode=O.CANON)
    le2 = h.sweep(12, ref_on_device=True)
    assert_sweep_equal(cfg, ro2, h, le2)


def test_predrawn_equals_in_kernel_draws(monkeypatch):
    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, 20, 9)
    cfg = _abi.make_config(m, 300_001, 20)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    le = h.sweep(5)
    x, a, w = h.states(20).copy(), h.ancestors(21).copy(), h.weights().copy()
    monkeypatch.delenv("APS_PREDRAW")
    monkeypatch.setenv("APS_NO_PREDRAW", "1")
    h2 = _lib.Handle(cfg)
    h2.set_observations(Y)
    assert h2.sweep(5) == le
    assert h2.last_sweep_launches() == 3 * 20 + 2
    assert np.array_equal(h2.states(20), x) and np.array_equal(h2.ancestors(21), a) and np.array_equal(h2.weights(), w)


@pytest.mark.parametrize("world,N,T,res,thr", [
    (2, 4096, 6, _abi.RESAMPLE_SYSTEMATIC, float("nan")),
    (2, 8192 * 3, 7, _abi.RESAMPLE_SYSTEMATIC, 0.5),
    (4, 8192 * 3, 7, _abi.RESAMPLE_SYSTEMATIC, 0.5),
    (2, 6400, 5, _abi.RESAMPLE_STRATIFIED, float("nan")),
    (2, 6400, 6, _abi.RESAMPLE_RESIDUAL, float("nan")),
    (8, 8192 * 2, 4, _abi.RESAMPLE_SYSTEMATIC, float("nan")),
])
def test_predrawn_sharded_equals_oracle(monkeypatch, world, N, T, res, thr):
    """The sharded sweep with the draws made ahead (ranks emulated on one GPU): every rank draws the normals of
    its own slot pairs; same result as one GPU and as the oracle. More than two EMULATED ranks launch their
    kernels directly (APS_NO_GRAPH): graphs with parallel branches from four host threads on one GPU do not
    run concurrently, and ranks that spin on each other then never meet (an artefact of the emulation, see
    aps_api.cu sweep_impl; one process per GPU replays the graph: tests/test_gpu_multiprocess.py)."""
    from test_gpu_sharded import assert_sharded_equal, run_sharded

    if world > 2:
        monkeypatch.setenv("APS_NO_GRAPH", "1")

    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, T, 0xDA7A0005)
    hs, out = run_sharded(m, N, T, Y, [11, 12], world, res, thr)
    assert hs[0].last_sweep_launches() > 4 * T
    ro = O.sweep(_abi.make_config(m, N, T, resampler=res, ess_threshold=thr), Y, 12, mode=O.CANON)
    assert_sharded_equal(hs, out[1], ro, N, T)
