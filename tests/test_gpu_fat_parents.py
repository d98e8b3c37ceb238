"""Degenerate weights: parents that own a large share of all children ("fat" parents) are not
expanded by the block that owns them; their child range is recorded in a per-step list and
resolved by the consumer (the next propagate kernel / k_fill_fat). APS_FAT_MIN lowers the
threshold so that small test problems exercise the path. Results must stay bit-equal to the oracle."""
import numpy as np
import pytest

import oracle as O
from advancedps_b200 import _abi, _lib, models
from test_gpu_sharded import assert_sharded_equal, collective, make_ranks
from test_gpu_sweep_parity import assert_sweep_equal

pytestmark = pytest.mark.gpu


@pytest.fixture
def low_fat_threshold(monkeypatch):
    monkeypatch.setenv("APS_FAT_MIN", "100")


@pytest.fixture(params=["fused", "three-kernel"])
def sweep_path(request, monkeypatch):
    """The fused persistent kernel (forced with APS_FUSED=1) resolves ancestors per child slot ("pull")
    and needs no fat-parent lists; APS_NO_FUSED=1 selects the three-kernel path, where the lists exist.
    Both must equal the oracle under degenerate weights."""
    monkeypatch.setenv("APS_NO_FUSED" if request.param == "three-kernel" else "APS_FUSED", "1")
    return request.param


@pytest.mark.parametrize("res", [_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED, _abi.RESAMPLE_MULTINOMIAL,
                                 _abi.RESAMPLE_RESIDUAL])
def test_fat_parents_single_gpu(low_fat_threshold, sweep_path, res):
    m = models.linear_gaussian(r=0.0004)            # very sharp likelihood: a handful of parents take everything
    N, T = 8192 + 77, 7
    _, Y = O.simulate_data(m, T, 5)
    cfg = _abi.make_config(m, N, T, resampler=res)
    ro = O.sweep(cfg, Y, 9, mode=O.CANON)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    le = h.sweep(9)
    if h.last_sweep_launches() == 1:               # fused persistent kernel: no lists by construction
        assert sweep_path == "fused" and h.fat_counts().sum() == 0
    else:
        assert h.fat_counts()[1:].sum() > 0        # the deferred path really ran (also at the final step)
        assert h.fat_counts()[T] > 0
    assert_sweep_equal(cfg, ro, h, le)
    assert np.array_equal(h.final_states(), ro.x_hist[T - 1][ro.anc_hist[T]])


@pytest.mark.parametrize("sampler", [_abi.SAMPLER_PG, _abi.SAMPLER_PGAS])
def test_fat_parents_conditional(low_fat_threshold, sweep_path, sampler):
    m = models.linear_gaussian(r=0.0004)
    N, T = 8192, 6
    _, Y = O.simulate_data(m, T, 5)
    cfg = _abi.make_config(m, N, T, sampler=sampler, ess_threshold=1.0)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    ref = None
    for seed in (1, 2):
        ro = O.sweep(cfg, Y, seed, ref_traj=ref, mode=O.CANON)
        le = h.sweep(seed, ref_traj=ref)
        assert_sweep_equal(cfg, ro, h, le)
        slot_o, traj_o = O.pick_trajectory(cfg, seed, ro, mode=O.CANON)
        slot_g, traj_g = h.pick_trajectory()
        assert slot_g == slot_o and np.array_equal(traj_g, traj_o)
        ref = traj_o
    assert (h.fat_counts().sum() > 0) == (h.last_sweep_launches() > 1)


@pytest.mark.parametrize("world,res", [(2, _abi.RESAMPLE_SYSTEMATIC), (4, _abi.RESAMPLE_SYSTEMATIC),
                                       (4, _abi.RESAMPLE_MULTINOMIAL), (2, _abi.RESAMPLE_RESIDUAL)])
def test_fat_parents_sharded(low_fat_threshold, world, res):
    """Fat ranges cross shard boundaries: the owner of the parent pushes the entry into the list of
    every rank that owns some of the children (peer atomics + stores)."""
    m = models.linear_gaussian(r=0.0004)
    N, T = 8192 * 2, 6
    _, Y = O.simulate_data(m, T, 5)
    hs = make_ranks(m, N, T, Y, world, res)
    les = collective(hs, lambda h: h.sweep(4))
    ro = O.sweep(_abi.make_config(m, N, T, resampler=res), Y, 4, mode=O.CANON)
    assert sum(int(h.fat_counts().sum()) for h in hs) > 0
    assert_sharded_equal(hs, les, ro, N, T)


@pytest.mark.parametrize("kind", [_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED, _abi.RESAMPLE_RESIDUAL])
def test_fat_parents_operator_level(low_fat_threshold, kind):
    rng = np.random.default_rng(3)
    w = rng.random(5000) * 1e-6
    w[[17, 2048, 4999]] = [0.5, 0.3, 0.2]
    w /= w.sum()
    n = 20000
    got = _lib.resample(kind, w, n, key=5, ctr=2)
    want = O.resample(kind, w, n, key=5, step=2, mode=O.CANON)
    assert np.array_equal(got, want)
