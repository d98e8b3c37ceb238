"""Sharded sweep == single-GPU sweep == oracle. Two (or four) ranks are emulated inside one
process on one GPU (threads + raw peer pointers), which exercises the same kernels, mailbox
exchanges and ancestor scatter as one process per GPU over CUDA IPC."""
import threading

import numpy as np
import pytest

import oracle as O
from advancedps_b200 import _abi, _lib, models

pytestmark = pytest.mark.gpu


def make_ranks(model, N, T, Y, world, resampler=_abi.RESAMPLE_SYSTEMATIC, thr=float("nan"), sampler=_abi.SAMPLER_SMC):
    hs = []
    for r in range(world):
        cfg = _abi.make_config(model, N, T, sampler=sampler, resampler=resampler, ess_threshold=thr, rank=r,
                               world_size=world)
        h = _lib.Handle(cfg)
        h.set_observations(Y)
        hs.append(h)
    blobs = [h.ipc_export() for h in hs]
    for h in hs:
        h.ipc_import(blobs)
    return hs


def collective(hs, fn):
    """Run fn(handle) on every emulated rank concurrently (the calls spin on each other)."""
    res = [None] * len(hs)

    def work(r):
        try:
            res[r] = fn(hs[r])
        except Exception as e:  # noqa: BLE001
            res[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(len(hs))]
    [t.start() for t in th]
    [t.join() for t in th]
    for r in res:
        if isinstance(r, Exception):
            raise r
    return res


def run_sharded(model, N, T, Y, seeds, world, resampler=_abi.RESAMPLE_SYSTEMATIC, thr=float("nan")):
    hs = make_ranks(model, N, T, Y, world, resampler, thr)
    return hs, [collective(hs, lambda h, seed=seed: h.sweep(seed)) for seed in seeds]


def assert_sharded_equal(hs, les, ro, N, T):
    assert all(le == ro.logevidence for le in les)
    for t in range(1, T + 1):
        x = np.concatenate([h.states(t) for h in hs])
        assert np.array_equal(x, ro.x_hist[t - 1]), f"states differ at t={t}"
    for t in range(2, T + 2):
        a = np.concatenate([h.ancestors(t) for h in hs])
        bad = np.nonzero(a != ro.anc_hist[t - 1])[0]
        assert bad.size == 0, f"{bad.size} ancestors differ at t={t}, first {bad[:5]}"
    assert np.array_equal(np.concatenate([h.weights() for h in hs]), ro.final_w)
    assert np.array_equal(np.concatenate([h.logweights() for h in hs]), ro.final_logw)
    for h in hs:
        logz, ess, rs = h.step_stats()
        assert np.array_equal(logz, ro.logz) and np.array_equal(ess, ro.ess) and np.array_equal(rs, ro.resampled)


@pytest.mark.parametrize("world,N,T,res,thr", [
    (2, 4096, 6, _abi.RESAMPLE_SYSTEMATIC, float("nan")),
    (2, 20480, 9, _abi.RESAMPLE_SYSTEMATIC, 0.5),
    (4, 8192 * 3, 7, _abi.RESAMPLE_SYSTEMATIC, float("nan")),
    (2, 6400, 5, _abi.RESAMPLE_STRATIFIED, float("nan")),
])
def test_sharded_equals_oracle(world, N, T, res, thr):
    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, T, 0xDA7A0005)
    hs, out = run_sharded(m, N, T, Y, [11, 12], world, res, thr)
    cfg = _abi.make_config(m, N, T, resampler=res, ess_threshold=thr)
    ro = O.sweep(cfg, Y, 12, mode=O.CANON)   # the handles hold the second sweep
    assert_sharded_equal(hs, out[1], ro, N, T)


def test_sharded_skewed_weights_cross_rank_children():
    """A sharp likelihood puts most offspring on a few parents, so children land on other ranks."""
    m = models.linear_gaussian(r=0.01)
    N, T, world = 8192, 5, 4
    _, Y = O.simulate_data(m, T, 3)
    hs, out = run_sharded(m, N, T, Y, [5], world)
    ro = O.sweep(_abi.make_config(m, N, T), Y, 5, mode=O.CANON)
    assert out[0][0] == ro.logevidence
    a = np.concatenate([h.ancestors(T + 1) for h in hs])
    assert np.array_equal(a, ro.anc_hist[T])
    nl = N // world
    owner_of_parent = a // nl
    owner_of_child = np.arange(N) // nl
    assert np.any(owner_of_parent != owner_of_child)  # the scatter really crossed shard boundaries


@pytest.mark.parametrize("world,N,T,res,thr", [
    (2, 4096, 6, _abi.RESAMPLE_MULTINOMIAL, float("nan")),
    (4, 8192 * 3, 5, _abi.RESAMPLE_MULTINOMIAL, 0.5),
    (2, 6400, 6, _abi.RESAMPLE_RESIDUAL, float("nan")),
    (4, 8192 * 3, 5, _abi.RESAMPLE_RESIDUAL, 0.5),
    (8, 8192 * 2, 4, _abi.RESAMPLE_RESIDUAL, float("nan")),
    (8, 8192 * 2, 4, _abi.RESAMPLE_SYSTEMATIC, float("nan")),
])
def test_sharded_multinomial_residual(world, N, T, res, thr):
    """configs[4]: the resampler sweep {systematic, stratified, residual, multinomial}, sharded."""
    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, T, 0xDA7A0005)
    hs, out = run_sharded(m, N, T, Y, [3, 4], world, res, thr)
    ro = O.sweep(_abi.make_config(m, N, T, resampler=res, ess_threshold=thr), Y, 4, mode=O.CANON)
    assert_sharded_equal(hs, out[1], ro, N, T)
    # collect(pc): final particle set gathered through the peer-mapped stores
    xf = np.concatenate(collective(hs, lambda h: h.final_states()))
    assert np.array_equal(xf, ro.x_hist[T - 1][ro.anc_hist[T]])


@pytest.mark.parametrize("sampler,world,res", [
    (_abi.SAMPLER_PG, 2, _abi.RESAMPLE_SYSTEMATIC),
    (_abi.SAMPLER_PGAS, 2, _abi.RESAMPLE_SYSTEMATIC),
    (_abi.SAMPLER_PGAS, 4, _abi.RESAMPLE_SYSTEMATIC),
    (_abi.SAMPLER_PG, 4, _abi.RESAMPLE_MULTINOMIAL),
    (_abi.SAMPLER_PGAS, 2, _abi.RESAMPLE_RESIDUAL),
])
def test_sharded_conditional_sweeps(sampler, world, res):
    """PG / PGAS sharded: the reference particle is the globally last slot (last rank), the PGAS
    ancestor draw and the final pick are exchanges between the ranks, trajectories are walked
    through the peer-mapped genealogy (src/smc.jl:101-129, src/pgas.jl:113-128)."""
    m = models.stochastic_volatility() if sampler == _abi.SAMPLER_PGAS else models.linear_gaussian()
    N, T = 8192, 9
    thr = 1.0 if sampler == _abi.SAMPLER_PGAS else 0.5
    _, Y = O.simulate_data(m, T, 0xDA7A0004)
    cfg = _abi.make_config(m, N, T, sampler=sampler, resampler=res, ess_threshold=thr)
    hs = make_ranks(m, N, T, Y, world, res, thr, sampler)
    ref = None
    for seed in [1, 2, 3]:
        ro = O.sweep(cfg, Y, seed, ref_traj=ref, mode=O.CANON)
        les = collective(hs, lambda h: h.sweep(seed, ref_traj=ref))
        assert_sharded_equal(hs, les, ro, N, T)
        slot_o, traj_o = O.pick_trajectory(cfg, seed, ro, mode=O.CANON)
        picks = collective(hs, lambda h: h.pick_trajectory())
        for slot_g, traj_g in picks:   # every rank holds the same global slot and trajectory
            assert slot_g == slot_o
            assert np.array_equal(traj_g, traj_o)
        if ref is not None:
            for t in range(1, T + 1):
                assert np.array_equal(hs[-1].states(t)[-1], ref[t - 1])
        ref = traj_o
    ro = O.sweep(cfg, Y, 77, ref_traj=ref, mode=O.CANON)
    les = collective(hs, lambda h: h.sweep(77, ref_on_device=True))
    assert all(le == ro.logevidence for le in les)
    nl = N // world
    tr = collective(hs, lambda h: h.trajectory(nl - 1))
    for r, t_r in enumerate(tr):
        assert np.array_equal(t_r, O.trajectory(cfg, (r + 1) * nl - 1, ro))


@pytest.mark.parametrize("world,N,T", [(2, 64, 1), (2, 64, 2), (4, 128, 3), (8, 256, 2)])
def test_sharded_minimal_sizes(world, N, T):
    """Smallest legal shards (32 slots per rank), one- and two-step sweeps."""
    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, T, 1)
    hs, out = run_sharded(m, N, T, Y, [8], world)
    ro = O.sweep(_abi.make_config(m, N, T), Y, 8, mode=O.CANON)
    assert_sharded_equal(hs, out[0], ro, N, T)


def test_sharded_without_history_even_and_odd_steps():
    """keep_history = 0 keeps two state / ancestor slabs: peers read slab (t-1) % 2 while a fast rank
    may already be in the next sweep -- the exchange at the start of k_propagate(t = 1) orders them."""
    m = models.linear_gaussian()
    for T in (6, 7):
        _, Y = O.simulate_data(m, T, 2)
        hs = []
        for r in range(2):
            cfg = _abi.make_config(m, 8192, T, keep_history=False, rank=r, world_size=2)
            h = _lib.Handle(cfg)
            h.set_observations(Y)
            hs.append(h)
        blobs = [h.ipc_export() for h in hs]
        [h.ipc_import(blobs) for h in hs]
        for seed in (1, 2, 3, 4):
            les = collective(hs, lambda h: h.sweep(seed))
            ro = O.sweep(_abi.make_config(m, 8192, T), Y, seed, mode=O.CANON)
            assert all(le == ro.logevidence for le in les)
            assert np.array_equal(np.concatenate([h.weights() for h in hs]), ro.final_w)


def test_sharded_not_normalisable_raises_on_every_rank():
    m = models.constant_loglik()
    hs = make_ranks(m, 4096, 2, np.full((2, 1), -np.inf), 2)

    def run(h):
        try:
            h.sweep(1)
            return None
        except _lib.ApsError as e:
            return e.code

    assert collective(hs, run) == [_abi.ERR_WEIGHTS, _abi.ERR_WEIGHTS]
