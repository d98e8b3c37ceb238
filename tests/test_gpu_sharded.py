"""Sharded sweep == single-GPU sweep == oracle. Two (or four) ranks are emulated inside one
process on one GPU (threads + raw peer pointers), which exercises the same kernels, mailbox
exchanges and ancestor scatter as one process per GPU over CUDA IPC."""
import threading

import numpy as np
import pytest

import oracle as O
from advancedps_b200 import _abi, _lib, models

pytestmark = pytest.mark.gpu


def run_sharded(model, N, T, Y, seeds, world, resampler=_abi.RESAMPLE_SYSTEMATIC, thr=float("nan")):
    hs = []
    for r in range(world):
        cfg = _abi.make_config(model, N, T, resampler=resampler, ess_threshold=thr, rank=r, world_size=world)
        h = _lib.Handle(cfg)
        h.set_observations(Y)
        hs.append(h)
    blobs = [h.ipc_export() for h in hs]
    for h in hs:
        h.ipc_import(blobs)
    out = []
    for seed in seeds:
        res = [None] * world

        def work(r):
            try:
                res[r] = hs[r].sweep(seed)
            except Exception as e:  # noqa: BLE001
                res[r] = e

        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        [t.start() for t in th]
        [t.join() for t in th]
        for r in res:
            if isinstance(r, Exception):
                raise r
        out.append(res)
    return hs, out


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
    assert all(le == ro.logevidence for le in out[1])
    nl = N // world
    for t in range(1, T + 1):
        x = np.concatenate([h.states(t) for h in hs])
        assert np.array_equal(x, ro.x_hist[t - 1]), f"states differ at t={t}"
    for t in range(2, T + 2):
        a = np.concatenate([h.ancestors(t) for h in hs])
        assert np.array_equal(a, ro.anc_hist[t - 1]), f"ancestors differ at t={t}"
    w = np.concatenate([h.weights() for h in hs])
    assert np.array_equal(w, ro.final_w)
    logz, ess, rs = hs[0].step_stats()
    assert np.array_equal(logz, ro.logz) and np.array_equal(ess, ro.ess) and np.array_equal(rs, ro.resampled)
    assert nl * world == N


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
