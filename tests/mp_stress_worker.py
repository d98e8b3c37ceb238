"""Stress of the multi-GPU exchanges: thousands of back-to-back sharded sweeps
with two-slab history (keep_history = 0: peers read slab (t-1) % 2 while a fast rank may already be
in the next sweep), alternating well-spread and degenerate weights (observations near the particles
/ far outliers under a sharp likelihood), every sweep compared with the unsharded oracle: evidence
bit-equal, this rank's final ancestors and states equal.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29544 tests/mp_stress_worker.py [n_sweeps]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def observations(T, kind, rng):
    if kind == 0:                                   # near the prior / transition mean: ESS a good share of N
        return 0.4 + 0.05 * rng.normal(size=(T, 1))
    y = 0.4 + 0.05 * rng.normal(size=(T, 1))         # outliers: a handful of particles take everything
    y[rng.integers(0, T, size=max(1, T // 2))] += rng.choice([-1.5, 1.5])
    return y


def main():
    import torch
    import torch.distributed as dist

    import oracle as O
    from advancedps_b200 import _abi, models
    from advancedps_b200 import distributed as D

    n_sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else int(os.environ.get("APS_STRESS_SWEEPS", 5000))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = models.linear_gaussian(r=0.02)              # sharp likelihood
    T = 6
    sizes = [(256 * world, n_sweeps), (16384 * world, max(20, n_sweeps // 100))]
    t0 = time.time()
    total = 0
    for N, count in sizes:
        for res, thr in ((_abi.RESAMPLE_SYSTEMATIC, float("nan")), (_abi.RESAMPLE_STRATIFIED, 0.5)):
            rng = np.random.default_rng(1000 + N)
            Ys = [observations(T, k % 2, rng) for k in range(8)]
            h = D.create_sharded_handle(m, N, T, Ys[0], resampler=res, ess_threshold=thr, keep_history=False, device=local)
            cfg = _abi.make_config(m, N, T, resampler=res, ess_threshold=thr)
            lo, hi = D.shard_bounds(N, world, rank)
            n_here = count if res == _abi.RESAMPLE_SYSTEMATIC else max(10, count // 10)
            min_ess = 1e300
            for k in range(n_here):
                Y = Ys[k % 8]
                h.set_observations(Y)
                le = h.sweep(7000 + k)
                ro = O.sweep(cfg, Y, 7000 + k, mode=O.CANON)
                min_ess = min(min_ess, ro.ess[1:].min() / N)
                assert le == ro.logevidence, (N, res, k, le, ro.logevidence)
                if k % 16 == 0 or k == n_here - 1:   # the accessors synchronise: not on every sweep, so sweeps also run back to back
                    assert np.array_equal(h.ancestors(T + 1), ro.anc_hist[T][lo:hi]), (N, res, k)
                    assert np.array_equal(h.states(T), ro.x_hist[T - 1][lo:hi]), (N, res, k)
                    dist.barrier()                   # peers must not overwrite the slabs the accessors read
            total += n_here
            assert min_ess < 0.01, min_ess           # the degenerate regime really occurred
            dist.barrier()
            h.close()
    dist.barrier()
    if rank == 0:
        print(f"MP_STRESS_OK world={world} sweeps={total} seconds={time.time() - t0:.1f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
