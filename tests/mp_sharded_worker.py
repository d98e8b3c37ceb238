"""Worker of tests/test_gpu_multiprocess.py: one process per GPU (torchrun), peers attached over
CUDA IPC. Every rank runs the sharded sweep for all four resamplers and for PG / PGAS and checks
its shard against the unsharded oracle (CANON mode) bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/mp_sharded_worker.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist

    import oracle as O
    from advancedps_b200 import _abi, models
    from advancedps_b200 import distributed as D

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nan = float("nan")
    N, T = 8192 * world, 8
    checked = 0

    # ---- SMC, the four resamplers (BASELINE configs[4])
    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, T, 0xDA7A0005)
    for res, thr in ((_abi.RESAMPLE_SYSTEMATIC, nan), (_abi.RESAMPLE_STRATIFIED, nan), (_abi.RESAMPLE_RESIDUAL, nan),
                     (_abi.RESAMPLE_MULTINOMIAL, nan), (_abi.RESAMPLE_MULTINOMIAL, 0.5)):
        h = D.create_sharded_handle(m, N, T, Y, resampler=res, ess_threshold=thr, device=local)
        for seed in (5, 6):
            le = h.sweep(seed)
        ro = O.sweep(_abi.make_config(m, N, T, resampler=res, ess_threshold=thr), Y, 6, mode=O.CANON)
        lo, hi = D.shard_bounds(N, world, rank)
        assert le == ro.logevidence, (res, le, ro.logevidence)
        for t in range(1, T + 1):
            assert np.array_equal(h.states(t), ro.x_hist[t - 1][lo:hi]), f"res {res}: states differ at t={t}"
        for t in range(2, T + 2):
            assert np.array_equal(h.ancestors(t), ro.anc_hist[t - 1][lo:hi]), f"res {res}: ancestors differ at t={t}"
        assert np.array_equal(h.weights(), ro.final_w[lo:hi])
        assert np.array_equal(h.final_states(), ro.x_hist[T - 1][ro.anc_hist[T]][lo:hi])
        dist.barrier()
        h.close()
        checked += 1

    # ---- configs[4] at its own T = 100: the resampler sweep, every field against the unsharded oracle
    T5 = 100
    _, Y5 = O.simulate_data(m, T5, 0xDA7A0005)
    for res in (_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED, _abi.RESAMPLE_RESIDUAL, _abi.RESAMPLE_MULTINOMIAL):
        h = D.create_sharded_handle(m, N, T5, Y5, resampler=res, device=local)
        le = h.sweep(4321)
        ro = O.sweep(_abi.make_config(m, N, T5, resampler=res), Y5, 4321, mode=O.CANON)
        lo, hi = D.shard_bounds(N, world, rank)
        assert le == ro.logevidence, ("c5", res, le, ro.logevidence)
        logz, ess, rs = h.step_stats()
        assert np.array_equal(logz, ro.logz) and np.array_equal(ess, ro.ess) and np.array_equal(rs, ro.resampled)
        for t in range(1, T5 + 1):
            assert np.array_equal(h.states(t), ro.x_hist[t - 1][lo:hi]), f"c5 res {res}: states differ at t={t}"
        for t in range(2, T5 + 2):
            assert np.array_equal(h.ancestors(t), ro.anc_hist[t - 1][lo:hi]), f"c5 res {res}: ancestors differ at t={t}"
        assert np.array_equal(h.weights(), ro.final_w[lo:hi])
        dist.barrier()
        h.close()
        checked += 1

    # ---- PG / PGAS: conditional sweeps, collective pick
    for sampler, model, thr in ((_abi.SAMPLER_PG, models.linear_gaussian(), 0.5),
                                (_abi.SAMPLER_PGAS, models.stochastic_volatility(), 1.0)):
        _, Y = O.simulate_data(model, T, 0xDA7A0004)
        cfg = _abi.make_config(model, N, T, sampler=sampler, ess_threshold=thr)
        h = D.create_sharded_handle(model, N, T, Y, ess_threshold=thr, device=local, sampler=sampler)
        lo, hi = D.shard_bounds(N, world, rank)
        ref = None
        for seed in (1, 2, 3):
            ro = O.sweep(cfg, Y, seed, ref_traj=ref, mode=O.CANON)
            le = h.sweep(seed, ref_on_device=ref is not None)
            assert le == ro.logevidence, (sampler, seed, le, ro.logevidence)
            for t in range(2, T + 2):
                assert np.array_equal(h.ancestors(t), ro.anc_hist[t - 1][lo:hi]), f"sampler {sampler}: ancestors differ at t={t}"
            slot_o, traj_o = O.pick_trajectory(cfg, seed, ro, mode=O.CANON)
            dist.barrier()  # accessors above read this rank's stores; peers must not start the next sweep early
            slot_g, traj_g = h.pick_trajectory()
            assert slot_g == slot_o and np.array_equal(traj_g, traj_o)
            ref = traj_o
        dist.barrier()
        h.close()
        checked += 1

    # ---- the sampler surface, sharded: same chain as the single-handle oracle recursion
    from advancedps_b200 import sampler as S

    sv = models.stochastic_volatility()
    _, Y = O.simulate_data(sv, T, 0xDA7A0004)
    tssm = S.TracedSSM(sv, Y)
    chain = D.sample(np.random.default_rng(5), tssm, S.PGAS(N), 3)
    rng0 = np.random.default_rng(5)
    cfg = _abi.make_config(sv, N, T, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
    ref = None
    for smp in chain:
        key = int(rng0.integers(0, 2**64, dtype=np.uint64))
        ro = O.sweep(cfg, Y, key, ref_traj=ref, mode=O.CANON)
        _, ref = O.pick_trajectory(cfg, key, ro, mode=O.CANON)
        assert smp.logevidence == ro.logevidence and np.array_equal(smp.trajectory.model.X, ref)
    smc = D.sample(np.random.default_rng(6), S.TracedSSM(models.linear_gaussian(), Y), S.SMC(N, S.resample_systematic))
    tot = torch.tensor([float(smc.weights.sum())], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    assert abs(float(tot.item()) - 1.0) < 1e-12
    checked += 2

    dist.barrier()
    if rank == 0:
        print(f"MP_SHARDED_OK world={world} cases={checked}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
