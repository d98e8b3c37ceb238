"""Sharded particle sweep: one process per GPU, particles split into contiguous blocks.

``torch.distributed`` is plumbing only: it carries the CUDA-IPC blobs once at set-up. During a
sweep the ranks talk through peer-mapped device memory inside the kernels (log-weight maximum,
integer weight totals, a barrier after the ancestor scatter) -- see csrc/aps_device.cuh.
"""
import numpy as np

from . import _abi, _lib


def shard_bounds(n_global, world, rank):
    """Global slots [lo, hi) owned by ``rank`` (contiguous blocks)."""
    if n_global % (world * 32):
        raise ValueError("n_particles must be a multiple of 32 * world_size")
    nl = n_global // world
    return rank * nl, (rank + 1) * nl


def combine_totals(totals, rank):
    """Rank-order combination of the per-shard integer weight totals: (global total, exclusive
    offset of ``rank``). Integers, so the result is identical on every rank."""
    totals = [int(t) for t in totals]
    return sum(totals), sum(totals[:rank])


def create_sharded_handle(model, n_global, n_steps, Y, resampler=_abi.RESAMPLE_SYSTEMATIC,
                          ess_threshold=float("nan"), keep_history=True, device=None, group=None,
                          sampler=_abi.SAMPLER_SMC):
    """Build this rank's handle of a sharded SMC / PG / PGAS sweep and attach the peers (collective
    call). ``sweep`` and ``pick_trajectory`` on the returned handle are collective too: every rank
    calls them with the same arguments."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.cuda.current_device()
    shard_bounds(n_global, world, rank)
    cfg = _abi.make_config(model, n_global, n_steps, sampler=sampler, resampler=resampler,
                           ess_threshold=ess_threshold, keep_history=keep_history, device=device,
                           rank=rank, world_size=world)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    blobs = [None] * world
    dist.all_gather_object(blobs, h.ipc_export().tobytes(), group=group)
    h.ipc_import([np.frombuffer(b, dtype=np.uint8) for b in blobs])
    dist.barrier(group=group)
    return h
