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


def draw_partition(n_draws, world, rank):
    """Philox blocks [p0, p1) whose i.i.d. draws (two per block) ``rank`` makes in the sharded
    multinomial / residual step -- each draw is made once and routed to the rank that owns its weight
    range (csrc: k_multi_route)."""
    npairs = (int(n_draws) + 1) // 2
    pp = (npairs + world - 1) // world
    return min(rank * pp, npairs), min((rank + 1) * pp, npairs)


def owner_of(tau, totals):
    """Rank whose weight range [sum(totals[:r]), sum(totals[:r+1])) holds the integer position ``tau``,
    and the position relative to that range."""
    off = 0
    for r, t in enumerate(totals):
        if tau < off + int(t):
            return r, tau - off
        off += int(t)
    raise ValueError("tau beyond the weight total")


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


# ------------------------------------------------------------------ sampler surface, sharded
_sharded = {}


def _shared_key(rng, group=None):
    """One rand(rng, UInt64) drawn on rank 0 and broadcast: every rank must seed the sweep alike."""
    import torch.distributed as dist

    from .sampler import _draw_key

    box = [_draw_key(rng) if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return int(box[0])


def _sharded_handle(model, sampler, group=None):
    from .sampler import _resampler_config

    kind, thr = _resampler_config(sampler.resampler)
    key = (id(model), sampler.kind, sampler.nparticles, kind, repr(thr))
    h = _sharded.get(key)
    if h is None:
        if _sharded:
            import torch.distributed as dist

            dist.barrier(group=group)  # peers still map the old handle's stores: drop them together
            _sharded.clear()
        h = create_sharded_handle(model.model, sampler.nparticles, model.Y.shape[0], model.Y, resampler=kind,
                                  ess_threshold=thr, group=group, sampler=sampler.kind)
        h._model_keepalive = model
        _sharded[key] = h
    else:
        h.set_observations(model.Y)
    return h


def sample(rng, model, sampler, n_iter=None, group=None):
    """Sharded counterpart of ``sampler.sample`` (collective: every rank calls it with the same
    arguments; only rank 0's ``rng`` is consumed). ``SMC`` -> SMCSample whose ``weights`` are this
    rank's shard; ``PG`` / ``PGAS`` with ``n_iter`` -> list of PGSample (identical on every rank)."""
    from . import sampler as S

    if isinstance(sampler, S.SMC):
        h = _sharded_handle(model, sampler, group)
        logev = h.sweep(_shared_key(rng, group))
        return S.SMCSample(h, model, h.weights(pinned=True), logev)
    if n_iter is None:
        raise TypeError("sample(rng, model, PG|PGAS, n_iter): n_iter is required")
    out, state = [], None
    for _ in range(int(n_iter)):
        smp, state = step(rng, model, sampler, state, group=group)
        out.append(smp)
    return out


def step(rng, model, sampler, state=None, group=None):
    """Sharded counterpart of ``sampler.step`` (src/smc.jl:101-129): one conditional sweep over all
    ranks and the collective pick; every rank returns the same PGSample."""
    from . import sampler as S

    h = _sharded_handle(model, sampler, group)
    key = _shared_key(rng, group)
    if state is None:
        logev = h.sweep(key)
    elif state._on_device(h):
        logev = h.sweep(key, ref_on_device=True)
    else:
        logev = h.sweep(key, ref_traj=state.trajectory.model.X)
    _, traj = h.pick_trajectory()
    tr = S.Trace(S.TracedSSM(model.model, model.Y, traj))
    return S.PGSample(tr, logev), S.PGState(tr, h)
