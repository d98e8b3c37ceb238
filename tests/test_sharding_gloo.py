"""N>1 path on CPU: two processes over the gloo backend replay the sharded resampling step
(shard maxima -> global max, shard totals -> rank-order offsets, per-shard child ranges, scatter by
owner rank) with the oracle as the compute, and must reproduce the unsharded oracle exactly.
Covers the host-side sharding logic in advancedps.jl_b200/distributed.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_global, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from advancedps_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)                      # same stream on every rank
        logw_all = -0.5 * (1.7 * rng.normal(size=n_global)) ** 2
        lo, hi = D.shard_bounds(n_global, world, rank)
        logw = logw_all[lo:hi]
        # exchange 1: all-reduce(max)
        m = torch.tensor([logw.max()], dtype=torch.float64)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        gmax = float(m.item())
        # exchange 2: all-gather of the integer totals, combined in rank order
        q, Qr = O.quantise_shard(logw, gmax, n_global)
        tot = [None] * world
        dist.all_gather_object(tot, int(Qr))
        Q, offset = D.combine_totals(tot, rank)
        # this rank's parents -> child ranges in global child index space (exact integers)
        key, step, n = 77, 3, n_global
        w0, _ = O.philox2x64(0, (step << 16) | (1 << 8), key)
        R = ((w0 >> 11) * Q + (1 << 53) - 1) >> 53

        def K(Cv):  # children with threshold at or below cumulative weight Cv
            v = Cv * n - R
            return 0 if v < 0 else min(n, v // Q + 1)

        C = offset
        klo = 0 if rank == 0 else K(offset)
        mine = []
        for j in range(hi - lo):
            C += int(q[j])
            khi = K(C)
            mine.extend((i, lo + j) for i in range(klo, khi))  # (child slot, global parent)
            klo = khi
        # scatter by owner rank of the child slot
        nl = n_global // world
        out = [[p for p in mine if p[0] // nl == r] for r in range(world)]
        gathered = [None] * world  # gloo: object collective for the ragged lists
        dist.all_gather_object(gathered, out)
        anc_local = np.full(nl, -1, dtype=np.int64)
        for src in range(world):
            for child, parent in gathered[src][rank]:
                anc_local[child - lo] = parent
        assert (anc_local >= 0).all()
        # reference: the unsharded oracle walk on the full integer-weight vector
        qa, _, Qa = O.quantise_logw(logw_all)
        assert Qa == Q
        refq = np.zeros(n_global, dtype=np.int64)
        import ctypes as Ct
        O._chk(O.lib().orc_resample_systematic_canon(O._ptr(qa), Ct.c_int64(n_global), Ct.c_int64(n_global),
                                                     Ct.c_uint64(key), Ct.c_uint64(step), O._ptr(refq)))
        ok = np.array_equal(anc_local, refq[lo:hi] - 1)
        np.save(os.path.join(tmpdir, f"ok_{rank}.npy"), np.array([int(ok), Q % (1 << 62)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global", [(2, 4096), (2, 64 * 1000)])
def test_two_rank_sharded_resample_matches_unsharded(tmp_path, world, n_global):
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(world, port, n_global, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"ok_{r}.npy") for r in range(world)]
    assert all(int(r[0]) == 1 for r in res)
    assert len({int(r[1]) for r in res}) == 1  # every rank derived the same global total


def test_shard_bounds_and_totals():
    sys.path.insert(0, ROOT)
    from advancedps_b200 import distributed as D

    assert D.shard_bounds(8_000_000, 8, 3) == (3_000_000, 4_000_000)
    with pytest.raises(ValueError):
        D.shard_bounds(1000, 8, 0)
    assert D.combine_totals([5, 7, 11], 2) == (23, 12)
    assert D.combine_totals([2**61, 2**61], 0) == (2**62, 0)


def _worker_multinomial(rank, world, port, n_global, tmpdir):
    """Sharded multinomial step as the kernels do it (round 2): every draw is made ONCE -- rank r makes
    the draws of its share of the Philox blocks (two draws per block) -- and ROUTED to the rank whose
    weight range holds it; the owner histograms what it received over its own parents; offspring
    totals are exchanged so that children are laid out grouped by parent in global parent order.
    Must equal the unsharded oracle's draw list."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ctypes as Ct

    import oracle as O
    from advancedps_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(321)
        logw_all = -0.5 * (2.0 * rng.normal(size=n_global)) ** 2
        lo, hi = D.shard_bounds(n_global, world, rank)
        m = torch.tensor([logw_all[lo:hi].max()], dtype=torch.float64)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        q, Qr = O.quantise_shard(logw_all[lo:hi], float(m.item()), n_global)
        tot = [None] * world
        dist.all_gather_object(tot, int(Qr))
        Q, offset = D.combine_totals(tot, rank)
        cum = np.cumsum(q.astype(object))                      # exact python ints, local inclusive sums
        key, step, n = 99, 4, n_global - 1                     # n = N - 1 draws: the PG case (odd count, last block half used)
        p0, p1 = D.draw_partition(n, world, rank)
        send = [[] for _ in range(world)]                      # positions relative to the owner's range
        for p in range(p0, p1):                                # this rank's share of the draws
            w = O.philox2x64(p, (step << 16) | (1 << 8), key)
            for h in range(2):
                if 2 * p + h < n:
                    o, rel = D.owner_of(((w[h] >> 11) * Q) >> 53, tot)
                    send[o].append(rel)
        routed = [None] * world                                # the all-to-all of k_multi_route
        dist.all_gather_object(routed, send)
        counts = np.zeros(hi - lo, dtype=np.int64)
        for src in range(world):
            for rel in routed[src][rank]:
                counts[int(np.searchsorted(cum, rel, side="right"))] += 1
        ctot = [None] * world
        dist.all_gather_object(ctot, int(counts.sum()))
        assert sum(ctot) == n
        child_off = sum(ctot[:rank])                           # children of lower ranks' parents come first
        mine = np.repeat(np.arange(lo, hi), counts)            # global parent ids of children child_off..
        allc = [None] * world
        dist.all_gather_object(allc, (child_off, mine))
        anc = np.full(n, -1, dtype=np.int64)
        for off, par in allc:
            anc[off:off + len(par)] = par
        # unsharded oracle: draw list -> counts -> grouped by parent (src/container.jl:185-217)
        qa, _, Qa = O.quantise_logw(logw_all)
        idx = np.zeros(n, dtype=np.int64)
        O._chk(O.lib().orc_resample_multinomial_canon(O._ptr(qa), Ct.c_int64(n_global), Ct.c_int64(n), Ct.c_uint64(key),
                                                      Ct.c_uint64(step), O._ptr(idx)))
        want = np.repeat(np.arange(n_global), np.bincount(idx - 1, minlength=n_global))
        np.save(os.path.join(tmpdir, f"okm_{rank}.npy"), np.array([int(Qa == Q and np.array_equal(anc, want))]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_multinomial_matches_unsharded(tmp_path):
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker_multinomial, args=(2, port, 2048, str(tmp_path)), nprocs=2, join=True)
    assert all(int(np.load(tmp_path / f"okm_{r}.npy")[0]) == 1 for r in range(2))


def test_draw_partition_and_owner():
    sys.path.insert(0, ROOT)
    from advancedps_b200 import distributed as D

    for n, world in ((10, 2), (11, 4), (1, 8), (8_000_000, 8), (7_999_999, 8)):
        parts = [D.draw_partition(n, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == (n + 1) // 2
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))   # contiguous, no draw twice
    assert D.owner_of(0, [5, 0, 7]) == (0, 0)
    assert D.owner_of(5, [5, 0, 7]) == (2, 0)      # an empty range owns nothing
    assert D.owner_of(11, [5, 0, 7]) == (2, 6)
    with pytest.raises(ValueError):
        D.owner_of(12, [5, 0, 7])
