"""Stress test of the sharded exchanges (mailbox sequence numbers, two-slab history, degenerate and
well-spread weights alternating, thousands of sweeps back to back). On a multi-GPU box: one process
per GPU over CUDA IPC, 5000 sweeps (tests/mp_stress_worker.py; run log in profiles/). On a one-GPU
box: two emulated ranks, 300 sweeps."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stress_one_process_per_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    sweeps = os.environ.get("APS_STRESS_SWEEPS", "5000")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29544", os.path.join(ROOT, "tests", "mp_stress_worker.py"), sweeps],
        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert out.returncode == 0 and f"MP_STRESS_OK world={world}" in out.stdout, out.stdout[-4000:]


def test_stress_emulated_two_ranks():
    import oracle as O
    from advancedps_b200 import _abi, _lib, models
    from mp_stress_worker import observations
    from test_gpu_sharded import collective

    m = models.linear_gaussian(r=0.02)
    N, T, world = 512, 6, 2
    rng = np.random.default_rng(5)
    Ys = [observations(T, k % 2, rng) for k in range(8)]
    hs = []
    for r in range(world):
        h = _lib.Handle(_abi.make_config(m, N, T, keep_history=False, rank=r, world_size=world))
        h.set_observations(Ys[0])
        hs.append(h)
    blobs = [h.ipc_export() for h in hs]
    [h.ipc_import(blobs) for h in hs]
    cfg = _abi.make_config(m, N, T)
    min_ess = 1.0
    for k in range(300):
        Y = Ys[k % 8]
        [h.set_observations(Y) for h in hs]
        les = collective(hs, lambda h: h.sweep(7000 + k))
        ro = O.sweep(cfg, Y, 7000 + k, mode=O.CANON)
        min_ess = min(min_ess, ro.ess[1:].min() / N)
        assert all(le == ro.logevidence for le in les), k
        assert np.array_equal(np.concatenate([h.ancestors(T + 1) for h in hs]), ro.anc_hist[T]), k
    assert min_ess < 0.02
