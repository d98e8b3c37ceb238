"""One process per GPU over CUDA IPC (the production multi-GPU path): needs >= 2 GPUs, skipped on a
single-GPU box (tests/test_gpu_sharded.py covers the same kernels with emulated ranks there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("predraw", [False, True])
def test_torchrun_sharded_equals_oracle(predraw):
    """predraw: the same parity run with the state draws forced onto the parallel graph branch
    (k_draw_normals; by default that path needs >= 4e5 particles per GPU, more than these cases have)."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29534" if predraw else "29533", os.path.join(ROOT, "tests", "mp_sharded_worker.py")],
        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
        env=dict(os.environ, **({"APS_PREDRAW": "1"} if predraw else {})))
    assert out.returncode == 0 and f"MP_SHARDED_OK world={world}" in out.stdout, out.stdout[-4000:]
