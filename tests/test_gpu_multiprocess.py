"""One process per GPU over CUDA IPC (the production multi-GPU path): needs >= 2 GPUs, skipped on a
single-GPU box (tests/test_gpu_sharded.py covers the same kernels with emulated ranks there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_torchrun_sharded_equals_oracle():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mp_sharded_worker.py")],
        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and f"MP_SHARDED_OK world={world}" in out.stdout, out.stdout[-4000:]
