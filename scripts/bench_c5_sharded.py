"""BASELINE configs[4] shape under torchrun: LG d=1, T=100, N = 1e6 x world sharded, SMC with each of
the four resamplers (and PGAS conditional sweeps). Device time per sweep, max over ranks."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from advancedps_b200 import _abi, models  # noqa: E402
from advancedps_b200 import distributed as D  # noqa: E402
import bench  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, T = 1_000_000 * world, 100
Y = bench.make_data()
nan = float("nan")
out = []


def timed(h, sweeps, cond=False):
    ms = []
    for k in range(sweeps):
        le = h.sweep(100 + k, ref_on_device=cond and k > 0)
        ms.append(h.last_sweep_ms())
        if cond:
            h.pick_trajectory(want_traj=False)
    t = torch.tensor([min(ms[1:])], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), le


for name, res in (("systematic", 3), ("stratified", 2), ("residual", 1), ("multinomial", 0)):
    h = D.create_sharded_handle(models.linear_gaussian(), N, T, Y, resampler=res, ess_threshold=nan, device=local)
    ms, le = timed(h, 4)
    out.append({"config": f"C5 SMC {name}", "n_gpus": world, "N": N, "T": T, "ms_per_sweep": ms,
                "particle_steps_per_s": N * T / (ms * 1e-3), "logevidence": le})
    dist.barrier()
    h.close()
sv = models.stochastic_volatility()
Ysv = np.random.default_rng(0).normal(size=(T, 1)) * 0.7
h = D.create_sharded_handle(sv, N, T, Ysv, ess_threshold=1.0, device=local, sampler=_abi.SAMPLER_PGAS)
ms, le = timed(h, 4, cond=True)
out.append({"config": "SV PGAS conditional (sharded)", "n_gpus": world, "N": N, "T": T, "ms_per_sweep": ms,
            "particle_steps_per_s": N * T / (ms * 1e-3), "logevidence": le})
dist.barrier()
h.close()
if rank == 0:
    for r in out:
        print(json.dumps(r), flush=True)
dist.destroy_process_group()
