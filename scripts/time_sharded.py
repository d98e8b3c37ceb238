"""torchrun target: the headline sharded sweep (1e6 particles per rank), device time per sweep; with a
-DAPS_TIMELINE=1 build and APS_DEBUG_MULTI=16 the in-graph timeline of K1/K2/K3 on every rank, with
APS_DEBUG_SPIN=1 the block-0 wait cycles per exchange."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from advancedps_b200 import _abi, models, distributed as D
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = D.create_sharded_handle(models.linear_gaussian(), 1_000_000 * world, 100, bench.make_data(), device=local)
for k in range(4): h.sweep(1 + k)
ms = []
for k in range(8):
    h.sweep(10 + k); ms.append(h.last_sweep_ms())
dist.barrier()
if rank == 0: print("world", world, "ms/sweep min %.3f med %.3f" % (min(ms), sorted(ms)[4]), flush=True)
dist.destroy_process_group()
