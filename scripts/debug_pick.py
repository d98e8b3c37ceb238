import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'oracle')); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np
import oracle as O
from advancedps_b200 import _abi, _lib, models
from test_gpu_sharded import make_ranks, collective
m = models.linear_gaussian(); N,T,world=8192,9,2
sampler=_abi.SAMPLER_PG; thr=0.5
_, Y = O.simulate_data(m, T, 0xDA7A0004)
cfg = _abi.make_config(m, N, T, sampler=sampler, ess_threshold=thr)
hs = make_ranks(m, N, T, Y, world, 3, thr, sampler)
ref=None
for seed in [1,2,3]:
    ro = O.sweep(cfg, Y, seed, ref_traj=ref, mode=O.CANON)
    les = collective(hs, lambda h: h.sweep(seed, ref_traj=ref))
    print("seed", seed, "logev", les, ro.logevidence, "resampled", ro.resampled)
    for r,h in enumerate(hs):
        logz, ess, rs = h.step_stats(); print("  rank", r, "res", rs, "w sum", h.weights().sum())
    slot_o, traj_o = O.pick_trajectory(cfg, seed, ro, mode=O.CANON)
    def pk(h):
        try: return h.pick_trajectory()[0]
        except Exception as e: return repr(e)
    print("  oracle slot", slot_o, "gpu", collective(hs, pk))
    ref = traj_o
