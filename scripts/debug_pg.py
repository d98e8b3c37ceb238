import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'oracle'))
import numpy as np
import oracle as O
from advancedps_b200 import _abi, _lib, models
m = models.linear_gaussian()
N, T = 3000, 12
cfg = _abi.make_config(m, N, T, sampler=_abi.SAMPLER_PG, ess_threshold=0.5)
_, Y = O.simulate_data(m, T, 0xDA7A0004)
h = _lib.Handle(cfg); h.set_observations(Y)
ref=None
for seed in [1,2,3]:
    ro = O.sweep(cfg, Y, seed, ref_traj=ref, mode=O.CANON)
    le = h.sweep(seed, ref_traj=ref)
    logz, ess, res = h.step_stats()
    print("seed",seed,"res equal",np.array_equal(res,ro.resampled), res, ro.resampled)
    for t in range(1,T+1):
        xg=h.states(t); bad=np.nonzero((xg!=ro.x_hist[t-1]).any(axis=1))[0]
        if bad.size: print(" t",t,"x bad",bad.size,bad[:8]); 
        if t>=2:
            ag=h.ancestors(t); b2=np.nonzero(ag!=ro.anc_hist[t-1])[0]
            if b2.size: print(" t",t,"anc bad",b2.size,b2[:8], ag[b2[:8]], ro.anc_hist[t-1][b2[:8]])
        if bad.size: break
    slot_g, traj_g = h.pick_trajectory()
    slot_o, traj_o = O.pick_trajectory(cfg, seed, ro, mode=O.CANON)
    print(" slots", slot_g, slot_o, np.array_equal(traj_g,traj_o))
    ref = traj_o
