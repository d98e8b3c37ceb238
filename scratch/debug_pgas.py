import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O
from advancedps_b200 import _abi, _lib, models

m = models.linear_gaussian(r=0.0004)
N, T = 8192, 6
_, Y = O.simulate_data(m, T, 5)
cfg = _abi.make_config(m, N, T, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
h = _lib.Handle(cfg); h.set_observations(Y)
ref = None
for seed in (1, 2):
    ro = O.sweep(cfg, Y, seed, ref_traj=ref, mode=O.CANON)
    le = h.sweep(seed, ref_traj=ref)
    print("seed", seed, "launches", h.last_sweep_launches(), "le equal", le == ro.logevidence)
    for t in range(2, T + 2):
        a = h.ancestors(t)
        bad = np.nonzero(a != ro.anc_hist[t - 1])[0]
        print("  t", t, "ndiff", bad.size, "gpu ref anc", a[N - 1], "oracle", ro.anc_hist[t - 1][N - 1], "resampled", ro.resampled[t - 1])
        if bad.size and ref is not None and t >= 3:
            s = t - 1
            # recompute lw' on the host from the oracle history
            xprev = ro.x_hist[s - 2][:, 0][ro.anc_hist[s - 1]]       # x_{s-1}[anc_s[i]]
            xr = ref[s - 1, 0]
            mean = 0.5 * xprev + 0.2
            lp = -0.5 * ((xr - mean) / 0.1) ** 2
            print("     host lp stats: max at", np.argmax(lp), "gpu pick lp", lp[a[N - 1]], "oracle pick lp", lp[ro.anc_hist[t - 1][N - 1]])
    slot_o, traj_o = O.pick_trajectory(cfg, seed, ro, mode=O.CANON)
    ref = traj_o
