import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from advancedps_b200 import _abi, _lib, models
m = models.lg4(); N, T = 4_000_000, 200
h = _lib.Handle(_abi.make_config(m, N, T, sampler=_abi.SAMPLER_PG, ess_threshold=0.5))
h.set_observations(np.random.default_rng(0).normal(size=(T, 4)) * 0.3)
h.sweep(1); h.pick_trajectory()
ms = []
for k in range(3):
    h.sweep(2 + k, ref_on_device=True); ms.append(h.last_sweep_ms()); h.pick_trajectory()
_, cms, cn = h.sweep_profiled(9)
print("pairs" if os.environ.get("APS_K1_PAIRS") else "per-slot", "C3 ms/sweep min %.2f" % min(ms), "| per launch us: K1 %.1f K2 %.1f K3 %.1f" % tuple(1e3 * cms[i] / cn[i] for i in range(3)), flush=True)
