import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from advancedps_b200 import _abi, _lib, models
N = int(os.environ.get("N", 1000000)); T = 100
m = models.linear_gaussian(); Y = bench.make_data()
h = _lib.Handle(_abi.make_config(m, N, T)); h.set_observations(Y)
for w in range(3): h.sweep(1 + w)
ms = []
for k in range(10):
    h.sweep(10 + k); ms.append(h.last_sweep_ms())
_, cms, cn = h.sweep_profiled(5)
_, cms, cn = h.sweep_profiled(6)
print(os.environ.get("APS_LIB_PATH", "default"), "ms/sweep min %.3f med %.3f" % (min(ms), sorted(ms)[5]),
      "| per launch us: K1 %.2f K2 %.2f K3 %.2f" % tuple(1e3 * cms[i] / cn[i] for i in range(3)), flush=True)
