"""ncu target: k_propagate<4,4> at the C3 shape (LG d=4, N=4e6), plain launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from advancedps_b200 import _abi, _lib, models
T = 3
h = _lib.Handle(_abi.make_config(models.lg4(), 4_000_000, T, ess_threshold=0.5))
h.set_observations(np.random.default_rng(0).normal(size=(T, 4)) * 0.1 + 0.4)
h.sweep_profiled(1)
print(h.sweep_profiled(2))
