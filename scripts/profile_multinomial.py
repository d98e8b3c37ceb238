"""ncu target: one short multinomial sweep at N=1e6 (plain launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from advancedps_b200 import _abi, _lib, models
import bench
T = 3
h = _lib.Handle(_abi.make_config(models.linear_gaussian(), 1_000_000, T, resampler=int(os.environ.get("APS_PROF_RES", "0"))))
h.set_observations(bench.make_data()[:T])
h.sweep_profiled(1)
print(h.sweep_profiled(2))
