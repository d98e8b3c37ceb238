import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from advancedps_b200 import _abi, _lib, models
kind = int(sys.argv[1])
T = 6
h = _lib.Handle(_abi.make_config(models.linear_gaussian(), 1_000_000, T, resampler=kind))
h.set_observations(bench.make_data()[:T])
h.sweep_profiled(1)
print(h.sweep_profiled(2))
