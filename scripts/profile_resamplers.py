"""ncu target: short C5-shape sweeps (N=1e6, plain launches) for each resampler, so the launch list
shows where the multinomial / residual / stratified steps spend their time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from advancedps_b200 import _abi, _lib, models
import bench
T = int(os.environ.get("APS_PROF_T", "4"))
for res in (0, 1, 2):
    h = _lib.Handle(_abi.make_config(models.linear_gaussian(), 1_000_000, T, resampler=res))
    h.set_observations(bench.make_data()[:T])
    h.sweep_profiled(1)
    print(res, h.sweep_profiled(2))
    h.close()
