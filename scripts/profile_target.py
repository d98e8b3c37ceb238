"""Short target for ncu: one C2 sweep (graph replay) + the isolated resample kernel at N=2^25."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from advancedps_b200 import _abi, _lib, models
sys.path.insert(0, ROOT)
import bench

mode = sys.argv[1] if len(sys.argv) > 1 else "both"
if mode in ("both", "sweep"):
    T = int(os.environ.get("APS_PROF_T", "10"))
    cfg = _abi.make_config(models.linear_gaussian(), 1_000_000, T)
    h = _lib.Handle(cfg)
    h.set_observations(bench.make_data()[:T])
    h.sweep_profiled(1)   # plain launches (no graph) so ncu sees every kernel
    le, ms, n = h.sweep_profiled(2)
    print("sweep", le, ms, n)
if mode == "resample_small":
    print("isolated-small", _lib.bench_resample(_abi.RESAMPLE_SYSTEMATIC, 1 << 20, iters=2, flush_l2=False))
if mode in ("both", "resample"):
    print("isolated", _lib.bench_resample(_abi.RESAMPLE_SYSTEMATIC, 1 << 25, iters=2, flush_l2=True))
