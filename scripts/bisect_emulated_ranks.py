import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from advancedps_b200 import _abi, _lib, models
from test_gpu_sharded import assert_sharded_equal, run_sharded
world, N, T, thr = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
m = models.linear_gaussian()
_, Y = O.simulate_data(m, T, 0xDA7A0005)
try:
    hs, out = run_sharded(m, N, T, Y, [11, 12], world, _abi.RESAMPLE_SYSTEMATIC, thr)
    ro = O.sweep(_abi.make_config(m, N, T, ess_threshold=thr), Y, 12, mode=O.CANON)
    assert_sharded_equal(hs, out[1], ro, N, T)
    print("OK launches", hs[0].last_sweep_launches(), "resampled", ro.resampled.tolist())
except Exception as e:
    print("FAIL", type(e).__name__, str(e)[:120])
os._exit(0)
