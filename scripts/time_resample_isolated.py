import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from advancedps_b200 import _abi, _lib
a, m = _lib.bench_resample(_abi.RESAMPLE_SYSTEMATIC, 1 << 25, iters=20, flush_l2=2)
print("isolated resample N=2^25: avg %.1f us min %.1f us -> %.0f GB/s" % (a * 1e3, m * 1e3, 12 * (1 << 25) / (a * 1e-3) / 1e9))
