"""Emulated ranks on one GPU, long T: which library / settings time out?"""
import os, sys, time, threading
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("CONN", "32"))
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from advancedps_b200 import _abi, _lib, models
if os.environ.get("OLDLIB"):
    _lib.EXPORTS = ["aps_create", "aps_sweep", "aps_last_error", "aps_version"]
from test_gpu_sharded import make_ranks, collective
world = int(os.environ.get("WORLD", 4)); T = int(os.environ.get("T", 100)); res = int(os.environ.get("RES", 1))
m = models.linear_gaussian()
N = 8192 * world
Y = np.random.default_rng(0).normal(size=(T, 1)) * 0.2 + 0.4
hs = make_ranks(m, N, T, Y, world, res)
t0 = time.time()
try:
    les = collective(hs, lambda h: h.sweep(4321))
    print("OK", "world", world, "T", T, "res", res, "time %.2f" % (time.time() - t0), les[0], flush=True)
    t0 = time.time()
    les = collective(hs, lambda h: h.sweep(4322))
    print("OK second sweep time %.2f" % (time.time() - t0), flush=True)
except Exception as e:
    print("FAIL", "world", world, "T", T, "res", res, "time %.2f" % (time.time() - t0), str(e)[:80], flush=True)
