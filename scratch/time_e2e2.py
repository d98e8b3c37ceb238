import os, sys, time, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from advancedps_b200 import _abi, _lib, models, sampler as S
m = models.linear_gaussian(); Y = bench.make_data()
tssm = S.TracedSSM(m, Y); smc = S.SMC(1_000_000, S.resample_systematic)
rng = np.random.default_rng(1)
for _ in range(4): w = S.sample(rng, tssm, smc).weights
def tm(f, n=10):
    t0 = time.perf_counter()
    for _ in range(n): r = f()
    return (time.perf_counter() - t0) / n * 1e3
print("no sampler:        sample() %.3f ms" % tm(lambda: S.sample(rng, tssm, smc).weights))
for lms in (20, 50, 100):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", str(lms), "-i", "0"], stdout=subprocess.DEVNULL)
    time.sleep(0.5)
    print("nvidia-smi -lms %3d: sample() %.3f ms" % (lms, tm(lambda: S.sample(rng, tssm, smc).weights, 20)))
    p.terminate(); p.wait()
h2 = _lib.Handle(_abi.make_config(m, 1_000_000, 100)); h2.set_observations(Y); h2.sweep(1)
print("second handle alive: sample() %.3f ms" % tm(lambda: S.sample(rng, tssm, smc).weights))
