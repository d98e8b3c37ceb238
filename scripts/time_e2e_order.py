"""Is the e2e figure of bench.py (sample() wall time) sensitive to warm-up order and to the nvidia-smi clock sampler?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from advancedps_b200 import models, sampler as S
m = models.linear_gaussian(); Y = bench.make_data()
tssm = S.TracedSSM(m, Y); smc = S.SMC(1_000_000, S.resample_systematic)
rng = np.random.default_rng(1)
def tm(f, n=20):
    t0 = time.perf_counter()
    for _ in range(n): r = f()
    return (time.perf_counter() - t0) / n * 1e3
f = lambda: S.sample(rng, tssm, smc).weights
t0 = time.perf_counter(); f(); print("first call (handle creation) %.1f ms" % ((time.perf_counter() - t0) * 1e3))
h = S._handle_for(tssm, smc)
per = []
for _ in range(40):
    t0 = time.perf_counter(); f(); per.append(((time.perf_counter() - t0) * 1e3, h.last_sweep_ms()))
print("calls 2..41 wall/device ms:", " ".join("%.2f/%.2f" % x for x in per), flush=True)
print("sample() x20, five blocks:", " ".join("%.3f" % tm(f) for _ in range(5)), flush=True)
print("sweep+weights(pinned):    ", " ".join("%.3f" % tm(lambda: (h.sweep(S._draw_key(rng)), h.weights(pinned=True))[1]) for _ in range(3)), flush=True)
print("sample() again:           ", " ".join("%.3f" % tm(f) for _ in range(3)), flush=True)
for lms in (20, 100, 500):
    cs = bench.ClockSampler(0)
    cs.Q = cs.Q  # same query
    import subprocess, threading
    cs.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={cs.Q}", "--format=csv,noheader,nounits", "-lms", str(lms), "-i", "0"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    cs.th = threading.Thread(target=cs._read, daemon=True); cs.th.start()
    time.sleep(0.3)
    print("with nvidia-smi -lms %d:   " % lms, " ".join("%.3f" % tm(f) for _ in range(3)), "| device-timed sweep %.3f" % h.last_sweep_ms(), flush=True)
    cs.stop()
print("sample() after sampler:   ", " ".join("%.3f" % tm(f) for _ in range(3)), flush=True)
