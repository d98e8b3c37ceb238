import os, sys, time
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
h = S._handle_for(tssm, smc)
print("sample()            %.3f ms" % tm(lambda: S.sample(rng, tssm, smc).weights))
print("h.sweep             %.3f ms (device %.3f)" % (tm(lambda: h.sweep(5)), h.last_sweep_ms()))
print("weights pageable    %.3f ms" % tm(lambda: h.weights()))
print("weights pinned pool %.3f ms" % tm(lambda: h.weights(pinned=True)))
print("weights_view        %.3f ms" % tm(lambda: h.weights_view()))
print("set_observations    %.3f ms" % tm(lambda: h.set_observations(Y)))
print("_handle_for         %.3f ms" % tm(lambda: S._handle_for(tssm, smc)))
print("sweep + weights(pinned)         %.3f ms" % tm(lambda: (h.sweep(5), h.weights(pinned=True))[1]))
print("handle_for + sweep + weights    %.3f ms" % tm(lambda: (S._handle_for(tssm, smc).sweep(5), h.weights(pinned=True))[1]))
print("... + SMCSample                 %.3f ms" % tm(lambda: S.SMCSample(h, tssm, (h.sweep(S._draw_key(rng)), h.weights(pinned=True))[1], 0.0).weights))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10): w = S.sample(rng, tssm, smc).weights
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
