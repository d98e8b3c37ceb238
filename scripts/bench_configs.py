"""Times every BASELINE.json config shape once (device time of aps_sweep, CUDA events in the library)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from advancedps_b200 import _abi, _lib, models
import bench

rng = np.random.default_rng(0)
out = []
def run(name, m, N, T, smp, res, thr, iters=3, cond=False):
    Y = bench.make_data()[:T] if m.d == 1 and m.obs_kind == 0 and T <= 100 else rng.normal(size=(T, m.dy)) * 0.3
    h = _lib.Handle(_abi.make_config(m, N, T, sampler=smp, resampler=res, ess_threshold=thr))
    h.set_observations(Y)
    h.sweep(1)
    if cond:
        h.pick_trajectory()
    ms = []
    for k in range(iters):
        le = h.sweep(2 + k, ref_on_device=cond)
        ms.append(h.last_sweep_ms())
        if cond:
            h.pick_trajectory()
    r = {"config": name, "N": N, "T": T, "ms_per_sweep": min(ms), "particle_steps_per_s": N * T / (min(ms) * 1e-3),
         "launches": h.last_sweep_launches(), "logevidence": le}
    print(json.dumps(r), flush=True)
    out.append(r)
    h.close()

nan = float("nan")
run("C1 LG1 T=50 N=1e3 SMC systematic", models.linear_gaussian(), 1000, 50, 0, 3, nan)
run("C2 LG1 T=100 N=1e6 SMC systematic (bare)", models.linear_gaussian(), 10**6, 100, 0, 3, nan)
run("C2' LG1 T=100 N=1e6 SMC(N) ESS 0.5", models.linear_gaussian(), 10**6, 100, 0, 3, 0.5)
for kind, nm in ((2, "stratified"), (1, "residual"), (0, "multinomial")):
    run(f"C5-shape LG1 T=100 N=1e6 SMC {nm} (1 GPU shard size)", models.linear_gaussian(), 10**6, 100, 0, kind, nan)
run("C3 LG4 T=200 N=4e6 PG conditional", models.lg4(), 4 * 10**6, 200, 1, 3, 0.5, iters=2, cond=True)
run("C4 SV T=500 N=2e6 PGAS conditional", models.stochastic_volatility(), 2 * 10**6, 500, 2, 3, 1.0, iters=2, cond=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_configs_r01.json"), "w"), indent=1)
