import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from advancedps_b200 import _abi, _lib, models

def run(tag, m, N, T, **kw):
    rng = np.random.default_rng(1)
    Y = rng.normal(size=(T, 1)) * 0.3 + 0.4
    out = []
    for nofused in (0, 1):
        if nofused: os.environ["APS_NO_FUSED"] = "1"
        else: os.environ.pop("APS_NO_FUSED", None)
        h = _lib.Handle(_abi.make_config(m, N, T, **kw)); h.set_observations(Y)
        ref = None
        h.sweep(1)
        if kw.get("sampler", 0) != 0:
            h.pick_trajectory()
        ms = []
        for k in range(6):
            h.sweep(10 + k, ref_on_device=kw.get("sampler", 0) != 0); ms.append(h.last_sweep_ms())
        out.append((h.last_sweep_launches(), min(ms)))
        h.close()
    print(f"{tag:28s} N={N:8d} T={T}: fused {out[0][1]:8.3f} ms ({out[0][0]} launches, {out[0][1]/T*1e3:6.1f} us/step) | three-kernel {out[1][1]:8.3f} ms ({out[1][0]} launches, {out[1][1]/T*1e3:6.1f} us/step)", flush=True)

lg, sv = models.linear_gaussian(), models.stochastic_volatility()
for N in (10_000, 100_000, 300_000, 1_000_000, 2_000_000, 4_000_000):
    run("LG1 SMC systematic", lg, N, 100)
run("LG1 SMC stratified", lg, 1_000_000, 100, resampler=_abi.RESAMPLE_STRATIFIED)
run("LG1 SMC systematic ESS 0.5", lg, 1_000_000, 100, ess_threshold=0.5)
for N in (100_000, 2_000_000):
    run("SV PGAS (C4 shape)", sv, N, 100, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
run("LG1 PG", lg, 1_000_000, 100, sampler=_abi.SAMPLER_PG, ess_threshold=0.5)
