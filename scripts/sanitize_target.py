"""Small sweeps for compute-sanitizer (memcheck / racecheck): the fused persistent kernel (systematic,
stratified, PGAS conditional), the three-kernel path, the stepwise container, multinomial / residual."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O
from advancedps_b200 import _abi, _lib, models

def run(tag, m, N, T, fused, **kw):
    if fused: os.environ["APS_FUSED"] = "1"; os.environ.pop("APS_NO_FUSED", None)
    else: os.environ["APS_NO_FUSED"] = "1"; os.environ.pop("APS_FUSED", None)
    _, Y = O.simulate_data(m, T, 7)
    cfg = _abi.make_config(m, N, T, **kw)
    h = _lib.Handle(cfg); h.set_observations(Y)
    ref = None
    ok = True
    for seed in (1, 2):
        ro = O.sweep(cfg, Y, seed, ref_traj=ref, mode=O.CANON)
        le = h.sweep(seed, ref_traj=ref)
        ok &= le == ro.logevidence and np.array_equal(h.ancestors(T + 1), ro.anc_hist[T])
        if kw.get("sampler", 0):
            _, ref = h.pick_trajectory()
    print(tag, "fused" if fused else "three-kernel", "launches", h.last_sweep_launches(), "OK" if ok else "MISMATCH", flush=True)

lg, sv = models.linear_gaussian(), models.stochastic_volatility()
for fused in (True, False):
    run("lg1 sys N=20011", lg, 20011, 4, fused)
    run("lg1 strat ess", lg, 9000, 4, fused, resampler=_abi.RESAMPLE_STRATIFIED, ess_threshold=0.6)
    run("sv pgas", sv, 12000, 5, fused, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
run("lg4 pg", models.lg4(), 5000, 4, False, sampler=_abi.SAMPLER_PG, ess_threshold=0.5)
run("lg1 multinomial", lg, 9000, 3, False, resampler=_abi.RESAMPLE_MULTINOMIAL)
run("lg1 residual", lg, 9000, 3, False, resampler=_abi.RESAMPLE_RESIDUAL)
