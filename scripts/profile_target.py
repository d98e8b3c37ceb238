"""Short targets for ncu: one C2 sweep launched kernel by kernel, the fused persistent sweep kernel,
the isolated resample kernel at N = 2^25.   usage: profile_target.py {both|sweep|predraw|resample|fused|pgas}"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
if mode in ("fused", "pgas"):
    os.environ["APS_FUSED"] = "1"
from advancedps_b200 import _abi, _lib, models
import bench

T = int(os.environ.get("APS_PROF_T", "10"))
if mode in ("both", "sweep"):
    cfg = _abi.make_config(models.linear_gaussian(), 1_000_000, T)
    h = _lib.Handle(cfg)
    h.set_observations(bench.make_data()[:T])
    h.sweep_profiled(1)   # plain launches (no graph) so ncu sees every kernel
    le, ms, n = h.sweep_profiled(2)
    print("sweep", le, ms, n)
if mode == "predraw":     # the product path of C2, launched directly (APS_NO_GRAPH) so ncu sees k_draw_normals and k_propagate<..., PRE>
    os.environ["APS_NO_GRAPH"] = "1"
    cfg = _abi.make_config(models.linear_gaussian(), 1_000_000, T)
    h = _lib.Handle(cfg)
    h.set_observations(bench.make_data()[:T])
    print("predraw", h.sweep(1), h.sweep(2), h.last_sweep_launches(), h.last_sweep_ms())
if mode == "fused":       # the whole sweep as ONE cooperative launch (csrc/aps_fused.cuh)
    cfg = _abi.make_config(models.linear_gaussian(), 1_000_000, T)
    h = _lib.Handle(cfg)
    h.set_observations(bench.make_data()[:T])
    print("fused", h.sweep(1), h.sweep(2), h.last_sweep_launches(), h.last_sweep_ms())
if mode == "pgas":        # configs[3] shape, conditional sweep with ancestor sampling (fused by default)
    sv = models.stochastic_volatility()
    cfg = _abi.make_config(sv, 2_000_000, T, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
    h = _lib.Handle(cfg)
    h.set_observations(np.random.default_rng(0).normal(size=(T, 1)) * 0.3)
    h.sweep(1)
    h.pick_trajectory()
    print("pgas", h.sweep(2, ref_on_device=True), h.last_sweep_launches(), h.last_sweep_ms())
if mode == "resample_small":
    print("isolated-small", _lib.bench_resample(_abi.RESAMPLE_SYSTEMATIC, 1 << 20, iters=2, flush_l2=False))
if mode in ("both", "resample"):
    print("isolated", _lib.bench_resample(_abi.RESAMPLE_SYSTEMATIC, 1 << 25, iters=2, flush_l2=True))
