#!/usr/bin/env python
"""Regenerates tests/golden/sweeps.json from the CPU oracle (CANON mode).

The reference (Julia) cannot be imported in this image, so these fixtures pin the ORACLE's
canonical results (not Julia bit-streams): per-step logZ / ESS / decisions, the log-evidence and
SHA-256 digests of the full state and ancestor histories. tests check (a) the oracle still
reproduces them on CPU and (b) the CUDA path reproduces them on the GPU without the oracle.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
from advancedps_b200 import _abi, models  # noqa: E402

CASES = {
    # name: (model factory, N, T, sampler, resampler, ess_threshold, seed, data_key, conditional)
    "lg1_c1_smc_systematic": ("linear_gaussian", 1000, 50, "SMC", "SYSTEMATIC", None, 1234, 0xDA7A0001, False),
    "lg1_smc_ess_half": ("linear_gaussian", 3001, 20, "SMC", "SYSTEMATIC", 0.5, 7, 0xDA7A0001, False),
    "lg1_smc_stratified": ("linear_gaussian", 2500, 12, "SMC", "STRATIFIED", None, 8, 0xDA7A0001, False),
    "lg1_smc_multinomial": ("linear_gaussian", 2304, 10, "SMC", "MULTINOMIAL", None, 11, 0xDA7A0001, False),
    "lg1_smc_residual_ess": ("linear_gaussian", 2304, 10, "SMC", "RESIDUAL", 0.5, 12, 0xDA7A0001, False),
    "lg4_pg_conditional": ("lg4", 4000, 10, "PG", "SYSTEMATIC", 0.5, 9, 0xDA7A0003, True),
    "sv_pgas_conditional": ("stochastic_volatility", 2048, 16, "PGAS", "SYSTEMATIC", 1.0, 10, 0xDA7A0004, True),
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(spec):
    fac, N, T, smp, res, thr, seed, dkey, cond = spec
    m = getattr(models, fac)()
    cfg = _abi.make_config(m, N, T, sampler=getattr(_abi, "SAMPLER_" + smp),
                           resampler=getattr(_abi, "RESAMPLE_" + res),
                           ess_threshold=float("nan") if thr is None else thr)
    _, Y = O.simulate_data(m, T, dkey)
    ref = None
    if cond:
        r0 = O.sweep(cfg, Y, seed)
        _, ref = O.pick_trajectory(cfg, seed, r0)
        seed += 1
    r = O.sweep(cfg, Y, seed, ref_traj=ref)
    slot, traj = O.pick_trajectory(cfg, seed, r)
    return cfg, Y, ref, seed, {
        "spec": [fac, N, T, smp, res, thr, spec[6], dkey, cond],
        "Y": Y.ravel().tolist(),
        "ref": None if ref is None else ref.ravel().tolist(),
        "sweep_seed": seed,
        "logevidence": r.logevidence,
        "logz": r.logz.tolist(),
        "ess": r.ess.tolist(),
        "resampled": r.resampled.tolist(),
        "x_sha256": digest(r.x_hist),
        "anc_sha256": digest(r.anc_hist[1:]),
        "final_logw_sha256": digest(r.final_logw),
        "picked_slot": slot,
        "picked_traj": traj.ravel().tolist(),
    }


if __name__ == "__main__":
    out = {name: run_case(spec)[4] for name, spec in CASES.items()}
    with open(os.path.join(HERE, "sweeps.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "sweeps.json"), os.path.getsize(os.path.join(HERE, "sweeps.json")), "bytes")
