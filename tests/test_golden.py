"""Golden fixtures (tests/golden/sweeps.json, made by tests/golden/make_golden.py):
CPU: the oracle still reproduces them. GPU: the CUDA path reproduces them with no oracle involved."""
import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest

import oracle as O
from advancedps_b200 import _abi, _lib, models

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "sweeps.json")) as f:
    GOLD = json.load(f)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_cfg(g):
    fac, N, T, smp, res, thr, _, _, _ = g["spec"]
    m = getattr(models, fac)()
    cfg = _abi.make_config(m, N, T, sampler=getattr(_abi, "SAMPLER_" + smp), resampler=getattr(_abi, "RESAMPLE_" + res),
                           ess_threshold=float("nan") if thr is None else thr)
    Y = np.array(g["Y"]).reshape(T, m.dy)
    ref = None if g["ref"] is None else np.array(g["ref"]).reshape(T, m.d)
    return cfg, Y, ref


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_reproduces_golden(name):
    g = GOLD[name]
    cfg, Y, ref = build_cfg(g)
    r = O.sweep(cfg, Y, g["sweep_seed"], ref_traj=ref)
    assert r.logevidence == g["logevidence"]
    assert r.logz.tolist() == g["logz"] and r.ess.tolist() == g["ess"]
    assert r.resampled.tolist() == g["resampled"]
    assert digest(r.x_hist) == g["x_sha256"]
    assert digest(r.anc_hist[1:]) == g["anc_sha256"]
    slot, traj = O.pick_trajectory(cfg, g["sweep_seed"], r)
    assert slot == g["picked_slot"] and traj.ravel().tolist() == g["picked_traj"]


def test_golden_data_is_the_simulated_data():
    """the generating script is committed and deterministic"""
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    name = "lg1_c1_smc_systematic"
    assert mg.run_case(mg.CASES[name])[4] == GOLD[name]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_gpu_reproduces_golden(name):
    g = GOLD[name]
    cfg, Y, ref = build_cfg(g)
    N, T, d = cfg.n_particles, cfg.n_steps, cfg.model.d
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    le = h.sweep(g["sweep_seed"], ref_traj=ref)
    logz, ess, res = h.step_stats()
    assert le == g["logevidence"]
    assert logz.tolist() == g["logz"] and ess.tolist() == g["ess"] and res.tolist() == g["resampled"]
    x_hist = np.stack([h.states(t) for t in range(1, T + 1)])
    anc = np.stack([h.ancestors(t) for t in range(2, T + 2)])
    assert digest(x_hist) == g["x_sha256"]
    assert digest(anc) == g["anc_sha256"]
    assert digest(h.logweights()) == g["final_logw_sha256"]
    slot, traj = h.pick_trajectory()
    assert slot == g["picked_slot"] and traj.ravel().tolist() == g["picked_traj"]
