"""C-ABI surface and host-mirror logic that need no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import advancedps_b200 as aps
from advancedps_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "aps_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(aps_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/aps_b200.h declares."""
    so = _lib.build()
    L = ctypes.CDLL(so)
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(L, name), f"{name} declared in aps_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == decl
    L.aps_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.aps_version()


def test_struct_layout_matches_header():
    """sizeof(aps_model) / sizeof(aps_config) computed from the header's field list."""
    D = _abi.APS_MAX_D
    assert ctypes.sizeof(_abi.ApsModel) == 16 + 8 * (D + D + D * D + D + D + D * D + D)
    assert ctypes.sizeof(_abi.ApsConfig) == ctypes.sizeof(_abi.ApsModel) + 8 + 8 + 4 + 4 + 8 + 4 * 4


def test_invalid_arguments_are_reported_without_a_gpu():
    L = _lib.lib()
    h = ctypes.c_void_p()
    cfg = _abi.make_config(aps.models.linear_gaussian(), 0, 10)
    assert L.aps_create(ctypes.byref(cfg), ctypes.byref(h)) == _abi.ERR_INVALID
    assert b"n_particles" in L.aps_last_error()
    cfg = _abi.make_config(aps.models.linear_gaussian(), 10, 10, sampler=_abi.SAMPLER_PG, keep_history=False)
    assert L.aps_create(ctypes.byref(cfg), ctypes.byref(h)) == _abi.ERR_INVALID
    out = (ctypes.c_int64 * 4)()
    w = (ctypes.c_double * 1)()
    assert L.aps_resample(_abi.RESAMPLE_SYSTEMATIC, w, ctypes.c_int64(0), ctypes.c_int64(4), ctypes.c_uint64(0),
                          ctypes.c_uint64(0), out) == _abi.ERR_INVALID
    assert b"weight vector is empty" in L.aps_last_error()  # src/resampling.jl:103,154


def test_sampler_constructors():
    """test/smc.jl:2-20,107-125 and test/pgas.jl:93-97."""
    s = aps.SMC(10)
    assert s.nparticles == 10 and s.resampler == aps.ResampleWithESSThreshold()
    s = aps.SMC(15, 0.6)
    assert s.nparticles == 15 and s.resampler == aps.ResampleWithESSThreshold(aps.resample_systematic, 0.6)
    s = aps.SMC(20, aps.resample_multinomial, 0.6)
    assert s.nparticles == 20 and s.resampler == aps.ResampleWithESSThreshold(aps.resample_multinomial, 0.6)
    s = aps.SMC(25, aps.resample_systematic)
    assert s.nparticles == 25 and s.resampler is aps.resample_systematic
    s = aps.PG(10)
    assert s.nparticles == 10 and s.resampler == aps.ResampleWithESSThreshold()
    s = aps.PG(60, 0.6)
    assert s.resampler == aps.ResampleWithESSThreshold(aps.resample_systematic, 0.6)
    s = aps.PG(80, aps.resample_multinomial, 0.6)
    assert s.resampler == aps.ResampleWithESSThreshold(aps.resample_multinomial, 0.6)
    s = aps.PG(100, aps.resample_systematic)
    assert s.resampler is aps.resample_systematic
    s = aps.PGAS(10)
    assert s.nparticles == 10 and s.resampler == aps.ResampleWithESSThreshold(1.0)
    assert aps.ResampleWithESSThreshold().threshold == 0.5
    assert aps.DEFAULT_RESAMPLER is aps.resample_systematic


def test_resampler_config_mapping():
    from advancedps_b200.sampler import _resampler_config
    k, thr = _resampler_config(aps.SMC(5, aps.resample_stratified).resampler)
    assert k == _abi.RESAMPLE_STRATIFIED and np.isnan(thr)
    k, thr = _resampler_config(aps.PGAS(5).resampler)
    assert k == _abi.RESAMPLE_SYSTEMATIC and thr == 1.0
    with pytest.raises(TypeError):
        _resampler_config(lambda rng, w, n: None)


def test_model_builders():
    m = aps.models.lg4()
    assert (m.d, m.dy) == (4, 4)
    A = np.array(m.A[:]).reshape(4, 4)
    assert np.allclose(np.linalg.eigvalsh(A), [0.4, 0.4, 0.4, 0.8])
    sv = aps.models.stochastic_volatility()
    assert sv.obs_kind == _abi.OBS_STOCH_VOL and sv.sigma0[0] == 0.5 and sv.A[0] == 0.9
    with pytest.raises(ValueError):
        _abi.make_model(_abi.OBS_LINEAR_GAUSS, 5, 1, np.zeros(5), np.ones(5), np.eye(5), np.zeros(5), np.ones(5))
