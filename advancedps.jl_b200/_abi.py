"""ctypes mirror of include/aps_b200.h and include/aps_model.h (struct layouts and enums only).

Shared by the product binding (``_lib.py``) and by the test-side oracle binding
(``oracle/oracle.py``); contains no compute.
"""
import ctypes as C

import numpy as np

APS_MAX_D = 4

# enum aps_obs_kind
OBS_LINEAR_GAUSS, OBS_STOCH_VOL, OBS_CONST = 0, 1, 2
# enum aps_resampler (src/resampling.jl)
RESAMPLE_MULTINOMIAL, RESAMPLE_RESIDUAL, RESAMPLE_STRATIFIED, RESAMPLE_SYSTEMATIC = 0, 1, 2, 3
# enum aps_sampler (src/smc.jl)
SAMPLER_SMC, SAMPLER_PG, SAMPLER_PGAS = 0, 1, 2
# enum aps_status
OK, ERR_INVALID, ERR_WEIGHTS, ERR_CUDA, ERR_COMM, ERR_NOMEM = 0, 1, 2, 3, 4, 5

IPC_BLOB_BYTES = 512

_D = C.c_double


class ApsModel(C.Structure):
    _fields_ = [
        ("obs_kind", C.c_int32),
        ("d", C.c_int32),
        ("dy", C.c_int32),
        ("reserved", C.c_int32),
        ("mu0", _D * APS_MAX_D),
        ("sigma0", _D * APS_MAX_D),
        ("A", _D * (APS_MAX_D * APS_MAX_D)),
        ("b", _D * APS_MAX_D),
        ("q", _D * APS_MAX_D),
        ("H", _D * (APS_MAX_D * APS_MAX_D)),
        ("r", _D * APS_MAX_D),
    ]


class ApsConfig(C.Structure):
    _fields_ = [
        ("model", ApsModel),
        ("n_particles", C.c_int64),
        ("n_steps", C.c_int64),
        ("sampler", C.c_int32),
        ("resampler", C.c_int32),
        ("ess_threshold", C.c_double),
        ("keep_history", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
    ]


def _vec(x, n, fill=0.0):
    a = np.full(n, fill, dtype=np.float64)
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    a[: x.size] = x
    return a


def make_model(obs_kind, d, dy, mu0, sigma0, A, b, q, H=None, r=None):
    """Fill an ``aps_model``. Matrices are given as (d, d) / (dy, d); padded to APS_MAX_D."""
    if not (1 <= d <= APS_MAX_D and 1 <= dy <= APS_MAX_D):
        raise ValueError(f"state / observation dimension must be in 1..{APS_MAX_D}")
    m = ApsModel()
    m.obs_kind, m.d, m.dy, m.reserved = obs_kind, d, dy, 0
    m.mu0[:] = _vec(mu0, APS_MAX_D)
    m.sigma0[:] = _vec(sigma0, APS_MAX_D)
    Ap = np.zeros((APS_MAX_D, APS_MAX_D))
    Ap[:d, :d] = np.asarray(A, dtype=np.float64).reshape(d, d)
    m.A[:] = Ap.ravel()
    m.b[:] = _vec(b, APS_MAX_D)
    m.q[:] = _vec(q, APS_MAX_D, 1.0)
    Hp = np.zeros((APS_MAX_D, APS_MAX_D))
    if H is not None:
        Hp[:dy, :d] = np.asarray(H, dtype=np.float64).reshape(dy, d)
    m.H[:] = Hp.ravel()
    m.r[:] = _vec(r if r is not None else 1.0, APS_MAX_D, 1.0)
    return m


def make_config(model, n_particles, n_steps, sampler=SAMPLER_SMC, resampler=RESAMPLE_SYSTEMATIC,
                ess_threshold=float("nan"), keep_history=True, device=0, rank=0, world_size=1):
    c = ApsConfig()
    c.model = model
    c.n_particles, c.n_steps = int(n_particles), int(n_steps)
    c.sampler, c.resampler = int(sampler), int(resampler)
    c.ess_threshold = float(ess_threshold)
    c.keep_history = 1 if keep_history else 0
    c.device, c.rank, c.world_size = int(device), int(rank), int(world_size)
    return c
