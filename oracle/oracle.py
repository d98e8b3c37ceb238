"""ctypes binding of the CPU oracle (oracle/libaps_oracle.so). TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the product package. See the header of oracle/aps_oracle.cpp for what is
restated from the reference and what is (un)pinned.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import advancedps_b200 as _pkg  # noqa: E402  (struct layouts only)

_abi = _pkg._abi
SEQ, CANON = 0, 1

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libaps_oracle.so")
    src = os.path.join(_HERE, "aps_oracle.cpp")
    hdrs = [os.path.join(_HERE, "..", "include", h) for h in ("aps_math.h", "aps_model.h", "aps_b200.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaps_oracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_exp.restype = C.c_double
        L.orc_exp.argtypes = [C.c_double]
        L.orc_log.restype = C.c_double
        L.orc_log.argtypes = [C.c_double]
        L.orc_u01.restype = C.c_double
        L.orc_u01.argtypes = [C.c_uint64]
        L.orc_randcat_seq.restype = C.c_int64
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ------------------------------------------------------------------ math
def set_threads(n):
    """OpenMP threads for the particle-parallel loops of the CANON mode (bit-identical results for
    any n); 1 = the serial sweep of the reference. Returns the previous setting's maximum."""
    lib().orc_set_threads(int(n))


def max_threads():
    return int(lib().orc_max_threads())


def philox2x64(c0, c1, key):
    out = np.zeros(2, dtype=np.uint64)
    lib().orc_philox2x64(C.c_uint64(c0), C.c_uint64(c1), C.c_uint64(key), _ptr(out))
    return int(out[0]), int(out[1])


def exp(x):
    return lib().orc_exp(float(x))


def log(x):
    return lib().orc_log(float(x))


def sincospi(t):
    s, c = C.c_double(), C.c_double()
    lib().orc_sincospi(C.c_double(t), C.byref(s), C.byref(c))
    return s.value, c.value


def normal_pair(w0, w1):
    z = np.zeros(2)
    lib().orc_normal_pair(C.c_uint64(w0), C.c_uint64(w1), _ptr(z))
    return z[0], z[1]


def weight_shift(n):
    return lib().orc_weight_shift(C.c_int64(n))


# ------------------------------------------------------------------ weights
class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle status {code}")
        self.code = code


def _chk(rc):
    if rc != 0:
        raise OracleError(rc)


def logsumexp(logw, mode=CANON):
    logw = np.ascontiguousarray(logw, dtype=np.float64)
    out = C.c_double()
    _chk(lib().orc_logsumexp(_ptr(logw), C.c_int64(logw.size), mode, C.byref(out)))
    return out.value


def softmax(logw, mode=CANON):
    logw = np.ascontiguousarray(logw, dtype=np.float64)
    w = np.empty_like(logw)
    _chk(lib().orc_softmax(_ptr(logw), C.c_int64(logw.size), mode, _ptr(w)))
    return w


def ess(logw, mode=CANON):
    logw = np.ascontiguousarray(logw, dtype=np.float64)
    out = C.c_double()
    _chk(lib().orc_ess(_ptr(logw), C.c_int64(logw.size), mode, C.byref(out)))
    return out.value


def quantise_logw(logw):
    logw = np.ascontiguousarray(logw, dtype=np.float64)
    q = np.zeros(logw.size, dtype=np.uint64)
    m, Q = C.c_double(), C.c_uint64()
    _chk(lib().orc_quantise_logw(_ptr(logw), C.c_int64(logw.size), _ptr(q), C.byref(m), C.byref(Q)))
    return q, m.value, Q.value


def quantise_shard(logw, global_max, n_global):
    logw = np.ascontiguousarray(logw, dtype=np.float64)
    q = np.zeros(logw.size, dtype=np.uint64)
    Q = C.c_uint64()
    _chk(lib().orc_quantise_shard(_ptr(logw), C.c_int64(logw.size), C.c_double(global_max), C.c_int64(n_global),
                                  _ptr(q), C.byref(Q)))
    return q, Q.value


def quantise_w(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    q = np.zeros(w.size, dtype=np.uint64)
    Q = C.c_uint64()
    _chk(lib().orc_quantise_w(_ptr(w), C.c_int64(w.size), _ptr(q), C.byref(Q)))
    return q, Q.value


# ------------------------------------------------------------------ resamplers
def resample(kind, w, n=None, key=0, step=0, mode=CANON):
    """(kind, fp64 weights, n) -> 1-based int64 indices; uniforms from Philox(key, step)."""
    w = np.ascontiguousarray(w, dtype=np.float64)
    n = w.size if n is None else int(n)
    out = np.zeros(max(n, 1), dtype=np.int64)
    _chk(lib().orc_resample(int(kind), mode, _ptr(w), C.c_int64(w.size), C.c_int64(n),
                            C.c_uint64(key), C.c_uint64(step), _ptr(out)))
    return out[:n]


def resample_systematic_seq(w, n, u0):
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.zeros(max(n, 1), dtype=np.int64)
    _chk(lib().orc_resample_systematic_seq(_ptr(w), C.c_int64(w.size), C.c_int64(n), C.c_double(u0), _ptr(out)))
    return out[:n]


def resample_stratified_seq(w, n, us):
    w = np.ascontiguousarray(w, dtype=np.float64)
    us = np.ascontiguousarray(us, dtype=np.float64)
    out = np.zeros(max(n, 1), dtype=np.int64)
    _chk(lib().orc_resample_stratified_seq(_ptr(w), C.c_int64(w.size), C.c_int64(n), _ptr(us), _ptr(out)))
    return out[:n]


def resample_multinomial_seq(w, n, us):
    w = np.ascontiguousarray(w, dtype=np.float64)
    us = np.ascontiguousarray(us, dtype=np.float64)
    out = np.zeros(max(n, 1), dtype=np.int64)
    _chk(lib().orc_resample_multinomial_seq(_ptr(w), C.c_int64(w.size), C.c_int64(n), _ptr(us), _ptr(out)))
    return out[:n]


def resample_residual_seq(w, n, us):
    w = np.ascontiguousarray(w, dtype=np.float64)
    us = np.ascontiguousarray(us, dtype=np.float64)
    out = np.zeros(max(n, 1), dtype=np.int64)
    _chk(lib().orc_resample_residual_seq(_ptr(w), C.c_int64(w.size), C.c_int64(n), _ptr(us), _ptr(out)))
    return out[:n]


def randcat_seq(p, r):
    p = np.ascontiguousarray(p, dtype=np.float64)
    return int(lib().orc_randcat_seq(_ptr(p), C.c_int64(p.size), C.c_double(r)))


def randcat(w, key=0, step=0, mode=CANON):
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = C.c_int64()
    _chk(lib().orc_randcat(_ptr(w), C.c_int64(w.size), mode, C.c_uint64(key), C.c_uint64(step), C.byref(out)))
    return out.value


# ------------------------------------------------------------------ sweep
class SweepResult:
    pass


def sweep(cfg, Y, seed, ref_traj=None, mode=CANON, history=True):
    """Run one (conditional) sweep. Returns an object with logevidence, x_hist (T, N, d),
    anc_hist (T+1, N), logz (T,), ess (T+1,), resampled (T+1,), final_logw, final_w."""
    N, T, d = cfg.n_particles, cfg.n_steps, cfg.model.d
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    assert Y.size == T * cfg.model.dy
    r = SweepResult()
    r.x_hist = np.zeros((T, N, d)) if history else None
    r.anc_hist = np.zeros((T + 1, N), dtype=np.int32) if history else None
    r.logz = np.zeros(T)
    r.ess = np.zeros(T + 1)
    r.resampled = np.zeros(T + 1, dtype=np.uint8)
    r.final_logw = np.zeros(N)
    r.final_w = np.zeros(N)
    ref = None
    if ref_traj is not None:
        ref = np.ascontiguousarray(ref_traj, dtype=np.float64)
        assert ref.size == T * d
    le = C.c_double()
    _chk(lib().orc_sweep(C.byref(cfg), _ptr(Y), C.c_uint64(seed), _ptr(ref), mode, C.byref(le),
                         _ptr(r.x_hist), _ptr(r.anc_hist), _ptr(r.logz), _ptr(r.ess), _ptr(r.resampled),
                         _ptr(r.final_logw), _ptr(r.final_w)))
    r.logevidence = le.value
    return r


def pick_trajectory(cfg, seed, res, mode=CANON):
    T, d = cfg.n_steps, cfg.model.d
    traj = np.zeros((T, d))
    slot = C.c_int64()
    _chk(lib().orc_pick_trajectory(C.byref(cfg), C.c_uint64(seed), mode, _ptr(res.final_logw), _ptr(res.x_hist),
                                   _ptr(res.anc_hist), C.byref(slot), _ptr(traj)))
    return slot.value, traj


def trajectory(cfg, slot, res):
    traj = np.zeros((cfg.n_steps, cfg.model.d))
    _chk(lib().orc_trajectory(C.byref(cfg), C.c_int64(slot), _ptr(res.x_hist), _ptr(res.anc_hist), _ptr(traj)))
    return traj


def simulate_data(model, T, data_key):
    x = np.zeros((T, model.d))
    y = np.zeros((T, model.dy))
    _chk(lib().orc_simulate_data(C.byref(model), C.c_int64(T), C.c_uint64(data_key), _ptr(x), _ptr(y)))
    return x, y
