"""ctypes binding of libaps_b200.so (the C ABI declared in include/aps_b200.h).

There is no CPU fallback: if the CUDA library is missing or fails to load, importing a compute
entry point raises. ``build()`` compiles it in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("APS_LIB_PATH", os.path.join(_HERE, "libaps_b200.so"))
_SRC = [os.path.join(_HERE, "csrc", f) for f in ("aps_api.cu", "aps_kernels.cuh", "aps_device.cuh")]
_HDR = [os.path.join(_HERE, "..", "include", f) for f in ("aps_b200.h", "aps_model.h", "aps_math.h")]

# every symbol include/aps_b200.h declares
EXPORTS = [
    "aps_create", "aps_destroy", "aps_set_observations", "aps_sweep", "aps_sweep_profiled",
    "aps_pick_trajectory",
    "aps_host_alloc", "aps_host_free", "aps_get_weights", "aps_get_weights_view", "aps_get_logweights", "aps_get_final_states", "aps_get_trajectory",
    "aps_get_trajectories", "aps_pc_begin", "aps_pc_resample_propagate", "aps_pc_reweight", "aps_pc_logz",
    "aps_set_logweights",
    "aps_get_step_stats", "aps_get_states", "aps_get_ancestors", "aps_get_fat_counts", "aps_smoothing_mean", "aps_last_sweep_ms",
    "aps_last_sweep_launches", "aps_resample", "aps_logsumexp", "aps_softmax", "aps_ess",
    "aps_randcat", "aps_bench_resample", "aps_ipc_export", "aps_ipc_import", "aps_last_error",
    "aps_version",
]


class ApsError(RuntimeError):
    """Mirror of the reference's thrown ErrorException (src/resampling.jl:103,120,154,169;
    src/container.jl:292-298); ``code`` is the aps_status."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def build(force=False, verbose=False):
    """Compile libaps_b200.so in-tree (nvcc, sm_100a). Cross-compiles without a GPU."""
    stale = (not os.path.exists(SO_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(SO_PATH) for f in _SRC + _HDR)
    if force or stale:
        out = subprocess.run([os.path.join(_HERE, "csrc", "build.sh")], stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        if verbose or out.returncode:
            print(out.stdout)
        if out.returncode:
            raise RuntimeError("nvcc build of libaps_b200.so failed")
    return SO_PATH


_lib = None


def lib():
    """Load the CUDA library; raises (no fallback) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ApsError(_abi.ERR_CUDA, f"{SO_PATH} is missing: build it with __graft_entry__.build() "
                           "(advancedps.jl_b200/csrc/build.sh); there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        L.aps_last_error.restype = C.c_char_p
        L.aps_version.restype = C.c_char_p
        for name in EXPORTS:
            getattr(L, name)  # AttributeError if the ABI is incomplete
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise ApsError(rc, lib().aps_last_error().decode())


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _PinnedPool:
    """Page-locked host buffers for owned results (``aps_host_alloc``). A buffer belongs to the numpy
    array handed out until that array is garbage-collected, then it is reused: in steady state
    ``sample()`` returns an owned weights vector without a page-locking call or a staging copy."""

    def __init__(self, keep=4):
        self.free, self.keep = {}, keep

    def array(self, n, dtype=np.float64):
        import weakref

        nbytes = int(n) * np.dtype(dtype).itemsize
        lst = self.free.get(nbytes)
        if lst:
            p = lst.pop()
        else:
            p = self._alloc(nbytes)
            if nbytes not in self.free:
                # first buffer of this size: page-lock a spare one as well. A caller that keeps the previous
                # result alive while the next call runs (`w = sample(...).weights` in a loop) alternates between
                # two buffers; without the spare, the second one would be page-locked (milliseconds, and it
                # stalls the device) inside the caller's steady-state loop.
                self.free[nbytes] = [self._alloc(nbytes)]
        buf = (C.c_char * nbytes).from_address(p)
        a = np.frombuffer(buf, dtype=dtype, count=int(n))
        weakref.finalize(buf, self._give_back, nbytes, p)   # `a` keeps `buf` alive through its base
        return a

    @staticmethod
    def _alloc(nbytes):
        p = C.c_void_p()
        check(lib().aps_host_alloc(C.c_int64(nbytes), C.byref(p)))
        return p.value

    def _give_back(self, nbytes, p):
        lst = self.free.setdefault(nbytes, [])
        if len(lst) < self.keep:
            lst.append(p)
        else:
            try:
                lib().aps_host_free(C.c_void_p(p))
            except Exception:
                pass


_pinned = _PinnedPool()


class Handle:
    """Owns one ``aps_handle`` (device buffers, stream, CUDA graph) -- the device-side
    ParticleContainer (src/container.jl:5-12) in SoA form."""

    def __init__(self, cfg):
        self.cfg = cfg
        self._h = C.c_void_p()
        check(lib().aps_create(C.byref(cfg), C.byref(self._h)))
        # N: slots held by this handle (the local shard when world_size > 1); Ng: global particle count
        self.Ng, self.rank, self.world = cfg.n_particles, cfg.rank, cfg.world_size
        self.N, self.T, self.d, self.dy = cfg.n_particles // cfg.world_size, cfg.n_steps, cfg.model.d, cfg.model.dy
        # generation: bumped by every call that overwrites the particle store (sweep / pc_begin), so
        # lazily materialised results can tell that they went stale; ref_token: bumped whenever the
        # device-resident reference trajectory is replaced (a pick, or a sweep given a host trajectory)
        self.generation = 0
        self.ref_token = 0

    def close(self):
        if self._h:
            lib().aps_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU plumbing: opaque blobs the host exchanges between ranks
    def ipc_export(self):
        blob = np.zeros(_abi.IPC_BLOB_BYTES, dtype=np.uint8)
        check(lib().aps_ipc_export(self._h, ptr(blob)))
        return blob

    def ipc_import(self, blobs):
        blobs = np.ascontiguousarray(np.stack(blobs), dtype=np.uint8)
        assert blobs.shape == (self.world, _abi.IPC_BLOB_BYTES)
        check(lib().aps_ipc_import(self._h, ptr(blobs)))

    def set_observations(self, Y):
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(self.T, self.dy)
        check(lib().aps_set_observations(self._h, ptr(Y), C.c_int64(self.T), C.c_int64(self.dy)))

    def sweep(self, seed, ref_traj=None, ref_on_device=False):
        le = C.c_double()
        self.generation += 1
        if ref_on_device:
            ref = C.c_void_p(1)
        elif ref_traj is not None:
            self._ref_keep = np.ascontiguousarray(ref_traj, dtype=np.float64).reshape(self.T, self.d)
            ref = ptr(self._ref_keep)
            self.ref_token += 1
        else:
            ref = None
        check(lib().aps_sweep(self._h, C.c_uint64(seed), ref, C.byref(le)))
        return le.value

    def sweep_profiled(self, seed):
        """Unconditional sweep with per-launch CUDA events; returns (logevidence, ms[4], launches[4])
        for the {propagate, normalise, resample, pgas} kernel classes."""
        le = C.c_double()
        self.generation += 1
        ms = (C.c_float * 4)()
        nl = (C.c_int64 * 4)()
        check(lib().aps_sweep_profiled(self._h, C.c_uint64(seed), None, C.byref(le), ms, nl))
        return le.value, list(ms), list(nl)

    def pick_trajectory(self, want_traj=True):
        traj = np.zeros((self.T, self.d)) if want_traj else None
        slot = C.c_int64()
        self.ref_token += 1
        check(lib().aps_pick_trajectory(self._h, ptr(traj), C.byref(slot)))
        return slot.value, traj

    def weights(self, pinned=False):
        """Normalised weights of this handle's slots, an owned array. ``pinned=True`` takes the buffer
        from the page-locked pool (DMA-speed copy; the buffer returns to the pool when the array dies)."""
        w = _pinned.array(self.N) if pinned else np.empty(self.N)
        check(lib().aps_get_weights(self._h, ptr(w)))
        return w

    def weights_view(self):
        """Normalised weights as a read-only numpy view of the handle's pinned host buffer (no staging
        copy); valid until the next call on this handle."""
        p = C.POINTER(C.c_double)()
        check(lib().aps_get_weights_view(self._h, C.byref(p)))
        w = np.ctypeslib.as_array(p, shape=(self.N,))
        w.flags.writeable = False
        return w

    def logweights(self):
        w = np.zeros(self.N)
        check(lib().aps_get_logweights(self._h, ptr(w)))
        return w

    def final_states(self):
        x = np.zeros((self.N, self.d))
        check(lib().aps_get_final_states(self._h, ptr(x)))
        return x

    def trajectory(self, slot):
        traj = np.zeros((self.T, self.d))
        check(lib().aps_get_trajectory(self._h, C.c_int64(slot), ptr(traj)))
        return traj

    def trajectories(self):
        """All N trajectories of the final particle set, (T, N, d) -- collect(pc) of SMCSample (src/smc.jl:56)."""
        out = np.zeros((self.T, self.N, self.d))
        check(lib().aps_get_trajectories(self._h, ptr(out)))
        return out

    # ---- container level: the device ParticleContainer driven call by call (src/container.jl:171-363)
    def pc_begin(self, seed, ref_traj=None, ref_on_device=False):
        self.generation += 1
        if ref_on_device:
            ref = C.c_void_p(1)
        elif ref_traj is not None:
            self._ref_keep = np.ascontiguousarray(ref_traj, dtype=np.float64).reshape(self.T, self.d)
            ref = ptr(self._ref_keep)
            self.ref_token += 1
        else:
            ref = None
        check(lib().aps_pc_begin(self._h, C.c_uint64(seed), ref))

    def pc_resample_propagate(self):
        r = C.c_int32()
        check(lib().aps_pc_resample_propagate(self._h, C.byref(r)))
        return bool(r.value)

    def pc_reweight(self):
        d = C.c_int32()
        check(lib().aps_pc_reweight(self._h, C.byref(d)))
        return bool(d.value)

    def pc_logZ(self):
        z = C.c_double()
        check(lib().aps_pc_logz(self._h, C.byref(z)))
        return z.value

    def set_logweights(self, logw):
        logw = np.ascontiguousarray(logw, dtype=np.float64)
        assert logw.shape == (self.N,)
        check(lib().aps_set_logweights(self._h, ptr(logw)))

    def step_stats(self):
        logz = np.zeros(self.T)
        ess = np.zeros(self.T + 1)
        res = np.zeros(self.T + 1, dtype=np.uint8)
        check(lib().aps_get_step_stats(self._h, ptr(logz), ptr(ess), ptr(res)))
        return logz, ess, res

    def states(self, t):
        x = np.zeros((self.N, self.d))
        check(lib().aps_get_states(self._h, C.c_int64(t), ptr(x)))
        return x

    def ancestors(self, t):
        a = np.zeros(self.N, dtype=np.int32)
        check(lib().aps_get_ancestors(self._h, C.c_int64(t), ptr(a)))
        return a

    def smoothing_mean(self):
        """sum_i W_i X_i[t] for t = 1..T over the final weighted particle set (T x d); sharded
        handles return this rank's partial sums."""
        m = np.zeros((self.T, self.d))
        check(lib().aps_smoothing_mean(self._h, ptr(m)))
        return m

    def fat_counts(self):
        a = np.zeros(self.T + 1, dtype=np.int32)
        check(lib().aps_get_fat_counts(self._h, ptr(a)))
        return a

    def last_sweep_ms(self):
        ms = C.c_float()
        check(lib().aps_last_sweep_ms(self._h, C.byref(ms)))
        return ms.value

    def last_sweep_launches(self):
        n = C.c_int64()
        check(lib().aps_last_sweep_launches(self._h, C.byref(n)))
        return n.value


# ------------------------------------------------------------------ operator level
def _as_arg(a, dtype):
    """numpy array -> (keepalive, pointer); objects exposing data_ptr() (torch CUDA tensors) pass through."""
    if hasattr(a, "data_ptr"):
        return a, C.c_void_p(a.data_ptr()), int(a.numel())
    arr = np.ascontiguousarray(a, dtype=dtype)
    return arr, ptr(arr), arr.size


def resample(kind, w, n=None, key=0, ctr=0, out=None):
    """(kind, weights, n) -> 1-based int64 ancestor indices (the resampler callable of
    src/container.jl:182). ``w`` may be a numpy array (host) or a CUDA tensor (device)."""
    keep, wp, m = _as_arg(w, np.float64)
    n = m if n is None else int(n)
    if out is None:
        out = np.zeros(n, dtype=np.int64)
    keep_o, op, _ = _as_arg(out, np.int64)
    check(lib().aps_resample(int(kind), wp, C.c_int64(m), C.c_int64(n), C.c_uint64(key), C.c_uint64(ctr), op))
    return keep_o if keep_o is not out else out


def logsumexp(logw):
    keep, p, n = _as_arg(logw, np.float64)
    out = C.c_double()
    check(lib().aps_logsumexp(p, C.c_int64(n), C.byref(out)))
    return out.value


def ess(logw):
    keep, p, n = _as_arg(logw, np.float64)
    out = C.c_double()
    check(lib().aps_ess(p, C.c_int64(n), C.byref(out)))
    return out.value


def softmax(logw):
    keep, p, n = _as_arg(logw, np.float64)
    w = np.zeros(n)
    check(lib().aps_softmax(p, C.c_int64(n), ptr(w)))
    return w


def randcat(w, key=0, ctr=0):
    keep, p, n = _as_arg(w, np.float64)
    out = C.c_int64()
    check(lib().aps_randcat(p, C.c_int64(n), C.c_uint64(key), C.c_uint64(ctr), C.byref(out)))
    return out.value


def bench_resample(kind, n, iters=20, flush_l2=2, seed=1):
    """flush_l2: 0 none, 1 write 512 MB between launches (L2 left full of dirty lines),
    2 write 512 MB then stream-read 256 MB (L2 cold and clean)."""
    avg, mn = C.c_float(), C.c_float()
    check(lib().aps_bench_resample(int(kind), C.c_int64(n), int(iters), int(flush_l2), C.c_uint64(seed),
                                   C.byref(avg), C.byref(mn)))
    return avg.value, mn.value
