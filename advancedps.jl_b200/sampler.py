"""Host-side mirror of the reference's sampler surface over the C ABI.

Names, argument meaning and error behaviour follow TuringLang/AdvancedPS.jl v0.7.2
(paths relative to /root/reference):

  SMC / PG / PGAS, SMCSample / PGSample / PGState          src/smc.jl:1-27,59-99
  ResampleWithESSThreshold, resample_*, randcat            src/resampling.jl:11-204
  TracedSSM                                                src/model.jl:13-28
  sample(rng, model, SMC) / step(rng, model, PG|PGAS, st)  src/smc.jl:35-57,101-129
  ParticleContainer + getweights / logZ / ESS / reweight! / resample_propagate! / sweep!
                                                           src/container.jl:5-363

Julia's ``f!`` becomes ``f_`` here. The sweep for recognised state-space families runs entirely
on the device (``aps_sweep``); ``ParticleContainer`` keeps the reference's generic-particle
extension API (advance!/fork/forkr/update_ref!, ext/README.md:5-10) with only the weight
arithmetic and the ancestor draw offloaded through the operator-level entry points.
"""
import warnings

import numpy as np

from . import _abi, _lib
from ._lib import ApsError  # noqa: F401


# ------------------------------------------------------------------ rng plumbing
def _draw_key(rng):
    """One rand(rng, UInt64) from the user's generator (replaces seed_from_rng!, container.jl:143-159)."""
    if rng is None:
        rng = np.random.default_rng()
    if isinstance(rng, (int, np.integer)):
        return int(rng) & (2**64 - 1)
    return int(rng.integers(0, 2**64, dtype=np.uint64))


# ------------------------------------------------------------------ resamplers (src/resampling.jl)
def _resampler(kind, name):
    def f(rng, w, n=None):
        return _lib.resample(kind, w, n, key=_draw_key(rng), ctr=0)

    f.__name__ = name
    f.kind = kind
    f.__doc__ = f"{name}(rng, weights, n) -> n 1-based ancestor indices (src/resampling.jl)."
    return f


resample_multinomial = _resampler(_abi.RESAMPLE_MULTINOMIAL, "resample_multinomial")  # :31-35
resample_residual = _resampler(_abi.RESAMPLE_RESIDUAL, "resample_residual")           # :53-81
resample_stratified = _resampler(_abi.RESAMPLE_STRATIFIED, "resample_stratified")     # :98-131
resample_systematic = _resampler(_abi.RESAMPLE_SYSTEMATIC, "resample_systematic")     # :149-183
DEFAULT_RESAMPLER = resample_systematic                                                 # :185


def randcat(rng, p):
    """Single categorical draw, 1-based (src/resampling.jl:11-21)."""
    return _lib.randcat(p, key=_draw_key(rng), ctr=0)


class ResampleWithESSThreshold:
    """Resample with ``resampler`` if ESS <= threshold * N (src/resampling.jl:193-204)."""

    def __init__(self, resampler=None, threshold=None):
        # ResampleWithESSThreshold(), (resampler), (threshold::Real), (resampler, threshold)
        if threshold is None and isinstance(resampler, (int, float)) and not callable(resampler):
            resampler, threshold = None, resampler
        self.resampler = DEFAULT_RESAMPLER if resampler is None else resampler
        self.threshold = 0.5 if threshold is None else threshold

    def __eq__(self, other):
        return (isinstance(other, ResampleWithESSThreshold) and self.resampler is other.resampler
                and self.threshold == other.threshold)

    def __repr__(self):
        return f"ResampleWithESSThreshold({self.resampler.__name__}, {self.threshold})"


def _resampler_config(resampler):
    """resampler object -> (aps_resampler kind, ess_threshold or NaN for a bare function)."""
    if isinstance(resampler, ResampleWithESSThreshold):
        return resampler.resampler.kind, float(resampler.threshold)
    if hasattr(resampler, "kind"):
        return resampler.kind, float("nan")
    raise TypeError("the device sweep needs one of the resample_* functions or ResampleWithESSThreshold")


# ------------------------------------------------------------------ samplers (src/smc.jl)
class _ParticleSampler:
    def __init__(self, nparticles, *args):
        # (n), (n, resampler), (n, threshold::Real), (n, resampler, threshold)
        self.nparticles = int(nparticles)
        if len(args) == 0:
            self.resampler = self._default_resampler()
        elif len(args) == 1:
            a = args[0]
            if isinstance(a, (int, float)) and not callable(a):
                self.resampler = ResampleWithESSThreshold(DEFAULT_RESAMPLER, a)
            else:
                self.resampler = a
        elif len(args) == 2:
            self.resampler = ResampleWithESSThreshold(args[0], args[1])
        else:
            raise TypeError("too many arguments")

    @staticmethod
    def _default_resampler():
        return ResampleWithESSThreshold()


class SMC(_ParticleSampler):
    """SMC(n[, resampler = ResampleWithESSThreshold()]) / SMC(n, [resampler,] threshold)  (src/smc.jl:1-21)."""
    kind = _abi.SAMPLER_SMC


class PG(_ParticleSampler):
    """PG(n[, resampler]) / PG(n, [resampler,] threshold)  (src/smc.jl:59-81)."""
    kind = _abi.SAMPLER_PG


class PGAS(_ParticleSampler):
    """PGAS(n) = particle Gibbs with ancestor sampling, ESS threshold 1.0  (src/smc.jl:92-99)."""
    kind = _abi.SAMPLER_PGAS

    def __init__(self, nparticles, *args):
        super().__init__(nparticles, *args)

    @staticmethod
    def _default_resampler():
        return ResampleWithESSThreshold(1.0)


class TracedSSM:
    """TracedSSM(model, Y): a state-space model with its observations (src/model.jl:13-22).
    ``model`` is an ``aps_model`` built by ``advancedps_b200.models``; ``X`` is the trajectory."""

    def __init__(self, model, Y, X=None):
        self.model = model
        self.Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(-1, model.dy)
        self.X = X


class Trace:
    """Trace(model, rng) (src/model.jl:4-7); here only the carrier of ``.model.X``."""

    def __init__(self, model):
        self.model = model


class SMCSample:
    """SMCSample(trajectories, weights, logevidence) (src/smc.jl:23-27). ``weights`` and
    ``logevidence`` are owned values, like upstream's. ``trajectories`` is materialised through
    the device genealogy on access: one at a time (``smp.trajectories[i]``) or all at once
    (``smp.trajectories.materialize()``, T x N x d). The genealogy lives in the device store that
    the next ``sample`` / ``step`` on the same model overwrites; access after that raises instead
    of returning another sweep's particles (``sample(..., materialize=True)`` copies up front)."""

    def __init__(self, handle, tssm, weights, logevidence, materialize=False):
        self._h, self._tssm = handle, tssm
        self.weights = weights
        self.logevidence = logevidence
        self.trajectories = _LazyTrajectories(handle, tssm)
        if materialize:
            self.trajectories.materialize()

    def smoothed_mean(self):
        """Weighted mean trajectory sum_i W_i X_i (T x d), computed on the device from the genealogy."""
        self.trajectories._check_fresh()
        return self._h.smoothing_mean()


class _LazyTrajectories:
    def __init__(self, handle, tssm):
        self._h, self._tssm = handle, tssm
        self._gen = handle.generation
        self._all = None

    def _check_fresh(self):
        if self._h.generation != self._gen:
            raise ApsError(_abi.ERR_INVALID,
                           "this SMCSample's trajectories were not materialised before a later sample()/step() "
                           "reused the device particle store; call .trajectories.materialize() (or pass "
                           "materialize=True to sample) before the next call on the same model")

    def materialize(self):
        """Copy all N trajectories to the host (T x N x d) and serve every later access from there."""
        if self._all is None:
            self._check_fresh()
            self._all = self._h.trajectories()
        return self._all

    def __len__(self):
        return self._h.N

    def __getitem__(self, i):
        if not 0 <= i < self._h.N:
            raise IndexError(i)
        if self._all is not None:
            X = self._all[:, i, :].copy()
        else:
            self._check_fresh()
            X = self._h.trajectory(i)
        return Trace(TracedSSM(self._tssm.model, self._tssm.Y, X))

    def final_states(self):
        if self._all is not None:
            return self._all[-1].copy()
        self._check_fresh()
        return self._h.final_states()


class PGState:
    """PGState(trajectory) (src/smc.jl:83-85)."""

    def __init__(self, trajectory, handle=None):
        self.trajectory = trajectory
        # lets the next step condition on the device-resident copy -- as long as that copy is still
        # this state's trajectory (another chain on the same model may have replaced it)
        self._handle = handle
        self._ref_token = handle.ref_token if handle is not None else None

    def _on_device(self, h):
        return self._handle is h and self._ref_token == h.ref_token


class PGSample:
    """PGSample(trajectory, logevidence) (src/smc.jl:87-90)."""

    def __init__(self, trajectory, logevidence):
        self.trajectory = trajectory
        self.logevidence = logevidence


_handles = {}


def _handle_for(tssm, sampler, keep_history=True):
    """One device handle per (model, sampler shape); reused across step calls like the
    reference reuses nothing -- it rebuilds N traces per call (src/smc.jl:45-51,112-120)."""
    kind, thr = _resampler_config(sampler.resampler)
    key = (id(tssm), sampler.kind, sampler.nparticles, kind, repr(thr), keep_history)
    h = _handles.get(key)
    if h is None:
        cfg = _abi.make_config(tssm.model, sampler.nparticles, tssm.Y.shape[0], sampler=sampler.kind,
                               resampler=kind, ess_threshold=thr, keep_history=keep_history)
        h = _lib.Handle(cfg)
        h.set_observations(tssm.Y)
        h._tssm_keepalive = tssm
        _handles.clear()  # keep at most one live handle: particle stores can be tens of GB
        _handles[key] = h
    else:
        h.set_observations(tssm.Y)  # the caller may have changed Y in place; T x dy doubles
    return h


def sample(rng, model, sampler, n_iter=None, materialize=False, **kwargs):
    """AbstractMCMC.sample. ``SMC`` -> SMCSample (src/smc.jl:35-57); ``PG``/``PGAS`` with
    ``n_iter`` -> list of PGSample (AbstractMCMC's loop around ``step``)."""
    if isinstance(sampler, SMC):
        if kwargs:
            warnings.warn(f"keyword arguments {tuple(kwargs)} are not supported by `SMC`")  # smc.jl:41-43
        h = _handle_for(model, sampler)
        logev = h.sweep(_draw_key(rng))
        # weights: an owned copy, as upstream returns (src/smc.jl:56)
        return SMCSample(h, model, h.weights(pinned=True), logev, materialize=materialize)
    if n_iter is None:
        raise TypeError("sample(rng, model, PG|PGAS, n_iter): n_iter is required")
    out, state = [], None
    for _ in range(int(n_iter)):
        smp, state = step(rng, model, sampler, state)
        out.append(smp)
    return out


def step(rng, model, sampler, state=None, **kwargs):
    """AbstractMCMC.step for PG / PGAS: one conditional sweep + pick (src/smc.jl:101-129)."""
    if not isinstance(sampler, (PG, PGAS)):
        raise TypeError("step is defined for PG and PGAS")
    h = _handle_for(model, sampler)
    if state is None:
        logev = h.sweep(_draw_key(rng))
    elif state._on_device(h):
        logev = h.sweep(_draw_key(rng), ref_on_device=True)
    else:
        logev = h.sweep(_draw_key(rng), ref_traj=state.trajectory.model.X)
    _, traj = h.pick_trajectory()
    tr = Trace(TracedSSM(model.model, model.Y, traj))
    return PGSample(tr, logev), PGState(tr, h)


# ------------------------------------------------------------------ generic-particle container
class ParticleContainer:
    """ParticleContainer(vals[, logWs]) (src/container.jl:5-27) for arbitrary host particles.

    Particles implement the extension API of ext/README.md: ``advance(isref) -> score | None``,
    ``fork(isref)``. Weight arithmetic and ancestor draws run on the GPU through the
    operator-level ABI; the particles themselves stay on the host."""

    def __init__(self, vals, logWs=None, rng=None):
        self.vals = list(vals)
        self.logWs = np.zeros(len(self.vals)) if logWs is None else np.asarray(logWs, dtype=np.float64).copy()
        self.rng = rng if rng is not None else np.random.default_rng()

    def __len__(self):
        return len(self.vals)

    def __getitem__(self, i):
        return self.vals[i]

    def push_(self, p):  # Base.push!, :39-43
        self.vals.append(p)
        self.logWs = np.append(self.logWs, 0.0)
        return self


def reset_logweights_(pc):  # :75-78
    pc.logWs[:] = 0.0
    return pc


def increase_logweight_(pc, i, logw):  # :85-88 (0-based i)
    pc.logWs[i] += logw
    return pc


def getweights(pc):  # :95
    return _lib.softmax(pc.logWs)


def logZ(pc):  # :109
    return _lib.logsumexp(pc.logWs)


def getweight(pc, i):  # :102
    return float(np.exp(pc.logWs[i] - logZ(pc)))


def effectiveSampleSize(pc):  # :116-119
    return _lib.ess(pc.logWs)


def resample_propagate_(rng, pc, sampler, resampler=None, ref=None, weights=None):
    """resample_propagate! (src/container.jl:171-251): ESS decision, ancestor draw on the GPU,
    children grouped by parent in increasing parent order, reference kept in the last slot."""
    if resampler is None:
        resampler = DEFAULT_RESAMPLER
    n = len(pc)
    if isinstance(resampler, ResampleWithESSThreshold):
        if not effectiveSampleSize(pc) <= resampler.threshold * n:  # :242-247
            return pc
        resampler = resampler.resampler
    if weights is None:
        weights = getweights(pc)
    nres = n if ref is None else n - 1
    indx = resampler(pc.rng, weights, nres)  # uses the CONTAINER rng (:182)
    counts = np.bincount(np.asarray(indx) - 1, minlength=n)
    children = []
    for i in range(n):
        ni = int(counts[i])
        if ni > 0:
            p = pc.vals[i]
            isref = p is ref
            first = p.fork(isref) if isref else p
            children.append(first)
            for _ in range(1, ni):
                children.append(first.fork(isref))
    if ref is not None:
        if hasattr(ref, "update_ref"):
            ref.update_ref(pc, sampler)
        children.append(ref)
    pc.vals = children
    reset_logweights_(pc)
    return pc


def reweight_(pc, ref=None):
    """reweight! (src/container.jl:259-302): advance every particle, add its score."""
    n = len(pc)
    numdone = 0
    for i, p in enumerate(pc.vals):
        score = p.advance(p is ref)
        if score is None:
            numdone += 1
        else:
            increase_logweight_(pc, i, score)
    if numdone == n:
        return True
    if numdone != 0:
        raise ApsError(_abi.ERR_INVALID,
                       f"mis-aligned execution traces: # particles = {n} # completed trajectories = {numdone}."
                       " Please make sure the number of observations is NOT random.")
    return False


def sweep_(rng, pc, resampler, sampler, ref=None):
    """sweep! (src/container.jl:316-363) for generic host particles."""
    resample_propagate_(rng, pc, sampler, resampler, ref)
    logZ0 = logZ(pc)
    isdone = reweight_(pc, ref)
    logZ1 = logZ(pc)
    logevidence = logZ1 - logZ0
    while not isdone:
        resample_propagate_(rng, pc, sampler, resampler, ref)
        logZ0 = logZ(pc)
        isdone = reweight_(pc, ref)
        logZ1 = logZ(pc)
        logevidence += logZ1 - logZ0
    return logevidence


# ------------------------------------------------------------------ device-resident container
class DeviceParticleContainer:
    """The ParticleContainer of a recognised state-space family, resident on the GPU and driven call
    by call like the reference's own tests drive theirs (test/container.jl:28-119,
    test/pgas.jl:61-91): ``reweight_``, ``resample_propagate_``, ``logZ``, assignable ``logWs``.
    ``sweep_`` below is the loop of src/container.jl:316-363 over these calls; ``sample`` / ``step``
    run the same sweep fused on the device (``aps_sweep``)."""

    def __init__(self, model, sampler, rng=None, ref_traj=None):
        self.model, self.sampler = model, sampler
        self._h = _handle_for(model, sampler)
        self._h.pc_begin(_draw_key(rng), ref_traj=ref_traj)
        self.ref = ref_traj

    def __len__(self):
        return self._h.N

    @property
    def logWs(self):
        return self._h.logweights()

    @logWs.setter
    def logWs(self, v):  # pc.logWs = [...] (test/pgas.jl:82)
        self._h.set_logweights(v)

    def reweight_(self):
        return self._h.pc_reweight()

    def resample_propagate_(self):
        return self._h.pc_resample_propagate()

    def logZ(self):
        return self._h.pc_logZ()

    def getweights(self):
        return self._h.weights()

    def trajectory(self, i):
        """X of particle ``i`` (0-based) of the current set: ``pc.vals[i+1].model.X`` upstream."""
        return self._h.trajectory(i)

    def sweep_(self):
        """sweep! (src/container.jl:316-363) call by call; same result as ``aps_sweep``."""
        self.resample_propagate_()
        logZ0 = self.logZ()
        isdone = self.reweight_()
        logZ1 = self.logZ()
        logevidence = logZ1 - logZ0
        while not isdone:
            self.resample_propagate_()
            logZ0 = self.logZ()
            isdone = self.reweight_()
            logZ1 = self.logZ()
            logevidence += logZ1 - logZ0
        return logevidence
