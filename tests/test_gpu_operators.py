"""GPU parity of the operator-level entry points (the resampler callable of src/container.jl:182
and the weight helpers of src/container.jl:95-119) against the oracle in CANON mode, bit-exact."""
import numpy as np
import pytest

import oracle as O
from advancedps_b200 import _abi, _lib

pytestmark = pytest.mark.gpu

KINDS = [_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED, _abi.RESAMPLE_MULTINOMIAL, _abi.RESAMPLE_RESIDUAL]


def weight_shapes(rng, m):
    yield "uniform", np.full(m, 1.0 / m)
    onehot = np.zeros(m)
    onehot[m // 2] = 1.0
    yield "one-hot", onehot
    yield "lognormal", rng.dirichlet(np.exp(rng.normal(size=m)))
    skew = rng.dirichlet(np.full(m, 0.05)) if m > 1 else np.ones(1)
    yield "skewed", skew
    yield "unnormalised", rng.uniform(0, 5, size=m)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("m", [1, 3, 1000, 2048, 2049, 100003])
def test_resample_matches_oracle(kind, m):
    rng = np.random.default_rng(m)
    for name, w in weight_shapes(rng, m):
        for n in sorted({m, max(1, m // 3), m + 17}):
            got = _lib.resample(kind, w, n, key=123 + n, ctr=5)
            ref = O.resample(kind, w, n, key=123 + n, step=5, mode=O.CANON)
            bad = np.nonzero(got != ref)[0]
            assert bad.size == 0, f"{name} m={m} n={n}: {bad.size} differ, first {bad[:5]} {got[bad[:5]]} {ref[bad[:5]]}"


@pytest.mark.parametrize("kind", KINDS)
def test_resampler_proportions_reference_test(kind):
    """test/resampling.jl:12-15 verbatim: D=[0.3,0.4,0.3], n=1e6."""
    D = np.array([0.3, 0.4, 0.3])
    n = 10**6
    idx = _lib.resample(kind, D, n, key=7, ctr=0)
    tol = 1e-3 if kind in (_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED) else 1e-2
    assert abs((idx == 2).sum() - 0.4 * n) <= tol * n
    assert np.array_equal(idx, O.resample(kind, D, n, key=7, step=0, mode=O.CANON))


def test_forced_middle_particle():
    """logWs = [-Inf, 0, -Inf] (test/pgas.jl:82): every draw is particle 2."""
    w = np.array([0.0, 1.0, 0.0])
    for kind in KINDS:
        assert np.all(_lib.resample(kind, w, 64, key=1, ctr=1) == 2)
    assert _lib.randcat(w, key=5, ctr=0) == 2


def test_large_resample_properties():
    """N = 2^24 + 5: sorted output, exact offspring totals, agreement with the oracle."""
    m = (1 << 24) + 5
    rng = np.random.default_rng(0)
    w = np.exp(-0.5 * rng.normal(size=m) ** 2)
    got = _lib.resample(_abi.RESAMPLE_SYSTEMATIC, w, m, key=99, ctr=2)
    assert got.min() >= 1 and got.max() <= m
    assert np.all(np.diff(got) >= 0)
    counts = np.bincount(got - 1, minlength=m)
    expect = m * w / w.sum()
    assert np.all(np.abs(counts - expect) < 1.0 + 1e-6)  # systematic: |offspring - N w| < 1
    ref = O.resample(_abi.RESAMPLE_SYSTEMATIC, w, m, key=99, step=2, mode=O.CANON)
    assert np.array_equal(got, ref)


def test_weight_helpers_match_oracle():
    rng = np.random.default_rng(4)
    for n in (1, 3, 1000, 4097, 250001):
        lw = rng.normal(size=n) * 3
        assert _lib.logsumexp(lw) == O.logsumexp(lw, O.CANON)
        assert _lib.ess(lw) == O.ess(lw, O.CANON)
        assert np.array_equal(_lib.softmax(lw), O.softmax(lw, O.CANON))
    lw = np.array([0.0, -1.0, -2.0])  # test/container.jl:52-58
    assert np.allclose(_lib.softmax(lw), np.exp(lw) / np.exp(lw).sum(), rtol=0, atol=1e-15)
    assert _lib.logsumexp(np.zeros(3)) == pytest.approx(np.log(3), abs=1e-15)
    assert _lib.ess(np.zeros(3)) == 3
    assert np.array_equal(_lib.softmax(np.zeros(3)), np.full(3, 1 / 3))
    assert _lib.logsumexp(np.full(5, -np.inf)) == -np.inf


def test_randcat_matches_oracle():
    rng = np.random.default_rng(5)
    for n in (1, 2, 10, 5000, 70001):
        w = rng.dirichlet(np.ones(n))
        for k in range(5):
            assert _lib.randcat(w, key=k, ctr=3) == O.randcat(w, key=k, step=3, mode=O.CANON)


def test_error_behaviour():
    with pytest.raises(_lib.ApsError) as e:  # src/resampling.jl:103,154
        _lib.resample(_abi.RESAMPLE_SYSTEMATIC, np.zeros(0), 4)
    assert e.value.code == _abi.ERR_INVALID and "empty" in str(e.value)
    for bad in (np.zeros(5), np.array([0.5, np.nan, 0.5]), np.array([0.5, -0.1, 0.6])):
        with pytest.raises(_lib.ApsError) as e:  # src/resampling.jl:120,169
            _lib.resample(_abi.RESAMPLE_SYSTEMATIC, bad, 5)
        assert e.value.code == _abi.ERR_WEIGHTS
    with pytest.raises(_lib.ApsError):
        _lib.ess(np.array([0.0, np.nan]))


def test_device_pointers_are_accepted():
    """weights and output living in HBM (torch CUDA tensors): no host round trip."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(6)
    w = rng.dirichlet(np.ones(30000))
    wd = torch.tensor(w, device="cuda")
    out = torch.zeros(30000, dtype=torch.int64, device="cuda")
    _lib.resample(_abi.RESAMPLE_SYSTEMATIC, wd, 30000, key=3, ctr=1, out=out)
    torch.cuda.synchronize()
    ref = O.resample(_abi.RESAMPLE_SYSTEMATIC, w, 30000, key=3, step=1, mode=O.CANON)
    assert np.array_equal(out.cpu().numpy(), ref)


def test_generic_particle_container():
    """test/container.jl:28-119 with host particles: weights/ESS/logZ on the GPU, resampling through
    the operator ABI, reference kept in the last slot."""
    import advancedps_b200 as aps

    class LogPParticle:  # LogPModel, test/container.jl:4-18
        def __init__(self, logp, steps=10):
            self.logp, self.c, self.steps = logp, 0, steps

        def advance(self, isref=False):
            if self.c >= self.steps:
                return None
            self.c += 1
            return self.logp

        def fork(self, isref=False):
            p = LogPParticle(self.logp, self.steps)
            p.c = self.c
            return p

    logps = [0.0, -1.0, -2.0]
    pc = aps.ParticleContainer([LogPParticle(l) for l in logps], rng=np.random.default_rng(0))
    assert np.array_equal(aps.getweights(pc), np.full(3, 1 / 3))
    assert aps.logZ(pc) == pytest.approx(np.log(3)) and aps.effectiveSampleSize(pc) == 3
    aps.reweight_(pc)
    assert np.array_equal(pc.logWs, logps)
    assert np.allclose(aps.getweights(pc), np.exp(logps) / np.exp(logps).sum(), atol=1e-15)
    aps.reweight_(pc)
    assert np.array_equal(pc.logWs, 2 * np.array(logps))
    assert aps.logZ(pc) == pytest.approx(np.log(np.exp(2 * np.array(logps)).sum()))
    ref = pc.vals[-1]
    aps.resample_propagate_(None, pc, aps.PG(3), aps.resample_systematic, ref)
    assert np.all(pc.logWs == 0) and pc.vals[-1] is ref and len(pc) == 3
    assert aps.effectiveSampleSize(pc) == 3
    aps.reweight_(pc)
    assert set(pc.logWs) <= set(logps)
    aps.increase_logweight_(pc, 1, 1.41)
    aps.reset_logweights_(pc)
    assert np.all(pc.logWs == 0)
    # evidence of a state-independent likelihood: -2 log 2 (test/smc.jl:104)
    pc2 = aps.ParticleContainer([LogPParticle(np.log(0.5), steps=2) for _ in range(100)], rng=np.random.default_rng(1))
    assert aps.sweep_(None, pc2, aps.ResampleWithESSThreshold(), aps.SMC(100)) == pytest.approx(-2 * np.log(2), abs=1e-14)
    # mis-aligned traces raise (src/container.jl:292-298, test/smc.jl:68)
    pc3 = aps.ParticleContainer([LogPParticle(0.0, steps=1), LogPParticle(0.0, steps=2)])
    aps.reweight_(pc3)
    with pytest.raises(aps.ApsError):
        aps.reweight_(pc3)
