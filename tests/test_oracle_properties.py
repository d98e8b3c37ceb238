"""Property tests of the oracle's resamplers (CPU only): the invariants every resampler of
src/resampling.jl has by construction, on random weight vectors with zeros, spikes and ragged sizes.
The GPU path is bit-compared with this oracle elsewhere, so these properties transfer."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle as O
from advancedps_b200 import _abi

KINDS = [_abi.RESAMPLE_MULTINOMIAL, _abi.RESAMPLE_RESIDUAL, _abi.RESAMPLE_STRATIFIED, _abi.RESAMPLE_SYSTEMATIC]


def weights(rng, m, style):
    if style == 0:
        w = rng.random(m)
    elif style == 1:                      # many exact zeros
        w = rng.random(m) * (rng.random(m) < 0.3)
        w[rng.integers(m)] += 1e-3
    elif style == 2:                      # one dominant spike
        w = rng.random(m) * 1e-9
        w[rng.integers(m)] = 1.0
    else:                                 # log-normal spread over many decades
        w = np.exp(6.0 * rng.normal(size=m))
    return w / w.sum()


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2**32 - 1), m=st.integers(1, 300), n=st.integers(1, 400), style=st.integers(0, 3),
       kind=st.sampled_from(KINDS), mode=st.sampled_from([O.CANON, O.SEQ]))
def test_resampler_invariants(seed, m, n, style, kind, mode):
    rng = np.random.default_rng(seed)
    w = weights(rng, m, style)
    idx = O.resample(kind, w, n, key=seed, step=3, mode=mode)
    assert idx.shape == (n,) and idx.min() >= 1 and idx.max() <= m
    counts = np.bincount(idx - 1, minlength=m)
    assert counts.sum() == n
    assert np.all(counts[w == 0.0] == 0)                       # zero weight, no offspring
    if kind in (_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED):
        assert np.all(np.diff(idx) >= 0)                       # sorted ascending (src/resampling.jl:98-183)
    if kind == _abi.RESAMPLE_SYSTEMATIC:
        assert np.all(np.abs(counts - n * w) < 1.0 + 1e-9 * n)  # offspring within one of n w_j
    if kind == _abi.RESAMPLE_STRATIFIED:
        assert np.all(np.abs(counts - n * w) < 2.0 + 1e-9 * n)
    if kind == _abi.RESAMPLE_RESIDUAL:
        assert np.all(counts >= np.floor(n * w * (1 - 1e-12)).astype(int))  # the deterministic copies (:62-71)


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 2**32 - 1), m=st.integers(1, 2000))
def test_seq_and_canon_agree_except_at_ties(seed, m):
    """SEQ keeps the reference's fp64 order, CANON is the exact-integer form the GPU reproduces: they may
    differ only where a cumulative weight lies within rounding distance of a threshold."""
    rng = np.random.default_rng(seed)
    w = weights(rng, m, 0)
    a = O.resample(_abi.RESAMPLE_SYSTEMATIC, w, m, key=seed, step=1, mode=O.SEQ)
    b = O.resample(_abi.RESAMPLE_SYSTEMATIC, w, m, key=seed, step=1, mode=O.CANON)
    assert np.count_nonzero(a != b) <= 1 and np.all(np.abs(a - b) <= 1)


@pytest.mark.parametrize("kind", KINDS)
def test_degenerate_inputs(kind):
    assert np.array_equal(O.resample(kind, np.array([1.0]), 5, key=1), np.ones(5, dtype=np.int64))
    one_hot = np.zeros(50)
    one_hot[17] = 1.0
    assert np.array_equal(O.resample(kind, one_hot, 33, key=2), np.full(33, 18))
    with pytest.raises(O.OracleError):
        O.resample(kind, np.zeros(4), 4, key=3)                # "sample could not be selected" (:120,169)
    with pytest.raises(O.OracleError):
        O.resample(kind, np.array([0.5, np.nan]), 2, key=3)


def test_canon_truncation_bias_is_bounded_by_n_over_2_pow_s():
    """q = floor(exp(logw - M) 2^S) truncates; every particle loses < 1 unit of 2^-S of the
    maximum, so Q (hence every logZ increment) is biased DOWN by at most N 2^-S relative:
    2^-22 at N = 2^20 (S = 42). Worst case: one dominant particle, all others just below one unit."""
    N = 1 << 20
    S = O.weight_shift(N)
    assert S == 62 - 20
    lw = np.full(N, np.log(0.999 * 2.0 ** -S))      # each truncates to 0
    lw[12345] = 0.0
    seq, canon = O.logsumexp(lw, O.SEQ), O.logsumexp(lw, O.CANON)
    assert canon == 0.0                              # only the dominant particle survives quantisation
    bound = np.log1p(N * 2.0 ** -S)
    assert 0.0 <= seq - canon <= bound and seq - canon > 0.9 * bound   # the bound is attained, not exceeded
    # typical heavy-tailed weights stay far inside it
    rng = np.random.default_rng(0)
    lw = -np.abs(rng.standard_cauchy(N)) * 3.0
    d = O.logsumexp(lw, O.SEQ) - O.logsumexp(lw, O.CANON)
    assert -1e-12 <= d <= bound
    # the ESS sums run on q >> h with h chosen so that sum (q >> h)^2 fits 63 bits: (63 - log2 N) / 2 = 21
    # bits per weight at N = 2^20, i.e. an ESS accurate to ~1e-6 relative here -- a threshold decision
    # (ess <= thr N) within that distance of the boundary can differ from the reference's fp64 value
    assert O.ess(lw, O.CANON) == pytest.approx(O.ess(lw, O.SEQ), rel=1e-5)
    assert abs(O.ess(lw, O.CANON) / O.ess(lw, O.SEQ) - 1) > 1e-8   # (and it is not accidentally exact)
