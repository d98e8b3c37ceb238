"""The oracle pinned on every RNG-independent known answer the reference's own tests hold for the
hot path, in both arithmetic modes, and on closed-form Kalman results."""
import numpy as np
import pytest

import oracle as O
from advancedps_b200 import _abi, models

MODES = [O.SEQ, O.CANON]


@pytest.mark.parametrize("mode", MODES)
def test_container_identities(mode):
    """test/container.jl:45-68: weights, logZ and ESS of 3 particles with log-likelihoods 0,-1,-2."""
    logps = np.array([0.0, -1.0, -2.0])
    z = np.zeros(3)
    assert np.array_equal(O.softmax(z, mode), np.full(3, 1 / 3))          # :46
    assert O.logsumexp(z, mode) == pytest.approx(np.log(3), abs=1e-15)   # :48
    assert O.ess(z, mode) == 3                                             # :49
    for k in (1, 2):                                                       # :52-68
        lw = k * logps
        ref = np.exp(lw) / np.exp(lw).sum()
        assert np.allclose(O.softmax(lw, mode), ref, rtol=0, atol=1e-15)
        assert O.logsumexp(lw, mode) == pytest.approx(np.log(np.exp(lw).sum()), abs=1e-15)


@pytest.mark.parametrize("mode", MODES)
def test_constant_likelihood_evidence(mode):
    """test/smc.jl:104: two observations of probability 1/2 each -> logevidence = -2 log 2."""
    cfg = _abi.make_config(models.constant_loglik(), 100, 2)
    r = O.sweep(cfg, np.full((2, 1), np.log(0.5)), 100, mode=mode)
    assert r.logevidence == pytest.approx(-2 * np.log(2), abs=1e-14)
    assert np.array_equal(r.final_w, np.full(100, 0.01))  # bare resampler: final set resampled


@pytest.mark.parametrize("mode", MODES)
def test_after_resampling_weights_are_uniform(mode):
    """test/container.jl:86-99: after resample_propagate! logWs == 0, weights 1/N, logZ = log N, ESS = N."""
    m = models.linear_gaussian()
    N, T = 300, 4
    _, Y = O.simulate_data(m, T, 7)
    r = O.sweep(_abi.make_config(m, N, T), Y, 5, mode=mode)
    assert np.all(r.final_logw == 0.0)
    assert np.array_equal(r.final_w, np.full(N, 1 / N))
    if mode == O.CANON:
        assert O.ess(r.final_logw, mode) == N
    else:  # fp64 order: N only up to rounding for N > a few
        assert O.ess(r.final_logw, mode) == pytest.approx(N, rel=1e-13)
    assert O.logsumexp(r.final_logw, mode) == pytest.approx(np.log(N), abs=1e-14)


@pytest.mark.parametrize("kind", [_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED,
                                  _abi.RESAMPLE_MULTINOMIAL, _abi.RESAMPLE_RESIDUAL])
@pytest.mark.parametrize("mode", MODES)
def test_resampler_proportions(kind, mode):
    """test/resampling.jl:12-15: D = [0.3, 0.4, 0.3], n = 1e6; share of index 2 within 1e-3 n
    (systematic, stratified) / 1e-2 n (multinomial, residual)."""
    D = np.array([0.3, 0.4, 0.3])
    n = 10**6
    idx = O.resample(kind, D, n, key=2024, step=3, mode=mode)
    assert idx.min() >= 1 and idx.max() <= 3
    tol = 1e-3 if kind in (_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED) else 1e-2
    assert abs((idx == 2).sum() - 0.4 * n) <= tol * n
    if kind in (_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED):
        assert np.all(np.diff(idx) >= 0)  # sorted ascending, like the reference's walk


def test_systematic_seq_is_the_reference_walk():
    """Literal check of src/resampling.jl:149-183 on a hand-computed case."""
    w = np.array([0.1, 0.2, 0.3, 0.4])
    # n w cumulative = .4, 1.2, 2.4, 4.0 ; u = .5, 1.5, 2.5, 3.5
    assert list(O.resample_systematic_seq(w, 4, 0.5)) == [2, 3, 4, 4]
    # strict `<` (src/resampling.jl:165): a threshold that only just reaches a cumulative weight stays
    # with the smaller index -- v = 1.0 is not < u = 1.0, so the second draw is still index 1
    assert list(O.resample_systematic_seq(np.array([0.5, 0.5]), 2, 0.0)) == [1, 1]
    assert list(O.resample_systematic_seq(np.array([0.5, 0.5]), 2, 0.25)) == [1, 2]
    with pytest.raises(O.OracleError) as e:  # "sample could not be selected (are the weights normalized?)"
        O.resample_systematic_seq(np.array([0.1, 0.1]), 2, 0.9)
    assert e.value.code == 2


def test_randcat_seq_walk():
    """src/resampling.jl:11-21: `while cp <= r && s < n`."""
    p = np.array([0.25, 0.25, 0.5])
    assert O.randcat_seq(p, 0.0) == 1
    assert O.randcat_seq(p, 0.25) == 2      # cp <= r advances
    assert O.randcat_seq(p, 0.4999) == 2
    assert O.randcat_seq(p, 0.5) == 3
    assert O.randcat_seq(p, 0.9999999) == 3


def test_empty_weights_are_an_error():
    """src/resampling.jl:103,154: "weight vector is empty"."""
    with pytest.raises(O.OracleError) as e:
        O.resample(_abi.RESAMPLE_SYSTEMATIC, np.zeros(0), 5)
    assert e.value.code == 1


@pytest.mark.parametrize("kind", [_abi.RESAMPLE_SYSTEMATIC, _abi.RESAMPLE_STRATIFIED])
def test_canon_equals_seq_on_well_separated_weights(kind):
    """CANON (exact integer) and SEQ (reference fp64 order) pick the same ancestors unless a
    threshold falls within rounding distance of a cumulative weight."""
    rng = np.random.default_rng(3)
    total = 0
    for rep in range(20):
        w = rng.dirichlet(np.ones(5000))
        a = O.resample(kind, w, 5000, key=rep, step=1, mode=O.SEQ)
        b = O.resample(kind, w, 5000, key=rep, step=1, mode=O.CANON)
        total += int((a != b).sum())
    assert total == 0


@pytest.mark.parametrize("mode", MODES)
def test_lg1_c1_matches_kalman(mode):
    """configs[0] (C1): LG d=1, T=50, N=1000, SMC + systematic; evidence estimator is unbiased."""
    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, 50, 0xDA7A0001)
    ll, _, _ = models.kalman_loglik(m, Y)
    errs = [O.sweep(_abi.make_config(m, 1000, 50), Y, s, mode=mode, history=False).logevidence - ll
            for s in range(16)]
    # E[Z_hat] = Z: the log-estimate is biased low by ~var/2; |mean| stays small at N = 1000
    assert abs(np.mean(errs)) < 0.25
    assert np.std(errs) < 0.6


def test_lg1_filter_moments_match_kalman():
    m = models.linear_gaussian()
    T, N = 20, 200000
    _, Y = O.simulate_data(m, T, 11)
    _, means, covs = models.kalman_loglik(m, Y)
    r = O.sweep(_abi.make_config(m, N, T, ess_threshold=0.5), Y, 1)
    w = r.final_w
    xT = r.x_hist[T - 1][r.anc_hist[T], 0]
    mean = float((w * xT).sum())
    var = float((w * (xT - mean) ** 2).sum())
    assert mean == pytest.approx(means[-1, 0], abs=5e-3)
    assert var == pytest.approx(covs[-1, 0, 0], rel=0.05)


def test_lg4_matches_kalman():
    m = models.lg4()
    T = 10
    _, Y = O.simulate_data(m, T, 0xDA7A0003)
    ll, _, _ = models.kalman_loglik(m, Y)
    r = O.sweep(_abi.make_config(m, 50000, T), Y, 3, history=False)
    assert r.logevidence == pytest.approx(ll, abs=0.6)


def test_same_seed_same_result():
    """test/pgas.jl:99-127."""
    m = models.stochastic_volatility()
    _, Y = O.simulate_data(m, 15, 5)
    cfg = _abi.make_config(m, 500, 15, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
    a = O.sweep(cfg, Y, 10)
    b = O.sweep(cfg, Y, 10)
    c = O.sweep(cfg, Y, 11)
    assert np.array_equal(a.x_hist, b.x_hist) and a.logevidence == b.logevidence
    assert not np.array_equal(a.x_hist, c.x_hist)


@pytest.mark.parametrize("sampler", [_abi.SAMPLER_PG, _abi.SAMPLER_PGAS])
def test_reference_particle_keeps_last_slot(sampler):
    """test/container.jl:91 / src/container.jl:219-224: the retained particle stays in slot N and
    replays its trajectory (src/pgas.jl:69-72)."""
    m = models.linear_gaussian()
    N, T = 50, 8
    _, Y = O.simulate_data(m, T, 9)
    cfg = _abi.make_config(m, N, T, sampler=sampler, ess_threshold=0.5)
    ref = np.linspace(-1, 1, T).reshape(T, 1)
    r = O.sweep(cfg, Y, 4, ref_traj=ref)
    assert np.array_equal(r.x_hist[:, N - 1, :], ref)
    if sampler == _abi.SAMPLER_PG:
        assert np.all(r.anc_hist[:, N - 1] == N - 1)


def test_pgas_ancestor_splice_forced():
    """test/pgas.jl:61-91: when one particle carries all the ancestor weight it becomes the
    reference's ancestor and X_ref[1:c-1] == X_a[1:c-1]. Forced here with a transition density
    so sharp that only the particle sitting on the reference's path has non-zero weight."""
    N, T = 3, 4
    m = _abi.make_model(_abi.OBS_CONST, 1, 1, [0.0], [1.0], [[1.0]], [0.0], [1e-3])
    cfg = _abi.make_config(m, N, T, sampler=_abi.SAMPLER_PGAS, ess_threshold=1.0)
    Y = np.zeros((T, 1))
    base = O.sweep(cfg, Y, 21)                       # unconditional run, to borrow a real path
    path = O.trajectory(cfg, 0, base)                # trajectory of final slot 0
    r = O.sweep(cfg, Y, 21, ref_traj=path)           # same seed: particles 0..N-2 repeat their draws
    # at every PGAS update (2 <= s <= T-1) the reference's ancestor is a particle whose state at
    # time s-1 continues into X_ref[s] with non-negligible transition density: the one on `path`
    for s in range(2, T):
        a = r.anc_hist[s, N - 1]
        x_prev_of_a = r.x_hist[s - 2, r.anc_hist[s - 1, a], 0]
        assert abs(path[s - 1, 0] - x_prev_of_a) < 0.01
    slot = N - 1
    traj = O.trajectory(cfg, slot, r) if r.anc_hist[T, slot] == N - 1 else None
    if traj is not None:
        assert np.array_equal(traj[-1], path[-1])


def test_ess_threshold_semantics():
    """src/container.jl:233-251 + src/resampling.jl:193-204."""
    m = models.linear_gaussian()
    T, N = 12, 2000
    _, Y = O.simulate_data(m, T, 13)
    never = O.sweep(_abi.make_config(m, N, T, ess_threshold=0.0), Y, 1)
    always = O.sweep(_abi.make_config(m, N, T), Y, 1)
    half = O.sweep(_abi.make_config(m, N, T, ess_threshold=0.5), Y, 1)
    assert never.resampled.sum() == 0 and always.resampled.all()
    assert 0 < half.resampled.sum() < T + 1
    for s in range(T + 1):
        assert bool(half.resampled[s]) == (half.ess[s] <= 0.5 * N)
        if not half.resampled[s]:
            assert np.array_equal(half.anc_hist[s], np.arange(N))
    assert np.any(never.final_logw != 0.0)
    # PGAS default threshold 1.0 resamples at the first decision point too (uniform weights: ESS = N)
    one = O.sweep(_abi.make_config(m, N, T, ess_threshold=1.0), Y, 1)
    assert one.resampled[0] == 1 and one.ess[0] == N


def test_not_normalisable_weights():
    cfg = _abi.make_config(models.constant_loglik(), 10, 2)
    with pytest.raises(O.OracleError) as e:
        O.sweep(cfg, np.full((2, 1), -np.inf), 1)
    assert e.value.code == 2


def test_threaded_oracle_is_bit_identical():
    """The oracle's optional threads (bench.py's all-cores figure) only split particle-independent
    loops and exact integer / max reductions: same results for any thread count."""
    m = models.linear_gaussian()
    N, T = 20000, 6
    _, Y = O.simulate_data(m, T, 3)
    cfg = _abi.make_config(m, N, T, ess_threshold=0.5)
    a = O.sweep(cfg, Y, 11, mode=O.CANON)
    O.set_threads(4)
    try:
        b = O.sweep(cfg, Y, 11, mode=O.CANON)
    finally:
        O.set_threads(1)
    assert a.logevidence == b.logevidence
    assert np.array_equal(a.x_hist, b.x_hist) and np.array_equal(a.anc_hist, b.anc_hist)
    assert np.array_equal(a.ess, b.ess) and np.array_equal(a.final_w, b.final_w)
