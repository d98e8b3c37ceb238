"""The device-resident ParticleContainer driven call by call (aps_pc_*), the way the reference's own
tests drive theirs: the literal port of test/pgas.jl:61-91 ("update reference"), the sweep! loop
of src/container.jl:316-363 over the single calls against the fused aps_sweep and the oracle, the
bulk trajectory export of SMCSample (src/smc.jl:56), and the ownership of returned results."""
import numpy as np
import pytest

import advancedps_b200 as aps
import oracle as O
from advancedps_b200 import _abi, _lib, models

pytestmark = pytest.mark.gpu


def base_model(a, q, r):
    """BaseModel(Params(a, q, r)) of test/pgas.jl:2-40: x1 ~ N(0, q), x' ~ N(a x, q), y ~ N(x, r), Y = zeros(3)."""
    return aps.TracedSSM(models.linear_gaussian(a=a, b=0.0, q=q, h=1.0, r=r, x0=0.0, sigma0=q), np.zeros((3, 1)))


def test_pgas_update_reference_exact():
    """test/pgas.jl:61-91 verbatim: three particles, the third is the reference with a complete
    trajectory; after two steps logWs = [-Inf, 0, -Inf] forces the ancestor update to the second
    particle; then X_2[1:2] == X_ref[1:2] (equality, not tolerance) and the terminal values are
    all distinct."""
    model = base_model(0.9, 0.31, 1.0)
    sampler = aps.PGAS(3)                                   # resampler = ResampleWithESSThreshold(1.0)
    # part = particles[3]; advance! x 3; ref = forkr(part): any complete trajectory serves
    X_part = np.array([[0.3], [-0.2], [0.45]])
    pc = aps.DeviceParticleContainer(model, sampler, rng=np.random.default_rng(31), ref_traj=X_part)

    pc.reweight_()                                          # reweight!(pc, ref)
    pc.resample_propagate_()                                # current_step(ref) <= 2: no update
    pc.reweight_()
    pc.logWs = [-np.inf, 0.0, -np.inf]                      # force ancestor update to second particle
    assert pc.resample_propagate_()                         # ESS = 1 <= 1.0 * 3
    assert pc.reweight_() is False
    X2, Xref = pc.trajectory(1), pc.trajectory(2)           # pc.vals[2], ref (1-based upstream)
    assert np.array_equal(X2[0:2], Xref[0:2])               # all(pc.vals[2].model.X[1:2] .== ref.model.X[1:2])
    assert np.array_equal(Xref[2], X_part[2])               # the reference keeps its own X[3]
    assert not np.array_equal(Xref[0:2], X_part[0:2])       # ... and its past really was replaced
    terminal = [pc.trajectory(i)[2, 0] for i in range(3)]
    assert len(set(terminal)) == 3                          # all distinct
    assert pc.reweight_() is True                           # every particle is done (container.jl:288)
    # both children of the forced resampling descend from the second particle
    assert np.array_equal(pc._h.ancestors(3)[:2], [1, 1]) and pc._h.ancestors(3)[2] == 1


def test_pgas_no_update_before_step_three():
    """src/pgas.jl:114: current_step(ref) <= 2 returns early -- the reference keeps its own past."""
    model = base_model(0.9, 0.31, 1.0)
    X_part = np.array([[0.3], [-0.2], [0.45]])
    pc = aps.DeviceParticleContainer(model, aps.PGAS(3), rng=np.random.default_rng(1), ref_traj=X_part)
    pc.reweight_()
    pc.logWs = [0.0, -np.inf, -np.inf]
    pc.resample_propagate_()
    pc.reweight_()
    pc.reweight_()
    assert np.array_equal(pc.trajectory(2), X_part)
    assert pc._h.ancestors(2)[2] == 2                       # the reference's ancestor is itself


CASES = [
    ("lg1-bare", models.linear_gaussian, 5000, 12, _abi.SAMPLER_SMC, _abi.RESAMPLE_SYSTEMATIC, float("nan")),
    ("lg1-ess", models.linear_gaussian, 5000, 12, _abi.SAMPLER_SMC, _abi.RESAMPLE_SYSTEMATIC, 0.5),
    ("lg1-strat", models.linear_gaussian, 4097, 9, _abi.SAMPLER_SMC, _abi.RESAMPLE_STRATIFIED, 0.5),
    ("lg1-multi", models.linear_gaussian, 4097, 9, _abi.SAMPLER_SMC, _abi.RESAMPLE_MULTINOMIAL, float("nan")),
    ("lg1-resid", models.linear_gaussian, 4097, 9, _abi.SAMPLER_SMC, _abi.RESAMPLE_RESIDUAL, 0.5),
    ("sv-pgas", models.stochastic_volatility, 3000, 10, _abi.SAMPLER_PGAS, _abi.RESAMPLE_SYSTEMATIC, 1.0),
    ("lg4-pg", models.lg4, 3000, 8, _abi.SAMPLER_PG, _abi.RESAMPLE_SYSTEMATIC, 0.5),
]


@pytest.mark.parametrize("name,mk,N,T,smp,res,thr", CASES, ids=[c[0] for c in CASES])
def test_stepwise_loop_equals_sweep_and_oracle(name, mk, N, T, smp, res, thr):
    """sweep! written as its loop over resample_propagate! / logZ / reweight! (src/container.jl:316-363)
    gives, call by call, exactly what the fused aps_sweep and the oracle give."""
    m = mk()
    _, Y = O.simulate_data(m, T, 0xDA7A0001)
    cfg = _abi.make_config(m, N, T, sampler=smp, resampler=res, ess_threshold=thr)
    ref = None
    if smp != _abi.SAMPLER_SMC:
        ref = np.cumsum(np.full((T, m.d), 0.05), axis=0)
    ro = O.sweep(cfg, Y, 77, ref_traj=ref, mode=O.CANON)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    h.pc_begin(77, ref_traj=ref)
    resampled = []
    resampled.append(h.pc_resample_propagate())
    z0 = h.pc_logZ()
    done = h.pc_reweight()
    z1 = h.pc_logZ()
    le = z1 - z0
    while not done:
        resampled.append(h.pc_resample_propagate())
        z0 = h.pc_logZ()
        done = h.pc_reweight()
        z1 = h.pc_logZ()
        le += z1 - z0
    assert le == ro.logevidence
    assert np.array_equal(np.array(resampled, dtype=np.uint8), ro.resampled)
    for t in range(1, T + 1):
        assert np.array_equal(h.states(t), ro.x_hist[t - 1]), f"states differ at t={t}"
    for t in range(2, T + 2):
        assert np.array_equal(h.ancestors(t), ro.anc_hist[t - 1]), f"ancestors differ at t={t}"
    assert np.array_equal(h.weights(), ro.final_w)
    # and the fused sweep on the same handle afterwards
    assert h.sweep(77, ref_traj=ref) == ro.logevidence
    assert np.array_equal(h.ancestors(T + 1), ro.anc_hist[T])


def test_stepwise_logz_and_weights_between_calls():
    """test/container.jl:45-68 on the device container: logZ / weights after each reweight!."""
    m = models.constant_loglik()
    cfg = _abi.make_config(m, 3, 2, ess_threshold=0.0)      # never resample: weights accumulate
    h = _lib.Handle(cfg)
    h.set_observations(np.array([[np.log(0.5)], [np.log(0.25)]]))
    h.pc_begin(1)
    assert h.pc_logZ() == pytest.approx(np.log(3), abs=1e-15)
    h.pc_reweight()
    assert np.array_equal(h.logweights(), np.full(3, np.log(0.5)))
    assert h.pc_logZ() == pytest.approx(np.log(3 * 0.5), abs=1e-15)
    h.set_logweights(np.array([0.0, -1.0, -2.0]))
    w = np.exp([0.0, -1.0, -2.0])
    assert h.pc_logZ() == pytest.approx(np.log(w.sum()), abs=1e-15)
    assert h.pc_reweight() is False
    assert np.allclose(h.logweights(), np.array([0.0, -1.0, -2.0]) + np.log(0.25), rtol=0, atol=1e-15)
    assert np.allclose(h.weights(), w / w.sum(), rtol=0, atol=1e-15)
    assert h.pc_reweight() is True
    h.set_logweights(np.full(3, -np.inf))
    assert h.pc_logZ() == -np.inf
    with pytest.raises(_lib.ApsError) as e:                 # src/resampling.jl:120,169
        h.pc_resample_propagate()
    assert e.value.code == _abi.ERR_WEIGHTS


def test_all_trajectories_at_once():
    """aps_get_trajectories: SMCSample(collect(pc), ...) of src/smc.jl:56 in one call."""
    m = models.lg4()
    N, T = 5000, 11
    cfg = _abi.make_config(m, N, T, ess_threshold=0.5)
    _, Y = O.simulate_data(m, T, 0xDA7A0003)
    ro = O.sweep(cfg, Y, 9, mode=O.CANON)
    h = _lib.Handle(cfg)
    h.set_observations(Y)
    h.sweep(9)
    X = h.trajectories()
    assert X.shape == (T, N, m.d)
    b = ro.anc_hist[T].astype(np.int64)
    for t in range(T, 0, -1):
        assert np.array_equal(X[t - 1], ro.x_hist[t - 1][b]), f"t={t}"
        b = ro.anc_hist[t - 1][b].astype(np.int64)
    for i in (0, 17, N - 1):
        assert np.array_equal(X[:, i, :], h.trajectory(i))


def test_results_are_owned_values():
    """samples must not alias the cached device handle (src/smc.jl:56,127-128 return
    owned values)."""
    m = models.linear_gaussian()
    _, Y = O.simulate_data(m, 6, 1)
    tssm = aps.TracedSSM(m, Y)
    smc = aps.SMC(4096, 0.0)                                # never resample: non-uniform final weights
    rng = np.random.default_rng(3)
    s1 = aps.sample(rng, tssm, smc)
    w1 = s1.weights.copy()
    x1 = s1.trajectories[5].model.X.copy()
    s2 = aps.sample(rng, tssm, smc)
    assert np.array_equal(s1.weights, w1) and not np.array_equal(s2.weights, w1)
    with pytest.raises(aps.ApsError):                       # stale lazy access raises instead of lying
        s1.trajectories[5]
    s3 = aps.sample(np.random.default_rng(3), tssm, smc, materialize=True)
    aps.sample(rng, tssm, smc)
    assert np.array_equal(s3.trajectories[5].model.X, x1)   # materialised before the store was reused
    # two interleaved PG chains on one model: each conditions on its OWN trajectory
    pg = aps.PG(2048)
    ra, rb = np.random.default_rng(10), np.random.default_rng(20)
    sa, sta = aps.step(ra, tssm, pg)
    sb, stb = aps.step(rb, tssm, pg)
    sa2, _ = aps.step(ra, tssm, pg, sta)                    # the handle now holds chain b's pick
    ra0 = np.random.default_rng(10)
    ea, st0 = aps.step(ra0, tssm, pg)
    ea2, _ = aps.step(ra0, tssm, pg, st0)                   # uninterrupted chain a
    assert np.array_equal(sa.trajectory.model.X, ea.trajectory.model.X)
    assert sa2.logevidence == ea2.logevidence
    assert np.array_equal(sa2.trajectory.model.X, ea2.trajectory.model.X)
