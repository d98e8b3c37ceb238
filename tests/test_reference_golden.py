"""Pins the oracle's SEQ mode (the restatement of the reference's fp64 order) on vectors produced
by the REAL reference: tests/golden/reference_*.json, written by bench/julia/dump_golden.jl on any
machine with Julia + AdvancedPS.jl. The build image has no Julia, so until someone commits those
files these tests skip and the parity of the oracle with Julia's bit streams stays UNPINNED
(DESIGN.md section 2); everything RNG-independent the reference's own tests assert is pinned in
tests/test_oracle_known_answers.py regardless."""
import json
import os

import numpy as np
import pytest

import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} absent: parity with the Julia reference unpinned (run bench/julia/dump_golden.jl)")
    with open(path) as f:
        return json.load(f)


def fl(v):
    return np.array([float(x) for x in v], dtype=np.float64)  # "Inf" / "-Inf" / "NaN" strings included


def test_resamplers_against_reference_vectors():
    """(weights, uniforms, n) -> indices of src/resampling.jl:11-21,98-183, index for index."""
    recs = load("reference_resample.json")["records"]
    seen = set()
    for r in recs:
        w, n, u, want = fl(r["w"]), int(r["n"]), fl(r["u"]), np.array(r["indices"], dtype=np.int64)
        if r["kind"] == "systematic":
            got = O.resample_systematic_seq(w, n, float(u[0]))
        elif r["kind"] == "stratified":
            got = O.resample_stratified_seq(w, n, u)
        elif r["kind"] == "randcat":
            got = np.array([O.randcat_seq(w, float(u[0]))])
        elif r["kind"] == "residual":
            # upstream only gets here when no residual slot remains: deterministic copies in order
            got = O.resample_residual_seq(w, n, np.zeros(1))
        else:   # multinomial draws through a third-party alias table: law only (proportions below)
            assert want.min() >= 1 and want.max() <= w.size
            assert np.all(w[want - 1] > 0)
            continue
        seen.add(r["kind"])
        assert np.array_equal(got, want), (r["kind"], r["case"], n)
    assert {"systematic", "stratified", "randcat"} <= seen


def test_weight_helpers_against_reference_vectors():
    """getweights / logZ / ESS of src/container.jl:95-119. The oracle's exp / log are within 1 ulp of
    libm (tests/test_oracle_math.py), Julia's are too; sums of m terms: tolerance 4 m ulp."""
    for r in load("reference_weights.json")["records"]:
        lw = fl(r["logWs"])
        tol = 4 * lw.size * np.finfo(float).eps
        assert np.allclose(O.softmax(lw, O.SEQ), fl(r["getweights"]), rtol=tol, atol=0)
        assert O.logsumexp(lw, O.SEQ) == pytest.approx(float(r["logZ"]), rel=tol, abs=tol)
        assert O.ess(lw, O.SEQ) == pytest.approx(float(r["ess"]), rel=tol)
        # CANON (what the GPU reproduces) against the reference: truncation to 2^-(62 - log2 m)
        assert np.allclose(O.softmax(lw, O.CANON), fl(r["getweights"]), rtol=1e-9, atol=1e-15)
        assert O.logsumexp(lw, O.CANON) == pytest.approx(float(r["logZ"]), rel=1e-9, abs=1e-9)


def test_known_answers_against_reference_vectors():
    k = load("reference_known.json")
    for mode in (O.SEQ, O.CANON):
        assert np.allclose(O.softmax(np.zeros(3), mode), fl(k["uniform3"]["getweights"]), rtol=0, atol=1e-16)
        assert O.ess(np.zeros(3), mode) == pytest.approx(float(k["uniform3"]["ess"]), rel=1e-15)
        for name, lw in (("logps1", [0.0, -1.0, -2.0]), ("logps2", [0.0, -2.0, -4.0])):
            assert np.allclose(O.softmax(np.array(lw), mode), fl(k[name]["getweights"]), rtol=0, atol=1e-15)
            assert O.logsumexp(np.array(lw), mode) == pytest.approx(float(k[name]["logZ"]), abs=1e-15)
    assert k["defaults"] == {"SMC": 0.5, "PG": 0.5, "PGAS": 1.0}
    for name, cnt in k["proportions"].items():
        tol = 1e-2 if ("multinomial" in name or "residual" in name) else 1e-3
        assert abs(cnt - 0.4e6) <= tol * 1e6
