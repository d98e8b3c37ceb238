"""Shared arithmetic (include/aps_math.h) pinned on published vectors and on libm."""
import math

import numpy as np

import oracle as O


def test_philox2x64_10_known_answers():
    """Random123's published Philox2x64-10 known-answer vectors (kat_vectors: philox2x64 10)."""
    assert O.philox2x64(0, 0, 0) == (0xCA00A0459843D731, 0x66C24222C9A845B5)
    m = 2**64 - 1
    assert O.philox2x64(m, m, m) == (0x65B021D60CD8310F, 0x4D02F3222F86DF20)
    assert O.philox2x64(0x243F6A8885A308D3, 0x13198A2E03707344, 0xA4093822299F31D0) == (
        0x0A5E742C2997341C, 0xB0F883D38000DE5D)


def _max_ulp(f, ref, xs):
    worst = 0.0
    for x in xs:
        a, b = f(x), ref(x)
        if b == 0.0 or math.isinf(b):
            assert a == b
            continue
        worst = max(worst, abs(a - b) / math.ulp(b))
    return worst


def test_exp_log_within_one_ulp_of_libm():
    rng = np.random.default_rng(0)
    assert _max_ulp(O.exp, math.exp, rng.uniform(-745, 709, 20000)) <= 1.0
    assert _max_ulp(O.exp, math.exp, rng.uniform(-40, 1, 20000)) <= 1.0
    assert _max_ulp(O.log, math.log, np.exp(rng.uniform(-700, 700, 20000))) <= 1.0
    assert _max_ulp(O.log, math.log, rng.uniform(0, 1, 20000)) <= 1.0


def test_exp_log_special_values():
    assert O.exp(0.0) == 1.0
    assert O.exp(-1000.0) == 0.0
    assert O.exp(float("-inf")) == 0.0
    assert math.isinf(O.exp(1000.0))
    assert math.isnan(O.exp(float("nan")))
    assert O.log(1.0) == 0.0
    assert O.log(0.0) == float("-inf")
    assert math.isnan(O.log(-1.0))
    assert O.log(5e-324) == math.log(5e-324)  # subnormal


def test_sincospi():
    rng = np.random.default_rng(1)
    for t in rng.uniform(0, 2, 5000):
        s, c = O.sincospi(t)
        assert abs(s - math.sin(math.pi * t)) < 2e-15
        assert abs(c - math.cos(math.pi * t)) < 2e-15
    for t, (s0, c0) in {0.0: (0, 1), 0.5: (1, 0), 1.0: (0, -1), 1.5: (-1, 0)}.items():
        s, c = O.sincospi(t)
        assert abs(s - s0) < 1e-16 and abs(c - c0) < 1e-16


def test_box_muller_moments():
    rng = np.random.default_rng(2)
    ws = rng.integers(0, 2**64, size=(50000, 2), dtype=np.uint64)
    z = np.array([O.normal_pair(int(a), int(b)) for a, b in ws]).ravel()
    assert abs(z.mean()) < 0.02
    assert abs(z.var() - 1.0) < 0.02
    assert abs(((z - z.mean()) ** 4).mean() / z.var() ** 2 - 3.0) < 0.1
    assert np.isfinite(z).all()
    # extreme words stay finite: u1 -> (k + 1/2) 2^-52 never hits 0 or 1
    for w0 in (0, 2**64 - 1):
        a, b = O.normal_pair(w0, 12345)
        assert math.isfinite(a) and math.isfinite(b)


def test_weight_shift_keeps_sums_in_62_bits():
    for n in (1, 2, 3, 1000, 10**6, 4 * 10**6, 8 * 10**6, 2**31 - 1):
        S = O.weight_shift(n)
        assert n * 2**S <= 2**62
        assert S >= 31
