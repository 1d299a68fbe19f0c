"""csrc/eg_math.hpp (exp / ln / pow shared by the host and the device form of the weight update) against 200-bit arithmetic:
the correctly rounded double on every argument checked, over the ranges the update rule feeds them (learning.rs:131-373,
scoring.rs:12-13). The platform libm is within 1 ulp but not correctly rounded; how often it differs is printed, not asserted.
The host == device half of the contract is tests/test_gpu_update_inorder.py::test_rule_arithmetic_host_and_device_agree_bit_for_bit."""
import math

import numpy as np
import pytest

from eirgrid_b200 import _lib

mpmath = pytest.importorskip("mpmath")


def _correctly_rounded(fn, xs, ys):
    mpmath.mp.prec = 200
    out = np.empty(len(xs))
    for i, (x, y) in enumerate(zip(xs, ys)):
        x, y = mpmath.mpf(float(x)), mpmath.mpf(float(y))
        out[i] = float({0: lambda: mpmath.exp(x), 1: lambda: mpmath.log(x), 2: lambda: mpmath.power(x, y)}[fn]())
    return out


@pytest.mark.parametrize("name,fn,make", [
    ("exp(-iwi/500), exp(-iwi/400)", 0, lambda r, n: (-r.integers(0, 60000, n) / r.choice([500.0, 400.0], n), np.zeros(n))),
    ("ln(normalised cost)", 1, lambda r, n: (10 ** r.uniform(0, 4, n), np.zeros(n))),
    ("deterioration ^ 0.3", 2, lambda r, n: (r.uniform(0, 1.0, n), np.full(n, 0.3))),
    ("(iwi / 10) ^ 1.8", 2, lambda r, n: (r.integers(0, 2000000, n) / 10.0, np.full(n, 1.8))),
    ("(iwi / 10) ^ 0.3", 2, lambda r, n: (r.integers(1, 2000000, n) / 10.0, np.full(n, 0.3))),
])
def test_correctly_rounded(name, fn, make):
    rng = np.random.default_rng(11)
    n = 4000
    x, y = make(rng, n)
    got = _lib.rule_math(fn, x, y)
    want = _correctly_rounded(fn, x, y)
    bad = int((got != want).sum())
    assert bad == 0, "%s: %d of %d results are not the correctly rounded double" % (name, bad, n)
    libm = np.array([{0: math.exp, 1: math.log}[fn](v) if fn < 2 else math.pow(v, w) for v, w in zip(x, y)])
    print("%s: platform libm differs from the correctly rounded value on %d of %d arguments" % (name, int((libm != want).sum()), n))


def test_special_values_behave_like_libm():
    """the cases the rule can produce at its edges: pow(0, 1.8) at iwi = 0, a negative deterioration (quirk Q9: NaN), exact ones"""
    f = lambda fn, x, y=0.0: float(_lib.rule_math(fn, np.array([x]), np.array([y]))[0])
    assert f(2, 0.0, 1.8) == 0.0 and f(2, 0.0, 0.3) == 0.0 and f(2, 1.0, 0.3) == 1.0 and f(2, 7.5, 0.0) == 1.0 and f(2, 7.5, 1.0) == 7.5
    assert math.isnan(f(2, -0.25, 0.3)) and f(2, -2.0, 2.0) == 4.0 and f(2, -2.0, 3.0) == -8.0
    assert f(0, 0.0) == 1.0 and f(0, -800.0) == 0.0 and f(0, 800.0) == math.inf and f(0, -745.0) == math.exp(-745.0)
    assert f(1, 1.0) == 0.0 and f(1, 0.0) == -math.inf and math.isnan(f(1, -1.0)) and f(1, 5e-324) == math.log(5e-324)
    # the contrast factors at the first iteration: stagnation 1, boost 1.4 exactly (learning.rs:162-179 with iwi = 0, lr = 0.2)
    assert f(5, 1.0, 0.0) == 1.4
