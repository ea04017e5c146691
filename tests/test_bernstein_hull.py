"""The marcher's miss proof (csrc/rt_march.cuh (3)): the Bernstein hull of a polynomial over [0, L], tightened by two
levels of de Casteljau subdivision, host build (rt_bernstein_clear) against dense sampling.

What the device relies on is SOUNDNESS: when the test says "clear", |p| really exceeds the threshold with constant sign
on the whole interval -- a ray for which it says so is dropped without marching (k_march_filter), and a wrong "clear"
would lose a hit.  Completeness only costs speed; it is measured, not required."""
import ctypes as C

import numpy as np
import pytest

from rs_pathtracing_b200 import _ffi


def clear(c, L, thr):
    c = np.ascontiguousarray(c, dtype=np.float64)
    out = C.c_int(-1)
    assert _ffi.core().rt_bernstein_clear(c.ctypes.data, len(c) - 1, float(L), float(thr), C.byref(out)) == 0
    return bool(out.value)


def heart_along_ray(o, d):
    """coefficients of f(o + tau d) for the Heart (ray_marching.rs:147-155), by polynomial arithmetic"""
    P = np.polynomial.polynomial
    x, y, z = (np.array([o[k], d[k]]) for k in range(3))
    x2, y2, z2 = P.polymul(x, x), P.polymul(y, y), P.polymul(z, z)
    z3 = P.polymul(z2, z)
    a = P.polysub(P.polyadd(P.polyadd(x2, 2.25 * y2), z2), [1.0])
    return P.polysub(P.polysub(P.polymul(P.polymul(a, a), a), P.polymul(x2, z3)), (9.0 / 80.0) * P.polymul(y2, z3))


@pytest.mark.parametrize("degree", [4, 6])
def test_clear_means_clear_on_random_polynomials(degree):
    rng = np.random.default_rng(degree)
    said_clear = 0
    for trial in range(4000):
        L = 10.0 ** rng.uniform(-2, 2)
        roots_inside = rng.random() < 0.5
        # a polynomial built from its roots: real ones inside / outside [0, L] and complex pairs near the interval
        roots = []
        while len(roots) < degree:
            if degree - len(roots) >= 2 and rng.random() < 0.5:
                re, im = rng.uniform(-0.5 * L, 1.5 * L), 10.0 ** rng.uniform(-3, 0) * L
                roots += [complex(re, im), complex(re, -im)]
            else:
                roots.append(rng.uniform(0, L) if roots_inside and rng.random() < 0.3
                             else rng.choice([-1, 1]) * rng.uniform(1.001, 3.0) * L + (L if rng.random() < 0.5 else 0.0))
        c = np.real(np.polynomial.polynomial.polyfromroots(roots)) * rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(-3, 3)
        xs = np.linspace(0.0, L, 20001)
        v = np.polynomial.polynomial.polyval(xs, c)
        thr = 10.0 ** rng.uniform(-6, -1) * np.abs(v).max()
        if clear(c, L, thr):
            said_clear += 1
            assert (np.sign(v) == np.sign(v[0])).all() and np.abs(v).min() > thr * (1 - 1e-9), (trial, c, L, thr)
    assert said_clear > 200   # (the test does fire: otherwise this checks nothing)


def test_heart_chords_soundness_and_success_rate():
    """random chords of the Heart's bounding ellipsoid (radii 1.45, 1.45 / 2.05, 1.45, ray_marching.rs:126-131): every
    chord the test calls clear misses the Heart; it proves > 85 % of the chords that do miss (93 % measured)"""
    rng = np.random.default_rng(1)
    R = np.array([1.45, 1.45 / 2.05, 1.45])
    misses = proven = 0
    for _ in range(6000):
        a, b = rng.normal(size=3), rng.normal(size=3)
        o, e = a / np.linalg.norm(a) * R, b / np.linalg.norm(b) * R
        c = heart_along_ray(o, e - o)          # tau in [0, 1]
        v = np.polynomial.polynomial.polyval(np.linspace(0, 1, 4001), c)
        really_misses = bool((v > 0).all())
        said = clear(c, 1.0, 1e-9)
        assert not said or really_misses
        misses += really_misses
        proven += said
    assert misses > 1000 and proven > 0.85 * misses, (misses, proven)


def test_end_values_and_arguments():
    # p(0) inside the threshold: never clear, whatever the rest looks like
    assert not clear([1e-12, 1.0, 0, 0, 0, 0, 0], 1.0, 1e-9)
    assert clear([1.0, 0, 0, 0, 0, 0, 0], 5.0, 0.5) and not clear([1.0, 0, 0, 0, 0, 0, 0], 5.0, 1.0)
    assert clear([-2.0, 0, 0, 0, 0], 3.0, 1.0)                       # negative sign, degree 4
    assert not clear([1.0, -2.0, 0, 0, 0], 1.0, 0.0)                 # 1 - 2x changes sign at 0.5
    assert not clear([float("nan"), 1, 1, 1, 1, 1, 1], 1.0, 0.0)     # NaN: nothing is proven
    out = C.c_int(0)
    c = np.zeros(6)
    assert _ffi.core().rt_bernstein_clear(c.ctypes.data, 5, 1.0, 0.0, C.byref(out)) != 0
