"""The marcher's exact multi-step advance (csrc/rt_march.cuh: advance_exact), host build, against the
literal loop it replaces: m iterations of `a = a + s` in IEEE double arithmetic.  The reference's marching
loop (ray_marching.rs:37-39: `t += step; p += step * dir`) is such an accumulation, and the candidate t it
returns depends on every rounded partial sum -- the skip is only exact if this function is."""
import ctypes as C

import numpy as np
import pytest

from rs_pathtracing_b200 import _ffi


def advance(a, s, m):
    out = C.c_double()
    assert _ffi.core().rt_advance_exact(float(a), float(s), int(m), C.byref(out)) == 0
    return out.value


def literal(a, s, m):
    a = np.array(a, dtype=np.float64).copy()
    s = np.asarray(s, dtype=np.float64)
    m = np.asarray(m, dtype=np.int64)
    for k in range(int(m.max())):
        live = m > k
        a[live] = a[live] + s[live]
    return a


def check(a, s, m):
    want = literal(a, s, m)
    for ai, si, mi, wi in zip(a, s, m, want):
        got = advance(ai, si, mi)
        assert got == wi or (np.isnan(got) and np.isnan(wi)), (float(ai).hex(), float(si).hex(), int(mi), got, wi)
        assert np.signbit(got) == np.signbit(wi) or got != 0.0


def test_marching_like_accumulators():
    """t and p components as the marcher sees them: steps of 0.01 * d (and the refinement steps -1e-4, 1e-6,
    -1e-8 times d), starting values around 1e-3 .. 1e3, thousands of steps, many zero crossings"""
    rng = np.random.default_rng(1)
    n = 1500
    d = rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-3, 0, n)
    level = rng.integers(0, 4, n)
    step = 0.01 * (-0.01) ** level
    s = step * d
    a = rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-3, 3, n)
    m = np.where(level == 0, rng.integers(1, 24000, n), rng.integers(1, 120, n))
    # force zero crossings for a third of them: start a few hundred steps before zero
    cross = rng.random(n) < 0.33
    a[cross] = -s[cross] * rng.uniform(1, 1500, cross.sum())
    check(a, s, m)


def test_edge_cases():
    tiny = np.nextafter(0.0, 1.0)
    cases = [
        (0.0, 0.01, 1000), (-0.0, 0.01, 3), (0.0, -0.01, 1000), (1.0, 0.0, 5), (1.0, -0.0, 5),
        (1.0, 2.0 ** -60, 100), (1.0, 2.0 ** -53, 1000), (1.0, -(2.0 ** -54), 1000),   # step below / at half an ulp
        (1.0 - 2.0 ** -53, 2.0 ** -54, 10), (2.0, -(2.0 ** -53), 7),                  # binade bottom approached from above
        (1.0, 1.0, 60), (1.0, 3.0, 40), (-5.0, 1.5, 10), (1e-300, 1e-301, 500),
        (tiny * 3, tiny, 50), (tiny * 3, -tiny, 50), (1e308, 1e307, 20), (np.inf, 1.0, 3), (np.nan, 1.0, 3), (1.0, np.nan, 3),
        (0.5, 2.0 ** -53 * 1.5, 300), (0.75, 2.0 ** -52 + 2.0 ** -53, 300),           # exact ties in the rounding
        (123.456, -0.01, 12346), (123.456, -0.0100000001, 20000), (8.0 - 1e-9, 1e-3, 10),
    ]
    a, s, m = (np.array(x) for x in zip(*cases))
    check(a.astype(float), s.astype(float), m.astype(np.int64))


def test_ties_and_power_of_two_steps():
    """steps with few significant bits hit the round-half-even ties that the closed form treats specially"""
    rng = np.random.default_rng(3)
    n = 800
    s = rng.integers(1, 8, n) * 2.0 ** rng.integers(-60, -40, n) * rng.choice([-1.0, 1.0], n)
    a = rng.uniform(0.25, 4.0, n) * rng.choice([-1.0, 1.0], n)
    a = np.where(rng.random(n) < 0.5, np.round(a * 2 ** 20) / 2 ** 20, a)   # short mantissas too
    m = rng.integers(1, 5000, n)
    check(a, s, m)


def test_zero_steps_and_no_steps():
    assert advance(1.5, 0.25, 0) == 1.5
    assert advance(1.5, 0.25, -3) == 1.5
    assert advance(1.5, 0.0, 10) == 1.5


def test_hypothesis_fuzz():
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=1500, deadline=None, derandomize=True)
    @hyp.given(a=st.floats(allow_nan=False, allow_infinity=False, width=64),
               e=st.integers(min_value=-70, max_value=8), frac=st.floats(min_value=1.0, max_value=2.0, exclude_max=True),
               neg=st.booleans(), m=st.integers(min_value=1, max_value=3000))
    def run(a, e, frac, neg, m):
        # the step relative to the accumulator: from 2^-70 |a| (below half an ulp) to 2^8 |a|
        scale = abs(a) if a != 0.0 and np.isfinite(abs(a) * 2.0 ** e * 2) else 1.0
        s = frac * 2.0 ** e * scale * (-1.0 if neg else 1.0)
        if not np.isfinite(s):
            return
        x = np.float64(a)
        with np.errstate(over="ignore"):
            for _ in range(m):
                x = x + np.float64(s)
        got = advance(a, s, m)
        assert got == x or (np.isnan(got) and np.isnan(x)), (float(a).hex(), float(s).hex(), m, got, float(x))

    run()
