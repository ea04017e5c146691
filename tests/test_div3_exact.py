"""The device's vector / scalar division (csrc/rt_math.cuh: div3_exact), host build, against the three IEEE
divisions the reference performs (`Vector3d / f64`, src/algebra/mod.rs:299-317; `normalize`, :107-110).
The device computes the quotients from ONE correctly rounded reciprocal plus two FMA correction steps per
component (nvcc's own division takes a long slow path for zero numerators -- every axis-aligned normal); the
normals, scatter directions and ray directions are only bit-identical to the reference's if every quotient is
the correctly rounded one, signed zeros included."""
import numpy as np

from rs_pathtracing_b200 import _ffi


def div3(a, s):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3)
    s = np.ascontiguousarray(s, dtype=np.float64).reshape(-1)
    assert len(a) == len(s)
    q = np.empty_like(a)
    assert _ffi.core().rt_div3_exact(a.ctypes.data, s.ctypes.data, len(s), q.ctypes.data) == 0
    return q


def check(a, s):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 3)
    s = np.asarray(s, dtype=np.float64).reshape(-1)
    with np.errstate(all="ignore"):
        want = a / s[:, None]
    got = div3(a, s)
    same = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))
    if not same.all():
        i, k = np.argwhere(~same)[0]
        raise AssertionError(f"{float(a[i, k]).hex()} / {float(s[i]).hex()}: got {float(got[i, k]).hex()}, "
                             f"want {float(want[i, k]).hex()}")


def from_parts(mant, exp):
    bits = ((exp.astype(np.int64) + 1023).astype(np.uint64) << np.uint64(52)) | (mant & np.uint64((1 << 52) - 1))
    return bits.view(np.float64)


def test_normalize_like_operands():
    """what normalize() feeds it: components of vectors of every length, divided by the length"""
    rng = np.random.default_rng(3)
    n = 2_000_000
    v = rng.standard_normal((n, 3)) * 10.0 ** rng.uniform(-8, 8, (n, 1))
    v[rng.random(n) < 0.3, 0] = 0.0           # axis-aligned normals: exact zeros, both signs
    v[rng.random(n) < 0.3, 1] = -0.0
    length = np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2])
    keep = length > 0
    check(v[keep], length[keep])


def test_random_mantissas_and_adversarial_divisors():
    """full-range mantissas; divisors that are powers of two, all-ones mantissas (where a rounded reciprocal is
    least accurate) and tiny mantissas; numerators next to the top of a binade"""
    rng = np.random.default_rng(4)
    n = 1_500_000
    ma = rng.integers(0, 1 << 52, (n, 3), dtype=np.uint64)
    ms = rng.integers(0, 1 << 52, n, dtype=np.uint64)
    mode = rng.integers(0, 6, n)
    ms[mode == 1] = (1 << 52) - 1
    ms[mode == 2] = 0
    ms[mode == 3] = rng.integers(0, 256, int((mode == 3).sum()), dtype=np.uint64)
    ms[mode == 4] = np.uint64((1 << 52) - 1) - rng.integers(0, 256, int((mode == 4).sum()), dtype=np.uint64)
    ma[mode == 5] = np.uint64((1 << 52) - 1) - rng.integers(0, 8, (int((mode == 5).sum()), 3), dtype=np.uint64)
    a = from_parts(ma, rng.integers(-40, 40, (n, 3))) * rng.choice([-1.0, 1.0], (n, 3))
    s = from_parts(ms, rng.integers(-40, 40, n))
    check(a, s)


def test_quotients_next_to_representable_values():
    """a = s * k moved by a few ulps: the exact quotient lies within 2^-52 relative of a double or of a midpoint"""
    rng = np.random.default_rng(5)
    n = 1_500_000
    s = from_parts(rng.integers(0, 1 << 52, n, dtype=np.uint64), rng.integers(-20, 20, n))
    k = from_parts(rng.integers(0, 1 << 52, (n, 3), dtype=np.uint64), rng.integers(-20, 20, (n, 3)))
    a = s[:, None] * k
    a = (a.view(np.int64) + rng.integers(-2, 3, (n, 3))).view(np.float64)
    check(a, s)


def test_operands_outside_the_fast_range_take_the_plain_division():
    """zero / negative / subnormal / huge / non-finite divisors and numerators: the guard hands them to `/`"""
    sp = np.array([0.0, -0.0, -1.5, 5e-324, 2.2250738585072014e-308, 1e-200, 1e200, 1.7976931348623157e308,
                   np.inf, -np.inf, np.nan, 1.0, 3.0])
    ap = np.array([0.0, -0.0, 5e-324, -5e-324, 1e-310, 1e-200, -1e200, 1.7976931348623157e308, np.inf, -np.inf,
                   np.nan, 1.0, -7.0])
    a = np.array([[x, y, z] for x in ap for y in ap[:5] for z in ap[5:9]])
    for s in sp:
        check(a, np.full(len(a), s))
    # the edges of the fast range, 2^-500 and 2^500
    rng = np.random.default_rng(6)
    n = 200_000
    a = from_parts(rng.integers(0, 1 << 52, (n, 3), dtype=np.uint64), rng.integers(-503, -497, (n, 3)))
    s = from_parts(rng.integers(0, 1 << 52, n, dtype=np.uint64), rng.integers(497, 503, n))
    check(a, s)
    check(1.0 / a, 1.0 / s)


def test_negative_divisors_like_the_cube_slabs():
    """Cube::ray_intersect (shapes/mod.rs:250-256) divides (-1 - o) and (1 - o) by the direction's components, of
    either sign: the device takes |d| and flips the numerators (div2_exact shares the code path).  Signed zeros
    included: (+0) / (-x) = -0."""
    rng = np.random.default_rng(8)
    n = 1_500_000
    o = rng.uniform(-3, 3, (n, 3))
    o[rng.random(n) < 0.05, 0] = -1.0          # the numerator -1 - o is then exactly +0 / -0
    o[rng.random(n) < 0.05, 1] = 1.0
    d = rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 1, n)
    num = np.stack([-1.0 - o[:, 0], 1.0 - o[:, 1], -1.0 - o[:, 2]], axis=1)
    check(num, d)
    check(-num, -np.abs(d))
    ma = rng.integers(0, 1 << 52, (n, 3), dtype=np.uint64)
    ms = rng.integers(0, 1 << 52, n, dtype=np.uint64)
    a = from_parts(ma, rng.integers(-40, 40, (n, 3))) * rng.choice([-1.0, 1.0], (n, 3))
    s = -from_parts(ms, rng.integers(-40, 40, n))
    check(a, s)
