"""Torus + solve_quantic_equation (src/world/shapes/mod.rs:403-494, src/algebra/equation.rs:17-67).

The solver runs on num::Complex<f64> and a root only counts as real when |im| < 1e-15 (shapes/mod.rs:456): its
accept / reject decision depends on the last ulp of libm's hypot / atan2 / cos / sin / cbrt, so neither the oracle
(glibc) nor the CUDA path (CUDA's libm) can be bit-identical to a Rust build -- or to each other.  The gate is
therefore, as VERDICT r1 item 10 asks: known answers by hand to 1e-9, t within 1e-9 relative where both sides hit
the same shape, a bounded share of rays on which the accept decision differs, and the statistical frame gate.
On the device a Torus never enters the hot loop: a ray whose line touches its bounding ball is replayed through
the literal ShapeCollection loop (k_replay), so FAST == BRUTE exactly.
"""
import json
import math
import os

import numpy as np
import pytest

import rs_pathtracing_b200 as rt
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IDENT = {"translate": [0, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}
CAMERA = {"position": [0, 0, -10], "direction": [0, 0, 1], "up": [0, 1, 0], "fov": 40.0, "focal_length": 1.0}
GREY = {"type": "Lambertian", "albedo": {"type": "SolidColor", "color": [0.5, 0.5, 0.5]}}


def torus_scene(radius=0.5, tube=0.1, transform=IDENT):
    text = json.dumps({"camera": CAMERA, "background": [0, 0, 0], "materials": {"M": GREY},
                       "shapes": [{"type": "Torus", "name": "T", "radius": radius, "tube_radius": tube,
                                   "transform": transform, "material": "M"}]})
    return rt.Scene.from_json(text, add_random_spheres=False)


# ---- the solver --------------------------------------------------------------------------------------
# equation.rs:76-119 (test_solve_quantic_equation) only prints; its three polynomials have roots one can state:
#   3 x^4 + 6 x^3 - 123 x^2 - 126 x + 1080 = 3 (x - 5)(x - 3)(x + 4)(x + 6)
#   the other two are compared with numpy's companion-matrix eigenvalues (an unrelated algorithm)
@pytest.mark.parametrize("coef,want", [
    ((3.0, 6.0, -123.0, -126.0, 1080.0), [5.0, 3.0, -4.0, -6.0]),
    ((-20.0, 5.0, 17.0, -29.0, 87.0), None),
    ((1.0, -4.0, 6.48, -4.96, 1.0376), None),
])
def test_quartic_roots(coef, want):
    got = po.solve_quartic(*coef)
    want = np.array(want, dtype=complex) if want is not None else np.roots(coef)
    assert len(got) == 4
    for w in want:   # every expected root is among the four, and the four are all roots
        assert np.min(np.abs(got - w)) < 1e-9 * max(1.0, abs(w))
    for g in got:
        assert abs(np.polyval(coef, g)) < 1e-8 * sum(abs(c) * max(1.0, abs(g)) ** (4 - k) for k, c in enumerate(coef))


# ---- one ray by hand -----------------------------------------------------------------------------------
# Torus radius 0.5, tube 0.1 in the z = 0 plane; ray o = (-10, 0, 0), d = +x crosses the tube at x = -0.6, -0.4, 0.4, 0.6:
#   shapes:429-452: the smallest root accepted as real and inside [min_t, max_t] is t = 10 - 0.6 = 9.4
#   p = (-0.6, 0, 0); normal = p - normalize((p.x, p.y, 0)) * radius = (-0.6 + 0.5, 0, 0) = (-0.1, 0, 0) -> (-1, 0, 0)
#   theta = asin(p.z / tube) = 0 -> v = 0; phi = acos(p.z / (radius + tube cos 0)) + pi = 3 pi / 2 -> u = 0.75
HAND = dict(o=(-10.0, 0.0, 0.0), d=(1.0, 0.0, 0.0), t=9.4, point=(-0.6, 0.0, 0.0), normal=(-1.0, 0.0, 0.0), uv=(0.75, 0.0))


def check_hand(got):
    assert got["index"][0] == 0
    assert abs(got["t"][0] - HAND["t"]) < 1e-9
    assert np.allclose(got["point"][0], HAND["point"], atol=1e-9)
    assert np.allclose(got["normal"][0], HAND["normal"], atol=1e-9)
    assert np.allclose(got["uv"][0], HAND["uv"], atol=1e-9)
    assert got["front"][0] == 1


def test_torus_hand_derived_hit_oracle():
    sc = torus_scene()
    rays = np.array([list(HAND["o"]) + list(HAND["d"])], dtype=np.float64)
    check_hand(po.OracleScene(sc.desc()).intersect_batch(rays))
    # the reference's own test ray (shapes:853-860) points away from the torus: None
    away = np.array([[0, 0, -10, 0.42233513247717097, 0.26611434880691537, -0.86649650272494549]], dtype=np.float64)
    assert po.OracleScene(sc.desc()).intersect_batch(away)["index"][0] == -1


def test_torus_bounding_box_and_loader():
    # shapes:486-493: (+-(radius + tube), +-(radius + tube), +-tube) through the transform; the loader reads
    # name / radius / tube_radius / transform / material (shapes:766-789)
    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", "torus.json"), add_random_spheres=False)
    d = sc.desc()
    kinds = [int(d.kind[i]) for i in range(d.n_shapes)]
    assert kinds.count(rt._ffi.RT_SHAPE_TORUS) == 3
    i = kinds.index(rt._ffi.RT_SHAPE_TORUS)
    assert (d.params[8 * i], d.params[8 * i + 1]) == (1.0, 0.35)
    with pytest.raises(Exception):
        rt.Scene.from_json(json.dumps({"camera": CAMERA, "background": [0, 0, 0], "materials": {"M": GREY},
                                       "shapes": [{"type": "Torus", "name": "T", "radius": 1.0, "transform": IDENT,
                                                   "material": "M"}]}), add_random_spheres=False)


def test_torus_cull_ball_is_conservative_cpu():
    """hit => reached: every ray the oracle's Torus accepts passes the FP32 ball pre-test that decides whether the
    device replays the ray (host build of the pre-test, rt_cull_reached)"""
    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", "torus.json"), add_random_spheres=False)
    cam = sc.camera()
    w, h = 64, 48
    rays = np.array([po.get_ray(cam, w, h, x + 0.5, y + 0.5) for y in range(h) for x in range(w)])
    from test_cull_cpu import check
    reached, hits = check(sc, rays)
    d = sc.desc()
    tor = [i for i in range(d.n_shapes) if int(d.kind[i]) == rt._ffi.RT_SHAPE_TORUS]
    assert hits[:, tor].sum() > 100
    assert reached[:, tor].mean() < 0.5   # ... and it does cull


# ---- the CUDA path ---------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_torus_hand_derived_hit_gpu():
    sc = torus_scene()
    rays = np.array([list(HAND["o"]) + list(HAND["d"])], dtype=np.float64)
    for mode in (rt.RT_ISECT_BRUTE, rt.RT_ISECT_FAST):
        check_hand(sc.closest_hit(rays, mode=mode))


@pytest.mark.gpu
def test_torus_nearest_hit_against_the_oracle():
    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", "torus.json"), add_random_spheres=False)
    cam = sc.camera()
    w, h = 160, 120
    rays = np.array([po.get_ray(cam, w, h, x + 0.5, y + 0.5) for y in range(h) for x in range(w)])
    want = po.OracleScene(sc.desc()).intersect_batch(rays)
    brute = sc.closest_hit(rays, mode=rt.RT_ISECT_BRUTE)
    fast = sc.closest_hit(rays, mode=rt.RT_ISECT_FAST)
    # the device's two modes run the same torus arithmetic: identical
    for k in ("index", "t", "normal", "point", "uv", "front"):
        assert np.array_equal(brute[k], fast[k], equal_nan=True), k
    d = sc.desc()
    tor = {i for i in range(d.n_shapes) if int(d.kind[i]) == rt._ffi.RT_SHAPE_TORUS}
    on_torus = np.isin(want["index"], list(tor)) | np.isin(fast["index"], list(tor))
    assert on_torus.sum() > 1500
    same = want["index"] == fast["index"]
    assert same[~on_torus].all()
    hit = same & (want["index"] >= 0)
    rel = np.zeros(len(rays))
    rel[hit] = np.abs(fast["t"][hit] - want["t"][hit]) / np.abs(want["t"][hit])
    # The |im| < 1e-15 acceptance differs between glibc and CUDA's libm on a small share of the torus rays: the ray
    # then misses the torus on one side (another index) or takes the next accepted root (same index, another t).
    differs = on_torus & (~same | (rel > 1e-9))
    assert differs.sum() <= 0.03 * on_torus.sum(), (differs.sum(), on_torus.sum())
    agree = hit & ~differs
    assert rel[agree].max() < 1e-9
    assert np.abs(fast["normal"][agree] - want["normal"][agree]).max() < 1e-6
    # no torus ray is lost to the ball pre-test
    sc.reset_stats()
    sc.closest_hit(rays, mode=rt.RT_ISECT_VERIFY, want=("index",))
    st = sc.stats()
    assert st.verify_rays == 0 and st.verify_false_culls == 0


@pytest.mark.gpu
def test_torus_frame_statistical_gate():
    """BASELINE.md's gate on the torus scene (Lambertian, Metal and Dielectric tori): RMSE <= 1.25 r0 + 1e-3 against
    an independent oracle stream, r0 = the RMSE between two independent oracle frames; mean luminance within 0.5 %"""
    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", "torus.json"), add_random_spheres=False)
    cam = sc.camera()
    w, h, spp, depth = 48, 36, 64, 8
    osc = po.OracleScene(sc.desc())
    a, _ = osc.render(cam, w, h, spp, depth, seed=11, rng="xoshiro")
    b, _ = osc.render(cam, w, h, spp, depth, seed=12, rng="xoshiro")
    got = rt.GpuRenderer(sc, 8, depth, seed=5).render(cam, w, h, spp)
    lum = lambda f: float(np.clip(f, 0, 4).mean())
    r0 = float(np.sqrt(np.mean((np.clip(a, 0, 4) - np.clip(b, 0, 4)) ** 2)))
    r = float(np.sqrt(np.mean((np.clip(got, 0, 4) - np.clip(a, 0, 4)) ** 2)))
    assert r <= 1.25 * r0 + 1e-3, (r, r0)
    assert abs(lum(got) - 0.5 * (lum(a) + lum(b))) <= 0.005 * lum(a) + 2e-3, (lum(got), lum(a), lum(b))
