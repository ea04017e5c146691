"""GPU parity of the batched nearest-hit kernel (rt_intersect_batch through the C ABI) against the
oracle: shape index bit-exact, t / normal / point bit-exact (the contract allows 1e-5 relative; the
FP64 no-FMA kernel is expected to give 0 ulp and the test says so), uv within 1e-12."""
import math

import numpy as np
import pytest

import rs_pathtracing_b200 as rt
from oracle import pyoracle as po

from conftest import scene_path

pytestmark = pytest.mark.gpu

TRIO = {
    "camera": {"position": [0, 0, -10], "direction": [0, 0, 1], "up": [0, 1, 0], "fov": 40.0, "focal_length": 1.0},
    "background": [0, 0, 0],
    "materials": {"M": {"type": "Lambertian", "albedo": {"type": "SolidColor", "color": [0.9, 0.1, 0.1]}}},
    "shapes": [
        {"type": "Sphere", "name": "Test sphere", "material": "M",
         "transform": {"translate": [0, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}},
        {"type": "Cube", "name": "Test cube", "material": "M",
         "transform": {"translate": [0, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}},
        {"type": "BruteForsableShape", "shape": {"type": "Heart"}, "step": 0.01, "material": "M",
         "transform": {"translate": [0, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}},
    ],
}


def bench_rays(n, seed=42, target_radius=3.0):
    """benches/bench_intersections.rs:69-70: pos = -random_in_sphere(10), ray toward the origin;
    here the target is jittered inside a ball so that about half of the rays miss (SURVEY §8d cfg 2)."""
    rng = np.random.default_rng(seed)

    def ball(m, r):
        out = np.empty((0, 3))
        while out.shape[0] < m:
            v = rng.uniform(-r, r, (2 * m, 3))
            out = np.concatenate([out, v[(v * v).sum(1) <= r * r]])
        return out[:m]

    o = -ball(n, 10.0)
    tgt = ball(n, target_radius)
    return rt.make_rays(o, tgt - o)


def scene_rays(sc, n, seed):
    """half primary-like rays from the camera through the viewport, half rays between random points of
    the scene's bounding region (secondary-like)"""
    import json
    cam = sc.camera()
    rng = np.random.default_rng(seed)
    w = h = 256
    xs, ys = rng.uniform(0, w, n // 2), rng.uniform(0, h, n // 2)
    prim = np.array([po.get_ray(cam, w, h, x, y) for x, y in zip(xs, ys)])
    d = sc.desc()
    dirm = np.ctypeslib.as_array(d.direct, shape=(d.n_shapes, 12))
    centres = dirm[:, [3, 7, 11]]
    scale = np.abs(dirm[:, [0, 5, 10]]).max(1)
    ok = scale < 1e4
    a = centres[ok][rng.integers(0, ok.sum(), n - n // 2)] + rng.normal(size=(n - n // 2, 3)) * (scale[ok].mean() + 1.0)
    b = centres[ok][rng.integers(0, ok.sum(), n - n // 2)] + rng.normal(size=(n - n // 2, 3)) * 0.3
    sec = rt.make_rays(a, b - a)
    return np.concatenate([prim, sec])


def check_parity(sc, rays, t_min=0.001, t_max=math.inf, modes=(rt.RT_ISECT_BRUTE, rt.RT_ISECT_FAST)):
    """both kernels — the literal brute-force loop and the reorganised one (analytic shapes first,
    exact-skip marching) — must reproduce the oracle"""
    osc = po.OracleScene(sc.desc())
    want = osc.intersect_batch(rays, t_min, t_max)
    for mode in modes:
        _compare(sc.closest_hit(rays, t_min, t_max, mode=mode), want, len(rays), mode)
    return want


def _compare(got, want, n, mode):
    assert np.array_equal(got["index"], want["index"]), \
        f"mode {mode}: {(got['index'] != want['index']).sum()} of {n} nearest-hit indices differ"
    hit = want["index"] >= 0
    for k in ("t", "normal", "point"):
        g, w_ = got[k][hit], want[k][hit]
        same = (g == w_) | (np.isnan(g) & np.isnan(w_))
        rel = np.abs(g - w_) / np.maximum(np.abs(w_), 1e-300)
        assert same.all(), f"mode {mode} {k}: {(~same).sum()} values differ, max rel err {np.nanmax(rel[~same]):.3e}"
    assert np.array_equal(got["front"][hit], want["front"][hit])
    assert np.allclose(got["uv"][hit], want["uv"][hit], rtol=0, atol=1e-12, equal_nan=True)


def test_bench_trio_parity():
    import json
    sc = rt.Scene.from_json(json.dumps(TRIO), add_random_spheres=False)
    rays = bench_rays(1 << 17)
    want = check_parity(sc, rays)
    frac = (want["index"] >= 0).mean()
    assert 0.2 < frac < 0.9, frac
    # the unit cube encloses the unit sphere, so the sphere (index 0) can never be the nearest hit
    assert set(np.unique(want["index"])) == {-1, 1, 2}


@pytest.mark.parametrize("name", ["spheres.json", "cornell_box.json", "detached_materials.json", "dupin.json",
                                  "cube_test.json"])
def test_fixture_scene_parity(name):
    sc = rt.Scene.from_file(scene_path(name), random_spheres_seed=1)
    rays = scene_rays(sc, 1 << 14, seed=7)
    want = check_parity(sc, rays)
    assert (want["index"] >= 0).mean() > 0.3
    kinds = sc.shape_kinds()
    assert len(set(kinds[want["index"][want["index"] >= 0]])) >= 2   # more than one shape type wins


def test_finite_t_max_and_t_min():
    sc = rt.Scene.from_file(scene_path("cornell_box.json"), random_spheres_seed=1)
    rays = scene_rays(sc, 4096, seed=3)
    a = check_parity(sc, rays, t_min=0.001, t_max=700.0)
    b = check_parity(sc, rays, t_min=50.0, t_max=math.inf)
    assert (a["index"] != b["index"]).any()


def test_empty_and_ragged_inputs():
    sc = rt.Scene.from_file(scene_path("cube_test.json"), random_spheres_seed=1)
    out = sc.closest_hit(np.zeros((0, 6)))
    assert out["index"].shape == (0,)
    for n in (1, 31, 33, 257):   # not multiples of the warp / block size
        check_parity(sc, scene_rays(sc, 2 * n, seed=n)[:n])
    empty = rt.Scene.from_file(scene_path("empty.json"), add_random_spheres=False)
    out = empty.closest_hit(bench_rays(100))
    assert (out["index"] == -1).all()


def test_degenerate_rays_follow_the_reference():
    """tangent ray (D == 0 accepted without range check, t = -half_b * a),
    origin inside a cube (hit at t_min), ray inside a rectangle's plane (NaN t), zero direction."""
    import json
    scene = json.loads(json.dumps(TRIO))
    scene["shapes"] = [
        {"type": "Rectangle", "x0": -1, "y0": -1, "x1": 1, "y1": 1, "material": "M",
         "transform": {"translate": [0, 0, 5], "rotate": [0, 0, 0], "scale": [1, 1, 1]}},
        {"type": "Cube", "name": "c", "material": "M",
         "transform": {"translate": [10, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}},
        {"type": "Sphere", "name": "s", "material": "M",
         "transform": {"translate": [0, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}},
        {"type": "Sphere", "name": "s2", "material": "M",
         "transform": {"translate": [0, 0, 20], "rotate": [0, 0, 0], "scale": [2, 2, 2]}},
    ]
    sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
    rays = np.array([
        [1.0, 0.0, -10.0, 0.0, 0.0, 1.0],      # tangent to the unit sphere: D == 0
        [2.0, 0.0, -10.0, 0.0, 0.0, 1.0],      # tangent to the scaled sphere behind it
        [10.0, 0.2, 0.1, 0.0, 0.0, 1.0],       # origin inside the cube
        [-5.0, 0.0, 5.0, 1.0, 0.0, 0.0],       # in the rectangle's plane: t = -0/0 = NaN
        [0.0, 0.0, -10.0, 0.0, 0.0, 0.0],      # zero direction
        [0.0, 0.0, -10.0, math.nan, 0.0, 1.0],  # NaN direction
        [0.0, 0.0, -10.0, 0.0, 0.0, 1.0],      # plain hit through everything
    ])
    want = check_parity(sc, rays)
    assert want["index"][2] == 1 and want["t"][2] == 0.001


def test_later_shape_wins_ties():
    """A.3: equal t -> the later shape in the list wins (shrinking max_t accepts t == max_t)"""
    import json
    scene = json.loads(json.dumps(TRIO))
    one = {"type": "Sphere", "name": "s", "material": "M",
           "transform": {"translate": [0, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}}
    scene["shapes"] = [one, dict(one), dict(one)]
    sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
    want = check_parity(sc, bench_rays(2048, target_radius=0.8))
    assert set(np.unique(want["index"])) <= {-1, 2}


def test_marched_candidate_at_the_best_hit_wins_the_tie():
    """A marched shape of depth 0 returns t = the start of its bounding chord (ray_marching.rs:27-57: no loop, then
    the range check).  With a unit Sphere in front of a Sine surface bounded by the unit sphere the two candidates are
    the SAME double (both solve a = d.d, half_b = d.o, c = o.o - 1), and ShapeCollection's loop lets the later shape
    win the tie (shapes/mod.rs:573-597).  The marcher's prune `chord starts beyond the best hit` must therefore be
    strict: start == best still has to be marched"""
    import json
    scene = json.loads(json.dumps(TRIO))
    ident = {"translate": [0, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}
    scene["shapes"] = [
        {"type": "Sphere", "name": "s", "material": "M", "transform": ident},
        {"type": "BruteForsableShape", "name": "m", "material": "M", "transform": ident, "step": 0.01, "depth": 0,
         "shape": {"type": "Sine", "a": 0.7, "sphere_radius": 1.0}},
    ]
    sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
    want = check_parity(sc, bench_rays(2048, target_radius=0.8))
    assert (want["index"] == 1).mean() > 0.9   # (a handful of rays: the two roots differ in the last place)


def test_depth_zero_marched_shape_is_hit_at_the_start_of_its_chord():
    """depth 0: `for _ in 0..self.depth` runs no iteration, the candidate is t = start for EVERY ray that enters the
    bound (ray_marching.rs:27-57) -- also for the rays whose chord never comes near the surface, which the miss proof
    of rt_march.cuh (3) must therefore leave alone"""
    import json
    scene = json.loads(json.dumps(TRIO))
    ident = {"translate": [0, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}
    scene["shapes"] = [{"type": "BruteForsableShape", "name": "m", "material": "M", "transform": ident, "step": 0.01,
                        "depth": 0, "shape": {"type": "Star", "a": 1.3, "sphere_radius": 2.0}}]
    sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
    want = check_parity(sc, bench_rays(2048, target_radius=1.9))
    assert (want["index"] == 0).mean() > 0.9


SURFACES = {
    "Heart": {"type": "Heart"},
    "Sine": {"type": "Sine", "a": 0.7, "sphere_radius": 2.0},
    "Star": {"type": "Star", "a": 1.3, "sphere_radius": 2.0},
    "DupinCyclide": {"type": "DupinCyclide", "a": 1.11, "b": 0.99, "c": 0.5, "d": 0.1, "sphere_radius": 2.5},
    "HuntsSurface": {"type": "HuntsSurface", "sphere_radius": 5.0},
    "Cushion": {"type": "Cushion", "sphere_radius": 1.5},
}


@pytest.mark.parametrize("surface", list(SURFACES))
@pytest.mark.parametrize("scale,rotate,step,depth", [
    (1.0, [0.0, 0.0, 0.0], 0.01, 4),           # axis-aligned object space
    (82.5, [-95.0, -18.0, 0.0], 0.01, 4),      # cornell_box.json's heart: ~24 000 steps per chord
    (2.0, [0.0, -100.0, 0.0], 0.01, 4),        # dupin.json's transform
    (1.0, [10.0, 20.0, 30.0], 0.003, 3),
])
def test_exact_skip_marching_is_bit_exact(surface, scale, rotate, step, depth):
    """the exact-skip marcher (binade arithmetic progression + derivative bound) must return the very
    t the plain loop returns, for every implicit surface, at object scales from 1 to 82.5"""
    import json
    scene = json.loads(json.dumps(TRIO))
    R = 2.6 * scale if surface != "HuntsSurface" else 5.5 * scale
    scene["shapes"] = [
        {"type": "BruteForsableShape", "shape": SURFACES[surface], "step": step, "depth": depth, "material": "M",
         "transform": {"translate": [3.0, -2.0, 5.0], "rotate": rotate, "scale": [scale] * 3}},
        # a wall behind half of the object: exercises the clipped end and the `start < best` pruning
        {"type": "Rectangle", "x0": -1e4, "y0": -1e4, "x1": 1e4, "y1": 1e4, "material": "M",
         "transform": {"translate": [3.0, -2.0, 5.0 + 0.3 * R], "rotate": [0, 0, 0], "scale": [1, 1, 1]}},
    ]
    sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
    rng = np.random.default_rng(hash(surface) % 1000 + int(scale))
    centre = np.array([3.0, -2.0, 5.0])
    n = 3000 if scale > 10 else 6000
    def ball(m, r):
        v = rng.normal(size=(m, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        return v * r * rng.uniform(0, 1, (m, 1)) ** (1 / 3)
    o = centre + ball(n, 4.0 * R)                       # outside and inside the bound
    tgt = centre + ball(n, 0.9 * R)
    rays = rt.make_rays(o, tgt - o)
    # axis-aligned rays through the centre region: some step*dir components are exactly zero
    ax = []
    for a in range(3):
        for s in (-1.0, 1.0):
            for off in rng.uniform(-0.5, 0.5, (8, 3)) * scale:
                d = np.zeros(3); d[a] = s
                oo = centre + off - d * 3.0 * R
                oo[(a + 1) % 3] = centre[(a + 1) % 3]     # exactly on the object's coordinate plane
                ax.append(np.concatenate([oo, d]))
    rays = np.concatenate([rays, np.array(ax)])
    want = check_parity(sc, rays, modes=(rt.RT_ISECT_FAST,))
    assert (want["index"] == 0).mean() > 0.05, "the test rays must actually hit the surface"


def test_exact_skip_skips(monkeypatch):
    """the skipping marcher evaluates the polynomial far fewer times than the plain loop"""
    sc = rt.Scene.from_file(scene_path("cornell_box.json"), random_spheres_seed=1)
    rays = scene_rays(sc, 4096, seed=11)
    sc.set_counters(True)
    sc.reset_stats()
    sc.closest_hit(rays, mode=rt.RT_ISECT_BRUTE, want=("index",))
    plain = sc.stats().march_steps
    sc.reset_stats()
    sc.closest_hit(rays, mode=rt.RT_ISECT_FAST, want=("index",))
    skip = sc.stats().march_steps
    assert plain > 500_000 and skip * 10 < plain, (plain, skip)


# ---- conservative cull (rt_cull.cuh) ---------------------------------------------------------------
def grazing_rays(sc, n_shapes, seed):
    """rays that pass at distance r*(1 + eps) from the centre of small spheres / cubes, from near and
    far: the inputs on which an over-eager cull would drop a real hit"""
    rng = np.random.default_rng(seed)
    d = sc.desc()
    dirm = np.ctypeslib.as_array(d.direct, shape=(d.n_shapes, 12))
    kinds = sc.shape_kinds()
    cand = np.where((kinds == 0) | (kinds == 1))[0]
    pick = cand[rng.integers(0, len(cand), n_shapes)]
    rays = []
    for i in pick:
        C = dirm[i, [3, 7, 11]]
        r = np.abs(dirm[i, [0, 5, 10]]).max() * (math.sqrt(3.0) if kinds[i] == 1 else 1.0)
        for eps in (0.0, 1e-14, -1e-14, 1e-9, -1e-9, 1e-6, -1e-6, 1e-3, -1e-3, 0.02, 0.045, 0.06, 0.2, -0.3):
            for dist in (0.0, 0.5 * r, 3.0 * r, 50.0, 1e3, 1e5):
                u = rng.normal(size=3); u /= np.linalg.norm(u)
                w = np.cross(u, rng.normal(size=3)); w /= np.linalg.norm(w)
                o = C + w * r * (1.0 + eps) - u * dist
                rays.append(np.concatenate([o, u]))
                rays.append(np.concatenate([o, -u]))      # the same line walked the other way (behind test)
    r = np.array(rays)
    return rt.make_rays(r[:, :3], r[:, 3:])


@pytest.mark.parametrize("name", ["spheres.json", "cornell_box.json", "detached_materials.json", "dupin.json",
                                  "cube_test.json", "light_source.json"])
def test_cull_is_conservative_on_fixture_scenes(name):
    """RT_ISECT_VERIFY: FAST == BRUTE on every ray and no culled (ray, shape) pair hits in the exact test"""
    sc = rt.Scene.from_file(scene_path(name), random_spheres_seed=1)
    rays = np.concatenate([scene_rays(sc, 1 << 13, seed=5), grazing_rays(sc, 40, seed=6)])
    sc.reset_stats()
    got = sc.closest_hit(rays, mode=rt.RT_ISECT_VERIFY)
    st = sc.stats()
    assert st.verify_rays == 0, f"{st.verify_rays} rays differ between the culled and the literal loop"
    assert st.verify_false_culls == 0, f"{st.verify_false_culls} culled pairs are hits in the exact test"
    osc = po.OracleScene(sc.desc())
    _compare(got, osc.intersect_batch(rays), len(rays), rt.RT_ISECT_VERIFY)
    _compare(sc.closest_hit(rays, mode=rt.RT_ISECT_FAST), osc.intersect_batch(rays), len(rays), rt.RT_ISECT_FAST)


def test_cull_far_and_ill_conditioned_shapes():
    """small shapes far from the origin (FP32 cancellation in C - o), flattened / rotated shapes
    (condition number > 30: never culled), huge shapes, rays from 1e6 away"""
    import json
    scene = json.loads(json.dumps(TRIO))
    T = lambda t, r, s: {"translate": t, "rotate": r, "scale": s}
    scene["shapes"] = [
        {"type": "Sphere", "name": "far small", "material": "M", "transform": T([1e6, 2e6, -3e6], [0, 0, 0], [0.2] * 3)},
        {"type": "Sphere", "name": "flat", "material": "M", "transform": T([0, 5, 0], [0, -100, 0], [10, 0.1, 10])},
        {"type": "Cube", "name": "rot cube", "material": "M", "transform": T([4, 1, 2], [30, 45, 60], [1, 2, 3])},
        {"type": "Cube", "name": "far cube", "material": "M", "transform": T([-5e5, 10, 7e5], [10, 20, 30], [0.5] * 3)},
        {"type": "Sphere", "name": "sun", "material": "M", "transform": T([0, 1.476e11, 0], [0, 0, 0], [7e8] * 3)},
        {"type": "Sphere", "name": "ellipsoid", "material": "M", "transform": T([-3, 0, 4], [15, 25, 35], [1, 2, 4])},
        {"type": "Sphere", "name": "tiny", "material": "M", "transform": T([0.5, 0.5, 0.5], [0, 0, 0], [1e-4] * 3)},
    ]
    sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
    rng = np.random.default_rng(12)
    d = sc.desc()
    dirm = np.ctypeslib.as_array(d.direct, shape=(d.n_shapes, 12))
    rays = [grazing_rays(sc, 60, seed=13)]
    for i in range(d.n_shapes):   # rays aimed at / near each shape from random distances
        C = dirm[i, [3, 7, 11]]
        r = np.abs(dirm[i, [0, 5, 10]]).max()
        for dist in (2.0, 1e2, 1e4, 1e6):
            u = rng.normal(size=(300, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
            o = C + u * (r + dist)
            tgt = C + rng.normal(size=(300, 3)) * r * 0.8
            rays.append(rt.make_rays(o, tgt - o))
    rays = np.concatenate(rays)
    sc.reset_stats()
    got = sc.closest_hit(rays, mode=rt.RT_ISECT_VERIFY)
    st = sc.stats()
    assert st.verify_rays == 0 and st.verify_false_culls == 0, (st.verify_rays, st.verify_false_culls)
    want = po.OracleScene(sc.desc()).intersect_batch(rays)
    _compare(got, want, len(rays), rt.RT_ISECT_VERIFY)
    _compare(sc.closest_hit(rays, mode=rt.RT_ISECT_FAST), want, len(rays), rt.RT_ISECT_FAST)
    assert len(set(np.unique(want["index"])) - {-1}) >= 6   # (nearly) every shape is hit by some ray


def test_cull_tree_removes_most_tests():
    """on cornell_box the 481 small spheres sit in one corner under one root ball: a segment costs a few
    dozen FP32 ball tests (flat list + root + the groups / leaves its line touches) instead of ~490, and
    the exact FP64 test runs for a few percent of the (segment, shape) pairs"""
    sc = rt.Scene.from_file(scene_path("cornell_box.json"), random_spheres_seed=1)
    rays = scene_rays(sc, 1 << 14, seed=21)
    sc.set_counters(True)
    sc.reset_stats()
    sc.closest_hit(rays, mode=rt.RT_ISECT_FAST, want=("index",))
    st = sc.stats()
    sc.set_counters(False)
    pairs = len(rays) * (sc.shape_count - 1)
    assert st.cull_tests >= len(rays) * 8           # the never-culled shapes and the root are always looked at
    assert st.cull_tests * 4 < pairs, (st.cull_tests, pairs)
    assert st.shape_tests * 20 < pairs, (st.shape_tests, pairs)


@pytest.mark.parametrize("n_spheres,spread,offset", [(40, 3.0, 0.0), (700, 30.0, 0.0), (1500, 8.0, 0.0),
                                                     (300, 4.0, 2e5)])
def test_cull_tree_sizes(n_spheres, spread, offset):
    """scenes below the tree threshold (flat list only), with more than one root (> 512 tree shapes) and
    with heavily overlapping groups, and a cluster 3.7e5 away from the origin (FP32 cancellation in the node
    tests): FAST == BRUTE == oracle bit for bit, and RT_ISECT_VERIFY finds no false cull at any level"""
    rng = np.random.default_rng(n_spheres)
    shapes = []
    off = np.array([offset, 1.5 * offset, -2.5 * offset])
    for k in range(n_spheres):
        c = rng.uniform(-spread, spread, 3) + off
        r = float(rng.uniform(0.05, 0.6))
        kind = "Cube" if k % 7 == 0 else "Sphere"
        shapes.append({"type": kind, "name": f"s{k}", "material": "M",
                       "transform": {"translate": c.tolist(), "rotate": rng.uniform(-90, 90, 3).tolist(),
                                     "scale": [r, r * float(rng.uniform(0.5, 1.5)), r]}})
    shapes.append({"type": "Sphere", "name": "ground", "material": "M",
                   "transform": {"translate": [0, -1000 - spread, 0], "rotate": [0, 0, 0], "scale": [1000, 1000, 1000]}})
    data = dict(TRIO)
    data["shapes"] = shapes
    import json
    sc = rt.Scene.from_json(json.dumps(data), add_random_spheres=False)
    o = rng.uniform(-1.5 * spread, 1.5 * spread, (6000, 3)) + off
    o[::3] = rng.uniform(-1.5 * spread, 1.5 * spread, (2000, 3))   # a third of the rays start near the world origin
    tgt = rng.uniform(-spread, spread, (6000, 3)) + off
    rays = rt.make_rays(o, tgt - o)
    # rays no pre-test can judge (NaN / zero / infinite components pass every ball test, padding entries
    # included): they must come out like the literal loop's
    odd = np.array([[0, 0, 0, math.nan, 0, 1], [1, 2, 3, 0, 0, 0], [math.nan, 0, 0, 0, 0, 1],
                    [0, 0, 0, math.inf, 0, 0], [math.inf, 1, 1, 0, 1, 0]], dtype=np.float64)
    rays = np.concatenate([rays, odd + np.concatenate([off, [0, 0, 0]])])
    want = check_parity(sc, rays)
    assert (want["index"] >= 0).mean() > (0.3 if offset == 0.0 else 0.05)
    sc.reset_stats()
    sc.closest_hit(rays, mode=rt.RT_ISECT_VERIFY, want=("index",))
    st = sc.stats()
    assert st.verify_rays == 0 and st.verify_false_culls == 0


@pytest.mark.parametrize("n_shapes,extent,n_rays", [(10000, 24.0, 8192), (100000, 52.0, 4096)])
def test_arbitrary_depth_cull_tree_on_generated_scenes(n_shapes, extent, n_rays):
    """the cull tree generalised to arbitrary depth (levels of bounding balls above the roots, walked with
    range skipping) on generated scenes of 10^4 and 10^5 shapes (+ Rectangles on the flat list + a large ground
    sphere): FAST == BRUTE == oracle bit for bit, RT_ISECT_VERIFY clean at every level, and the work per ray grows
    like log n (counters), not like n"""
    from test_cull_cpu import big_scene, box_rays
    sc = big_scene(n_shapes, seed=n_shapes, extent=extent)
    rays = box_rays(n_rays, seed=7, extent=extent)
    odd = np.array([[0, 0, 0, math.nan, 0, 1], [1, 2, 3, 0, 0, 0], [0, 0, 0, math.inf, 0, 0]], dtype=np.float64)
    rays = np.concatenate([rays, odd])
    want = check_parity(sc, rays)
    assert (want["index"] >= 0).mean() > 0.5
    sub = rays[:512]
    sc.reset_stats()
    sc.closest_hit(sub, mode=rt.RT_ISECT_VERIFY, want=("index",))
    st = sc.stats()
    assert st.verify_rays == 0 and st.verify_false_culls == 0
    sc.set_counters(True)
    sc.reset_stats()
    sc.closest_hit(rays[:n_rays], mode=rt.RT_ISECT_FAST, want=("index",))   # (without the NaN / inf rays: nothing culls those)
    st = sc.stats()
    sc.set_counters(False)
    per_ray_cull, per_ray_exact = st.cull_tests / n_rays, st.shape_tests / n_rays
    # 10^4 shapes: 20 roots under 1 upper node; 10^5: 196 roots under 7.  Measured (profiles/r2k_big_scenes.md):
    # 684 / 1986 ball tests and 8.3 / 9.6 exact tests per ray -- a ray through a box this dense crosses a few dozen
    # groups of 16 -- against 10 007 / 100 007 exact tests of the literal loop
    assert per_ray_cull < 0.05 * n_shapes + 600 and per_ray_exact < 25, (per_ray_cull, per_ray_exact)
