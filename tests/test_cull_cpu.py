"""Conservativeness of the FP32 cull (tree and marching-bound pre-test) on the CPU: rt_cull_reached emulates
k_extend's decisions with the same once-rounded FP32 operations (host build of csrc/rt_cull.cuh); the
oracle tests every (ray, shape) pair exactly with max_t = +inf.  Contract: oracle hits the shape  =>  the
cull hands it to the exact test.  (The GPU suite checks the same thing on the device through
RT_ISECT_VERIFY; this one needs no GPU.)"""
import ctypes as C
import json

import numpy as np
import pytest

import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import _ffi
from oracle import pyoracle as po

from conftest import scene_path
from test_gpu_intersect import TRIO, grazing_rays, scene_rays


def reached(sc, rays):
    d = sc.desc()
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    out = np.zeros((len(rays), d.n_shapes), np.uint8)
    rc = _ffi.core().rt_cull_reached(C.byref(d), rays.ctypes.data_as(C.c_void_p), len(rays), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def check(sc, rays, min_hit_fraction=0.0):
    osc = po.OracleScene(sc.desc())
    n = sc.shape_count
    got = reached(sc, rays)
    hits = np.stack([osc.shape_hits(r, n) for r in rays])
    missed = (hits == 1) & (got == 0)
    assert not missed.any(), f"{missed.sum()} (ray, shape) pairs are culled although the exact test hits: {np.argwhere(missed)[:5]}"
    assert (hits.sum(axis=1) > 0).mean() >= min_hit_fraction
    return got, hits


@pytest.mark.parametrize("name", ["spheres.json", "cornell_box.json", "detached_materials.json", "dupin.json",
                                  "light_source.json"])
def test_no_hit_is_culled_on_fixture_scenes(name):
    sc = rt.Scene.from_file(scene_path(name), random_spheres_seed=1)
    rays = np.concatenate([scene_rays(sc, 1500, seed=5), grazing_rays(sc, 12, seed=6)])
    got, hits = check(sc, rays, min_hit_fraction=0.5)
    # and the cull does cull: a ray looks at a small fraction of the ~490 shapes
    assert got.sum(axis=1).mean() < 0.08 * sc.shape_count


def test_no_hit_is_culled_for_far_rotated_and_anisotropic_shapes():
    rng = np.random.default_rng(4)
    shapes = []
    for k in range(300):
        c = rng.uniform(-6, 6, 3) + (np.array([2e5, 3e5, -5e5]) if k % 2 else 0.0)
        r = float(rng.uniform(0.05, 0.6))
        shapes.append({"type": "Cube" if k % 4 == 0 else "Sphere", "name": f"s{k}", "material": "M",
                       "transform": {"translate": c.tolist(), "rotate": rng.uniform(-90, 90, 3).tolist(),
                                     "scale": [r, r * float(rng.uniform(0.5, 2.0)), r * float(rng.uniform(0.5, 2.0))]}})
    scene = dict(TRIO)
    scene["shapes"] = shapes
    sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
    off = np.array([2e5, 3e5, -5e5])
    o = rng.uniform(-9, 9, (1200, 3))
    o[::2] += off
    tgt = rng.uniform(-6, 6, (1200, 3))
    tgt[::3] += off
    rays = rt.make_rays(o, tgt - o)
    check(sc, rays, min_hit_fraction=0.3)


def test_tangent_rays_reach_their_sphere():
    """lines tangent to a sphere within +-1e-9 relative (the D == 0 neighbourhood), from in front and from
    behind: the reference accepts D == 0 without a range check, so the sphere must always be reached"""
    scene = dict(TRIO)
    scene["shapes"] = [{"type": "Sphere", "name": f"s{k}", "material": "M",
                        "transform": {"translate": [3.0 * k - 30, 0.5 * k, 2.0], "rotate": [0, 0, 0],
                                      "scale": [0.2 + 0.05 * k] * 3}} for k in range(80)]
    sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
    rng = np.random.default_rng(8)
    rays = []
    for k in range(80):
        c, r = np.array([3.0 * k - 30, 0.5 * k, 2.0]), 0.2 + 0.05 * k
        for _ in range(6):
            u = rng.normal(size=3); u /= np.linalg.norm(u)
            w = np.cross(u, rng.normal(size=3)); w /= np.linalg.norm(w)
            touch = c + r * (1.0 + rng.uniform(-1e-9, 1e-9)) * w
            for back in (-1.0, 1.0):
                o = touch - back * u * rng.uniform(5, 40)
                rays.append(np.concatenate([o, u]))
    rays = np.array(rays)
    got = reached(sc, rays)
    idx = np.repeat(np.arange(80), 12)
    assert got[np.arange(len(rays)), idx].all()
    check(sc, rays)


def big_scene(n_shapes, seed, extent):
    """a generated scene of n_shapes small Spheres / Cubes in a box of half-width `extent` plus a few Rectangles
    (flat list: never culled) and one large ground sphere -- the arbitrary-depth case of the cull tree"""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (n_shapes, 3))
    r = rng.uniform(0.05, 0.4, n_shapes)
    rot = rng.uniform(-90, 90, (n_shapes, 3))
    shapes = [{"type": "Cube" if k % 5 == 0 else "Sphere", "name": "s", "material": "M",
               "transform": {"translate": c[k].tolist(), "rotate": rot[k].tolist() if k % 5 == 0 else [0, 0, 0],
                             "scale": [float(r[k])] * 3}} for k in range(n_shapes)]
    for k in range(6):
        shapes.append({"type": "Rectangle", "x0": -1.0, "y0": -1.0, "x1": 1.0, "y1": 1.0, "material": "M",
                       "transform": {"translate": rng.uniform(-extent, extent, 3).tolist(),
                                     "rotate": rng.uniform(-90, 90, 3).tolist(), "scale": [extent / 4] * 3}})
    shapes.append({"type": "Sphere", "name": "ground", "material": "M",
                   "transform": {"translate": [0, -extent - 1000.0, 0], "rotate": [0, 0, 0], "scale": [1000.0] * 3}})
    scene = dict(TRIO)
    scene["shapes"] = shapes
    return rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)


def box_rays(n, seed, extent):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-1.5 * extent, 1.5 * extent, (n, 3))
    tgt = rng.uniform(-extent, extent, (n, 3))
    return rt.make_rays(o, tgt - o)


def test_arbitrary_depth_tree_is_conservative_and_encloses():
    """20 000 shapes: 40 roots under two upper-level nodes.  Enclosure of every node (rt_cull_tree_check, FP64) and
    "oracle hits => reached" pair by pair (rt_cull_reached walks the levels like the device)"""
    sc = big_scene(20000, seed=3, extent=30.0)
    d = sc.desc()
    out = [C.c_uint32() for _ in range(4)]
    worst = C.c_double()
    assert _ffi.core().rt_cull_tree_check(C.byref(d), *[C.byref(x) for x in out], C.byref(worst)) == 0
    n_roots, n_groups, n_tree, n_flat = (x.value for x in out)
    assert n_tree == 20000 and n_flat == 7 and n_roots == (n_groups + 31) // 32 and n_roots > 32
    assert worst.value <= 1e-7, worst.value
    got, hits = check(sc, box_rays(160, seed=4, extent=30.0), min_hit_fraction=0.5)
    assert got.sum(axis=1).mean() < 0.01 * sc.shape_count      # a ray looks at < 1 % of the shapes
