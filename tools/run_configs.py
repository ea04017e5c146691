"""Runs the five configurations BASELINE.json names on one GPU and prints a markdown table
(profiles/README.md quotes it).  Frames go through the public Renderer API (host buffer out).

  python tools/run_configs.py [--quick]      # --quick: 1/16 of the samples of cfg 3-5
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
args = ap.parse_args()
div = 16 if args.quick else 1
rows = []


def render(name, sc, cam, w, h, spp, depth=8, frames=2):
    r = rt.GpuRenderer(sc, 12, depth, seed=5)
    best = None
    for _ in range(frames):
        t0 = time.perf_counter()
        frame = r.render(cam, w, h, spp)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = sc.stats()
    assert np.isfinite(frame).all()
    rows.append((name, f"{w}x{h}, {spp} spp, depth {depth}", f"{w * h * spp / 1e6:.1f} M", f"{best * 1e3:.1f}",
                 f"{st.last_frame_ms:.1f}", f"{w * h * spp / best / 1e6:.1f}", f"{frame.mean():.4f}"))


def scene(name):
    return rt.Scene.from_file(os.path.join(ROOT, "scenes", name), random_spheres_seed=1)


# cfg 1
sc = scene("spheres.json")
render("1 spheres.json", sc, sc.camera(), 640, 480, 16)
# cfg 2: 1 Mi rays against the bench trio and against the flattened cornell scene
from test_gpu_intersect import TRIO, bench_rays, scene_rays
trio = rt.Scene.from_json(json.dumps(TRIO), add_random_spheres=False)
rays = bench_rays(1 << 20, seed=42)
for label, s, r in (("2 bench trio {Sphere, Cube, Heart}", trio, rays),
                    ("2 cornell_box shape list (492)", scene("cornell_box.json"), None)):
    if r is None:
        r = scene_rays(s, 1 << 20, seed=9)
    for mode, mname in ((rt.RT_ISECT_BRUTE, "literal loop"), (rt.RT_ISECT_FAST, "cull tree + exact skip")):
        for _ in range(2):
            t0 = time.perf_counter()
            s.closest_hit(r, mode=mode, want=("index", "t", "normal"))
            dt = time.perf_counter() - t0
        ms = s.stats().last_intersect_ms
        rows.append((label, f"1 Mi rays, {mname}", "1.05 M rays", f"{dt * 1e3:.1f}", f"{ms:.2f}",
                     f"{len(r) / ms / 1e3:.1f} Mrays/s (kernel)", "-"))
# cfg 3
sc = scene("cornell_box.json")
render("3 cornell_box.json", sc, sc.camera(), 1024, 1024, 256 // div)
# cfg 4a / 4b
sc = scene("detached_materials.json")
render("4a detached_materials.json as shipped", sc, sc.camera(), 1920, 1080, 256 // div)
sc = scene("detached_materials.json")   # materials can only be reassigned before the scene is on the device
sc.assign_material(1, "EarthMap")
sc.assign_material(2, "Glass")
sc.assign_material(5, "Lambertian01")
sc.assign_material(6, "WhiteMirror")
c0 = sc.camera()
pos = np.array(c0.position.tuple())
render("4b detached_materials.json, look-at-origin, all material/texture kinds", sc,
       rt.camera_new(pos, -pos, (0, 1, 0), 1.0, c0.fov_rad), 1920, 1080, 256 // div)
# cfg 5
sc = scene("dupin.json")
render("5 dupin.json", sc, sc.camera(), 3840, 2160, 1024 // div, frames=1)

print("| cfg | workload | paths | wall ms (API, host frame out) | device ms | Mpaths/s (wall) | frame mean |")
print("|---|---|---|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")
