"""Per-kernel-class device time and work counters of one frame of each BASELINE configuration (reduced spp;
Mpaths/s does not depend on spp).  Single lane, CUDA event pairs around every launch (rt_set_kernel_timing).

  python tools/kernel_breakdown.py [--cfg 3 4a 4b 5] [--scale 16]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", nargs="*", default=["1", "3", "4a", "4b", "5"])
ap.add_argument("--scale", type=int, default=16, help="divide the samples of cfg 3-5 by this")
a = ap.parse_args()


def scene(name):
    return rt.Scene.from_file(os.path.join(ROOT, "scenes", name), random_spheres_seed=1)


def config(cfg):
    if cfg == "1":
        sc = scene("spheres.json")
        return sc, sc.camera(), 640, 480, 16
    if cfg == "3":
        sc = scene("cornell_box.json")
        return sc, sc.camera(), 1024, 1024, max(256 // a.scale, 1)
    if cfg == "4a":
        sc = scene("detached_materials.json")
        return sc, sc.camera(), 1920, 1080, max(256 // a.scale, 1)
    if cfg == "4b":
        sc = scene("detached_materials.json")
        sc.assign_material(1, "EarthMap")
        sc.assign_material(2, "Glass")
        sc.assign_material(5, "Lambertian01")
        sc.assign_material(6, "WhiteMirror")
        c0 = sc.camera()
        pos = np.array(c0.position.tuple())
        return sc, rt.camera_new(pos, -pos, (0, 1, 0), 1.0, c0.fov_rad), 1920, 1080, max(256 // a.scale, 1)
    if cfg == "5":
        sc = scene("dupin.json")
        return sc, sc.camera(), 3840, 2160, max(1024 // (a.scale * 4), 1)
    raise SystemExit(f"unknown cfg {cfg}")


print("| cfg | frame | Mpaths/s (2 lanes) | 1-lane ms: raygen | extend | march | shade | resolve | segments/path | "
      "exact tests/seg | cull tests/seg | marched rays/path | evals/marched ray | rays > 2048 evals | max evals | lit0/lit+/jumps/hops/hull misses/hop misses/empty plans at level 0/at refinement levels per marched ray |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for cfg in a.cfg:
    sc, cam, w, h, spp = config(cfg)
    ds = sc.device_scene(0)
    p = api.render_params(w, h, spp, 8, seed=1)
    for _ in range(2):  # warm-up + plain two-lane frame
        api.render_start(ds, cam, p)
        api.render_wait(ds, None)
    plain_ms = sc.stats().last_frame_ms
    sc.set_kernel_timing(True)
    sc.reset_stats()
    api.render_start(ds, cam, p)
    api.render_wait(ds, None)
    t = sc.stats()
    sc.set_kernel_timing(False)
    sc.set_counters(True)
    sc.reset_stats()
    api.render_start(ds, cam, p)
    api.render_wait(ds, None)
    c = sc.stats()
    sc.set_counters(False)
    paths = w * h * spp
    print(f"| {cfg} | {w}x{h}x{spp} | {paths / plain_ms / 1e3:.1f} | {t.ms_raygen:.2f} | {t.ms_extend:.2f} | "
          f"{t.ms_march:.2f} | {t.ms_shade:.2f} | {t.ms_resolve:.2f} | {c.segments / paths:.2f} | "
          f"{c.shape_tests / max(c.segments, 1):.2f} | {c.cull_tests / max(c.segments, 1):.1f} | "
          f"{c.march_rays / paths:.3f} | {c.march_steps / max(c.march_rays, 1):.1f} | {c.march_long_rays} | {c.march_max_evals} | "
          f"{' / '.join(f'{c.march_prof[k] / max(c.march_rays, 1):.2f}' for k in range(8))} |", flush=True)
