"""cfg 2 on the cornell shape list: 1 Mi rays through rt_intersect_batch, literal loop and cull tree + exact skip"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rs_pathtracing_b200 as rt
from test_gpu_intersect import scene_rays
s = rt.Scene.from_file(os.path.join(ROOT, "scenes", "cornell_box.json"), random_spheres_seed=1)
r = scene_rays(s, 1 << 20, seed=9)
for mode, name in ((rt.RT_ISECT_FAST, "cull tree + exact skip"),):
    for _ in range(3):
        s.closest_hit(r, mode=mode, want=("index", "t", "normal"))
    ms = s.stats().last_intersect_ms
    print(f"{name}: {ms:.3f} ms, {len(r) / ms / 1e3:.1f} Mrays/s")
