"""How much of the marching cost is SIMT divergence?  rt_intersect_batch (FAST) on a Heart-only scene:
(a) 2^20 different rays, (b) every ray repeated 32 times in a row (each warp marches one ray 32-fold:
perfectly convergent), (c) rays sorted by number of evaluations.  Device time from rt_stats."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import rs_pathtracing_b200 as rt
from test_gpu_intersect import TRIO, bench_rays

scene = json.loads(json.dumps(TRIO))
scene["shapes"] = [scene["shapes"][2]]
scene["shapes"][0]["transform"]["scale"] = [2, 2, 2]
sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
n = 1 << 20
rays = bench_rays(n, seed=3, target_radius=2.0)
def run(r, label):
    for _ in range(2):
        sc.closest_hit(r, mode=rt.RT_ISECT_FAST, want=("index",))
    ms = sc.stats().last_intersect_ms
    print(f"{label:40s} {ms:8.3f} ms  {len(r) / ms / 1e3:8.1f} Mrays/s")
    return ms
run(rays, "different rays")
rep = np.repeat(rays[: n // 32], 32, axis=0)
run(rep, "each ray x32 (convergent warps)")
perm = np.random.default_rng(0).permutation(n)
run(rep[perm], "the same multiset, shuffled")
