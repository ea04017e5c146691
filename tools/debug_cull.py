"""Debug helper: where do FAST and BRUTE disagree?  (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import rs_pathtracing_b200 as rt
from test_gpu_intersect import grazing_rays, scene_rays
from conftest import scene_path

name = sys.argv[1] if len(sys.argv) > 1 else "spheres.json"
sc = rt.Scene.from_file(scene_path(name), random_spheres_seed=1)
rays = np.concatenate([scene_rays(sc, 1 << 13, seed=5), grazing_rays(sc, 40, seed=6)])
sc.reset_stats()
v = sc.closest_hit(rays, mode=rt.RT_ISECT_VERIFY)
st = sc.stats()
print("verify_rays", st.verify_rays, "false_culls", st.verify_false_culls, "n", len(rays))
b = sc.closest_hit(rays, mode=rt.RT_ISECT_BRUTE)
f = sc.closest_hit(rays, mode=rt.RT_ISECT_FAST)
bad = np.where((b["index"] != f["index"]) | ((b["index"] >= 0) & (b["t"] != f["t"])))[0]
print("differing rays:", len(bad), "first in grazing part:", (bad >= (1 << 13)).sum())
kinds = sc.shape_kinds()
for i in bad[:25]:
    bi, fi = b["index"][i], f["index"][i]
    print(i, "ray", rays[i], "brute", bi, (kinds[bi] if bi >= 0 else -1), repr(b["t"][i]), "fast", fi, (kinds[fi] if fi >= 0 else -1), repr(f["t"][i]))
