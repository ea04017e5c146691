"""Nearest-hit throughput against scene size: generated scenes of 10^3 .. 10^5 shapes, 256 Ki rays, the
literal loop (RT_ISECT_BRUTE) against the arbitrary-depth cull tree (RT_ISECT_FAST); prints a markdown table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rs_pathtracing_b200 as rt
from test_cull_cpu import big_scene, box_rays

print("| shapes | box half-width | ball tests / ray | exact tests / ray | FAST ms (256 Ki rays) | Mrays/s | BRUTE ms (16 Ki rays) | Mrays/s | FAST == BRUTE |")
print("|---|---|---|---|---|---|---|---|---|")
for n, ext in ((1000, 11.0), (10000, 24.0), (100000, 52.0)):
    sc = big_scene(n, seed=n, extent=ext)
    rays = box_rays(1 << 18, seed=7, extent=ext)
    for _ in range(2):
        fast = sc.closest_hit(rays, mode=rt.RT_ISECT_FAST, want=("index", "t"))
    ms_fast = sc.stats().last_intersect_ms
    small = rays[: 1 << 14]
    for _ in range(2):
        brute = sc.closest_hit(small, mode=rt.RT_ISECT_BRUTE, want=("index", "t"))
    ms_brute = sc.stats().last_intersect_ms
    same = np.array_equal(fast["index"][: 1 << 14], brute["index"]) and np.array_equal(fast["t"][: 1 << 14], brute["t"], equal_nan=True)
    sc.set_counters(True); sc.reset_stats()
    sc.closest_hit(small, mode=rt.RT_ISECT_FAST, want=("index",))
    st = sc.stats(); sc.set_counters(False)
    print(f"| {n} (+6 Rectangles, ground) | {ext} | {st.cull_tests / len(small):.0f} | {st.shape_tests / len(small):.1f} | "
          f"{ms_fast:.2f} | {len(rays) / ms_fast / 1e3:.1f} | {ms_brute:.2f} | {len(small) / ms_brute / 1e3:.2f} | {same} |", flush=True)
