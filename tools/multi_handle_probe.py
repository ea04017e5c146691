"""One process, one handle, every GPU of the box (rt_scene_create_multi through GpuRenderer(devices=...)): frame time of
the bench frame against the single-device handle, and equality of the two frames."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rs_pathtracing_b200 as rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = rt.device_count()
w = h = 1024
spp = 64
res = {}
for devices in ([0], list(range(n))):
    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", "cornell_box.json"), random_spheres_seed=1)
    r = rt.GpuRenderer(sc, 12, 8, seed=3, devices=devices)
    cam = sc.camera()
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        frame = r.render(cam, w, h, spp)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    res[len(devices)] = (best, frame.copy())
    print(f"{len(devices)} device(s) behind one handle: {best * 1e3:.1f} ms / frame end to end (host frame out), "
          f"{w * h * spp / best / 1e6:.1f} Mpaths/s")
if n > 1:
    print("frames identical:", np.array_equal(res[1][1], res[n][1]), f"speed-up {res[1][0] / res[n][0]:.2f}x on {n} GPUs")
