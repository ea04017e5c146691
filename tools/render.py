"""The bins' render-and-save (src/bin/main_raylib.rs: RendererState::render + the 'S' key) without the window:

  python tools/render.py scenes/cornell_box.json out.png --size 1024 1024 --spp 256 [--depth 8] [--seed 1]

Scene::from_json -> GpuRenderer (Renderer trait: start_rendering / render_step until done) ->
rt_tonemap_rgba8 (sqrt, clamp(0, 0.999) * 256, on the device) -> PNG."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rs_pathtracing_b200 as rt

ap = argparse.ArgumentParser()
ap.add_argument("scene")
ap.add_argument("out")
ap.add_argument("--size", type=int, nargs=2, default=[640, 480])
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--depth", type=int, default=50)   # the bins construct their renderer with depth 50
ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()
sc = rt.Scene.from_file(a.scene, random_spheres_seed=a.seed)
cam = sc.camera()
r = rt.GpuRenderer(sc, 12, a.depth, seed=a.seed)
w, h = a.size
buf = np.zeros((h, w, 3))
t0 = time.perf_counter()
r.start_rendering(cam, rt.ImageParams(w, h), a.spp)
while not r.render_step(buf):
    time.sleep(0.001)
dt = time.perf_counter() - t0
rt.save_png(a.out, rt.tonemap_rgba8(sc, buf))
print(f"{a.scene}: {w}x{h}x{a.spp} in {dt * 1e3:.1f} ms ({w * h * a.spp / dt / 1e6:.1f} Mpaths/s) -> {a.out}")
