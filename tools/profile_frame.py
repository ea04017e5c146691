"""One frame of a scene at a chosen size — the short command ncu wraps (profiles/README.md)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="cornell_box.json")
ap.add_argument("--size", type=int, nargs=2, default=[256, 256])
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--depth", type=int, default=8)
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--variant", default="", help="4b: detached_materials.json with every material / texture kind "
                                              "assigned and the camera looking at the origin (tools/run_configs.py)")
a = ap.parse_args()
if a.variant == "4b":
    a.scene = "detached_materials.json"
sc = rt.Scene.from_file(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes", a.scene), 1)
cam = sc.camera()
if a.variant == "4b":
    import numpy as np
    sc.assign_material(1, "EarthMap")
    sc.assign_material(2, "Glass")
    sc.assign_material(5, "Lambertian01")
    sc.assign_material(6, "WhiteMirror")
    pos = np.array(cam.position.tuple())
    cam = rt.camera_new(pos, -pos, (0, 1, 0), 1.0, cam.fov_rad)
ds = sc.device_scene(0)
for _ in range(a.frames):
    api.render_start(ds, cam, api.render_params(a.size[0], a.size[1], a.spp, a.depth, seed=1))
    api.render_wait(ds, None)
st = sc.stats()
print(f"{a.scene} {a.size[0]}x{a.size[1]}x{a.spp}: {st.last_frame_ms:.2f} ms/frame, "
      f"{a.size[0]*a.size[1]*a.spp/st.last_frame_ms/1e3:.2f} Mpaths/s, launches {st.kernel_launches}")
