"""Sweep k_march scheduling thresholds (RT_B200_MARCH_TUNE) on two scenes; prints march ms."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import api
out = []
for name, w, h, spp in [("cornell_box.json", 1024, 1024, 4), ("dupin.json", 960, 540, 8), ("spheres.json", 640, 480, 16)]:
    sc = rt.Scene.from_file(os.path.join(%r, "scenes", name), 1)
    cam = sc.camera(); ds = sc.device_scene(0)
    for rep in range(3):
        sc.set_kernel_timing(rep == 2); sc.reset_stats()
        api.render_start(ds, cam, api.render_params(w, h, spp, 8, seed=1)); api.render_wait(ds, None)
    st = sc.stats()
    out.append(f"{name.split('.')[0]} march {st.ms_march:.2f} total {st.last_frame_ms:.2f}")
print(os.environ.get("RT_B200_MARCH_TUNE", "default"), "|", " | ".join(out))
''' % (ROOT, ROOT)
for tune in sys.argv[1:]:
    env = dict(os.environ, RT_B200_MARCH_TUNE=tune)
    subprocess.run([sys.executable, "-c", code], env=env)
