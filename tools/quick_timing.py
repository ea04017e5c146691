"""Ad-hoc timing probe (not the bench): frame time / Mpaths/s at a few sizes, FP64/FP32 peaks."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import api

print("peaks TFLOP/s (fp64, fp32):", rt.measure_peaks())
for name, w, h, spp in [("spheres.json", 640, 480, 16), ("cornell_box.json", 256, 256, 16), ("cornell_box.json", 1024, 1024, 4),
                        ("detached_materials.json", 480, 270, 16), ("dupin.json", 480, 270, 16)]:
    sc = rt.Scene.from_file(f"scenes/{name}", 1)
    cam = sc.camera()
    ds = sc.device_scene(0)
    for rep in range(2):
        sc.set_counters(rep == 1)
        sc.set_kernel_timing(rep == 0)
        sc.reset_stats()
        p = api.render_params(w, h, spp, 8, seed=1)
        t0 = time.time()
        api.render_start(ds, cam, p)
        api.render_wait(ds, None)
        dt = time.time() - t0
        st = sc.stats()
        print(f"{name} {w}x{h}x{spp} counters={rep}: wall {dt*1e3:.1f} ms, device {st.last_frame_ms:.1f} ms, "
              f"{w*h*spp/st.last_frame_ms/1e3:.2f} Mpaths/s, launches {st.kernel_launches}, seg {st.segments}, "
              f"exact {st.shape_tests}, cull {st.cull_tests}, march_steps {st.march_steps}, march_rays {st.march_rays}, "
              f"long {st.march_long_rays}, max {st.march_max_evals} | ms raygen {st.ms_raygen:.2f} extend {st.ms_extend:.2f} "
              f"march {st.ms_march:.2f} shade {st.ms_shade:.2f} resolve {st.ms_resolve:.2f} | prof lit0 {st.march_prof[0]} lit+ {st.march_prof[1]} jumps {st.march_prof[2]} hops {st.march_prof[3]}")
