"""Row-ordered (unsharded: one 'tile' per image row) against tile-ordered (32x32 tiles: the two shards of a
2-shard frame rendered one after the other on ONE GPU) batches: device ms of the same total work."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import api
for name, w, h, spp in (("cornell_box.json", 1024, 1024, 256), ("dupin.json", 3840, 2160, 128), ("spheres.json", 1920, 1080, 64)):
    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", name), random_spheres_seed=1)
    cam = sc.camera(); ds = sc.device_scene(0)
    def ms(shards, s):
        api.render_start(ds, cam, api.render_params(w, h, spp, 8, 2024, shards, s, tile=32)); api.render_wait(ds, None)
        return sc.stats().last_frame_ms
    ms(1, 0)
    rows = min(ms(1, 0) for _ in range(2))
    tiles = min(ms(2, 0) + ms(2, 1) for _ in range(2))
    print(f"{name} {w}x{h}x{spp}: row-ordered {rows:.1f} ms, tile-ordered (2 shards in sequence) {tiles:.1f} ms, ratio {rows / tiles:.3f}", flush=True)
