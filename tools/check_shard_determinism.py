"""Is the frame independent of the shard count at the sizes of BASELINE configs[4] (clipped tiles: 2160 is not a
multiple of 32; 1024 spp)?  Renders scenes/dupin.json unsharded and as 2 / 4 / 8 shards on ONE GPU (host path:
every shard scatters its owned pixels into the caller's buffer) and compares bit for bit; also the float4
accumulators assembled on the device (what the NCCL path delivers) against the f64 frame.

  python tools/check_shard_determinism.py [--size W H] [--spp S] [--scene dupin.json]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="dupin.json")
ap.add_argument("--size", type=int, nargs=2, default=[3840, 2160])
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--shards", type=int, nargs="*", default=[2, 4, 8])
a = ap.parse_args()
w, h = a.size
path = os.path.join(ROOT, "scenes", a.scene)


def load():
    return rt.Scene.from_file(path, random_spheres_seed=1)


sc = load()
cam = sc.camera()
ds = sc.device_scene(0)
ref = np.full((h, w, 3), -1.0)
api.render_start(ds, cam, api.render_params(w, h, a.spp, 8, 2024))
api.render_wait(ds, ref)
again = np.full((h, w, 3), -1.0)
api.render_start(ds, cam, api.render_params(w, h, a.spp, 8, 2024))
api.render_wait(ds, again)
print(f"{a.scene} {w}x{h}x{a.spp}: mean {ref.mean():.12f}; run-to-run differing pixels: "
      f"{int((ref != again).any(axis=2).sum())}")
rc = 0
for shards in a.shards:
    buf = np.full((h, w, 3), -1.0)
    scenes, ptrs = [], []
    for s in range(shards):
        scn = load()
        scenes.append(scn)
        d = scn.device_scene(0)
        p = api.render_params(w, h, a.spp, 8, 2024, shards, s, tile=32)
        api.render_start(d, cam, p)
        api.render_wait(d, buf)
        ptr, n = api.render_device_result(d)
        ptrs.append(ptr)
    diff = (buf != ref).any(axis=2)
    frame = torch.empty((h, w, 3), dtype=torch.float64, device="cuda")
    api.assemble_frame(scenes[0].device_scene(0), api.render_params(w, h, a.spp, 8, 2024, shards, 0, tile=32), ptrs,
                       frame.data_ptr())
    torch.cuda.synchronize()
    asm = frame.cpu().numpy()
    rel = np.abs(asm - ref) / np.maximum(np.abs(ref), 1e-9)
    print(f"  {shards} shards: host-path pixels differing from the unsharded frame: {int(diff.sum())}"
          f" (first: {np.argwhere(diff)[:3].tolist()}); assembled float4 frame max rel diff {rel.max():.3e}, "
          f"mean {asm.mean():.12f}")
    if diff.any() or rel.max() > 3e-7:
        rc = 1
sys.exit(rc)
