"""Does DistributedRenderer deliver, for EVERY frame of a sequence whose sample count changes from frame to
frame, the frame a single unsharded render gives (to the float rounding of the exchanged accumulators)?

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 \
      tools/check_distributed_frames.py [full]

Identical consecutive frames (bench.py) cannot show a frame assembled from the previous frame's shard buffers;
this sequence can.  `full` = BASELINE configs[4]'s size (3840x2160, 64 -> 1024 spp).  Exit code 1 on a deviation
(rank 0 prints which shards deviate)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import api
from rs_pathtracing_b200.distributed import DistributedRenderer, owned_pixel_coords

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
CASES = (("dupin.json", 480, 270, [64, 1024, 64, 1024]), ("cornell_box.json", 1000, 560, [4, 32, 4, 1, 32]),
         ("dupin.json", 3840, 2160, [4, 64, 4, 64]))
if len(sys.argv) > 1 and sys.argv[1] == "full":
    CASES = (("dupin.json", 3840, 2160, [64, 1024]),)
failed = 0
for name, w, h, seq in CASES:
    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", name), random_spheres_seed=1)
    cam = sc.camera()
    dr = DistributedRenderer(sc, 8, seed=2024, tile=32, device=lr)
    frames = []
    for i, spp in enumerate(seq):   # back to back: no barrier, no synchronisation between the frames
        f = dr.render(cam, w, h, spp)
        frames.append(None if f is None else f.copy())
    # ... and the same sequence PIPELINED (frame n's exchange overlaps frame n+1's render): the very same frames
    piped = dr.render_jobs([(cam, w, h, spp) for spp in seq], keep=True)
    if rank == 0:
        for i, f in enumerate(piped):
            same = np.array_equal(f, frames[i])
            failed += int(not same)
            print(f"{name} {w}x{h} pipelined frame {i}: {'identical to the unpipelined one' if same else 'DIFFERS'}", flush=True)
    if rank == 0:
        sc2 = rt.Scene.from_file(os.path.join(ROOT, "scenes", name), random_spheres_seed=1)
        d2 = sc2.device_scene(lr)
        refs = {}
        for i, spp in enumerate(seq):
            if spp not in refs:
                ref = np.full((h, w, 3), -1.0)
                api.render_start(d2, cam, api.render_params(w, h, spp, 8, 2024))
                api.render_wait(d2, ref)
                refs[spp] = ref
            frame, ref = frames[i], refs[spp]
            bad = (np.abs(frame - ref) > 3e-7 * np.maximum(np.abs(ref), 1e-9)).any(axis=2)
            per = []
            for s in range(world):
                x, y, ok = owned_pixel_coords(w, h, 32, world, s)
                per.append(int(bad[y[ok], x[ok]].sum()))
            failed += int(bad.any())
            print(f"{name} {w}x{h} frame {i} of spp sequence {seq}: deviating pixels {int(bad.sum())}, per shard {per}, "
                  f"means {frame.mean():.12f} / {ref.mean():.12f}", flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
if world > 1:
    flag = torch.tensor([failed], device="cuda")
    dist.broadcast(flag, 0)
    failed = int(flag[0])
    dist.destroy_process_group()
sys.exit(1 if failed else 0)
