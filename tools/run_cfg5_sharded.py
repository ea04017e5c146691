"""BASELINE.json configs[4]: scenes/dupin.json (re-authored in the current schema), 3840x2160,
1024 spp, max depth 8, sharded by interleaved 32x32 tiles over the N GPUs of one box (one process per GPU):

  python tools/run_cfg5_sharded.py                                        # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      tools/run_cfg5_sharded.py [--spp 1024]

Rank 0 prints one JSON line: Mpaths/s and ms/frame with the frame left on rank 0's device (barrier +
synchronize on both sides, max over ranks) and end to end through DistributedRenderer.render (frame in pinned
host memory).  Total work is fixed as N grows (strong scaling); the frame is independent of N bit for bit
(tests/test_gpu_render.py::test_sharded_render_assembles_to_the_unsharded_frame).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

W, H, DEPTH = 3840, 2160, 8


def main():
    out_fd = os.dup(1)
    os.dup2(2, 1)   # NCCL / torch banners must not share stdout with the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--check", action="store_true",
                    help="rank 0 also renders the frame unsharded and reports the pixels that differ beyond float rounding")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import numpy as np
    import torch
    import rs_pathtracing_b200 as rt
    from rs_pathtracing_b200.distributed import DistributedRenderer

    if not torch.cuda.is_available() or rt.device_count() == 0:
        raise SystemExit("needs a CUDA device: the core has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", "dupin.json"), random_spheres_seed=1)
    cam = sc.camera()
    dr = DistributedRenderer(sc, DEPTH, seed=2024, tile=32, device=local_rank)
    n_paths = W * H * args.spp
    for _ in range(3):   # warm-up at 1/16 of the samples: same kernels, same buffers
        dr.render_device(cam, W, H, max(args.spp // 16, 1))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.frames):
        dr.render_device(cam, W, H, args.spp)
    barrier()
    dev_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.frames)
    shard_ms = sc.stats(local_rank).last_frame_ms
    slowest_shard_ms = max_over_ranks(shard_ms)
    dr.render(cam, W, H, max(args.spp // 16, 1))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.frames):
        frame = dr.render(cam, W, H, args.spp)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.frames)
    check = None
    if rank == 0 and args.check:
        from rs_pathtracing_b200 import api
        from rs_pathtracing_b200.distributed import owned_pixel_coords
        sc2 = rt.Scene.from_file(os.path.join(ROOT, "scenes", "dupin.json"), random_spheres_seed=1)
        ref = np.full((H, W, 3), -1.0)
        d2 = sc2.device_scene(local_rank)
        api.render_start(d2, cam, api.render_params(W, H, args.spp, DEPTH, 2024))
        api.render_wait(d2, ref)
        bad = (np.abs(frame - ref) > 3e-7 * np.maximum(np.abs(ref), 1e-9)).any(axis=2)
        per = []
        for s in range(world):
            x, y, ok = owned_pixel_coords(W, H, 32, world, s)
            per.append(int(bad[y[ok], x[ok]].sum()))
        ys, xs = np.nonzero(bad)
        check = {"deviating_pixels": int(bad.sum()), "per_shard": per, "unsharded_mean": float(ref.mean()),
                 "first": [[int(a), int(b), frame[b, a].tolist(), ref[b, a].tolist()] for b, a in list(zip(ys, xs))[:6]],
                 "tiles": len(set(zip((xs // 32).tolist(), (ys // 32).tolist())))}
    if rank == 0:
        assert frame is not None and np.isfinite(frame).all()
        line = {"config": f"scenes/dupin.json +481 seeded random spheres ({sc.shape_count} shapes), {W}x{H}, "
                          f"{args.spp} spp, max depth {DEPTH}, interleaved 32x32 tiles over {world} GPU(s)",
                "n_gpus": world, "paths": n_paths, "frames_timed": args.frames,
                "value": n_paths / (dev_ms * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_frame": dev_ms,
                "rank0_shard_device_ms": shard_ms, "slowest_shard_device_ms": slowest_shard_ms,
                "e2e": {"value": n_paths / (e2e_ms * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_frame": e2e_ms,
                        "d2h_bytes_per_frame": W * H * 24},
                "scaling": "strong", "frame_mean": float(frame.mean()), "check_against_unsharded": check}
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
