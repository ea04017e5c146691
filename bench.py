#!/usr/bin/env python
"""bench.py — Mpaths/s (pixels x spp / s) of the path-tracing hot path on cornell_box.json.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one frame: scenes/cornell_box.json (+ seeded random spheres, N = 492 shapes),
1024 x 1024, 256 spp, max depth 8 (BASELINE.json configs[2], the configuration the metric is quoted
on).  One process per GPU (torchrun for N > 1); the frame is sharded by interleaved 32x32 tiles and
gathered with one NCCL exchange, so total work is fixed as N grows ("strong" scaling).

Output: ONE JSON line on rank 0 (see the task contract): value = device-timed whole-job Mpaths/s with
the scene resident, e2e = the same through the Renderer API with host buffers, roofline for the
dominant kernel (FP64 issue bound; HBM traffic reported beside it), cpu_baseline = the oracle's
threaded renderer (C++ restatement of the reference, BVH like Scene::new builds) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENE = os.path.join(ROOT, "scenes", "cornell_box.json")
WIDTH, HEIGHT, SPP, DEPTH = 1024, 1024, 256, 8
SCENE_SEED, RNG_SEED = 1, 2024
WORKLOAD = "scenes/cornell_box.json +481 seeded random spheres (492 shapes), 1024x1024, 256 spp, max depth 8"

# Algorithmic FP64 work per unit, counted on the reference's formulation:
FLOPS_PER_SHAPE_TEST = 52      # ray -> object space (33) + unit-sphere discriminant (19)
FLOPS_PER_CULL_TEST = 20       # executed FP32 flops of one conservative ball pre-test (rt_cull.cuh: 3 sub, 9 FMA, 1 mul, 1 add)
FLOPS_PER_MARCH_STEP = 22      # t/p advance (7) + Heart polynomial (15)
FLOPS_PER_SEGMENT = 135        # winner's hit record (75) + shade (~60)
# Algorithmic HBM bytes per segment (ray 48 + throughput 24 + id 4, read once and written once) and
# per path (float4 radiance write + read, float4 accumulate + f64 frame write)
BYTES_PER_SEGMENT = 2 * 76
BYTES_PER_PATH = 16 + 16 + (16 + 24) / SPP


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(steps: int, warmup: int, sample=None):
    """The reference arm: the oracle's threaded renderer (dispatcher + workers, BVH as Scene::new
    builds) on all host threads, each step a bounded sample of the workload."""
    import rs_pathtracing_b200 as rt
    from oracle import pyoracle as po

    sc = rt.Scene.from_file(SCENE, random_spheres_seed=SCENE_SEED)
    cam = sc.camera()
    osc = po.OracleScene(sc.desc())
    osc.build_bvh(seed=1)
    threads = po.hardware_threads()
    stride, spp = sample or ((8, 8), 64)
    n_paths = (WIDTH // stride[0]) * (HEIGHT // stride[1]) * spp
    secs = []
    for i in range(warmup + steps):
        _, info = osc.render(cam, WIDTH, HEIGHT, spp, DEPTH, seed=RNG_SEED + i, rng="xoshiro", use_bvh=True,
                             threads=threads, stride=stride)
        if i >= warmup:
            secs.append(info["seconds"])
    mean_s = sum(secs) / len(secs)
    desc = (f"every {stride[0]}x{stride[1]}-th pixel of the 1024x1024 frame at {spp} spp, depth 8 = {n_paths} paths "
            f"per step; C++ restatement of the reference's threaded renderer with its BVH, {threads} worker threads "
            f"+ 1 serial dispatcher (not the Rust binary: no Rust toolchain in the image)")
    return {"mpaths": n_paths / mean_s / 1e6, "ms": mean_s * 1e3, "cores": threads, "sample": desc,
            "n_paths": n_paths}


def emit(line: dict) -> None:
    """the ONE JSON line, on the real stdout"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # libraries (NCCL's version banner, torch warnings) must not share stdout with the JSON line: keep a
    # private duplicate of fd 1 for it and point fd 1 at stderr for everybody else
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": WORKLOAD, "scene_seed": SCENE_SEED, "parallelism": f"interleaved 32x32 tiles over {world} GPU(s)",
              "l2_policy": "per-step inputs+queues (>= 1.9 GB of path state streamed per frame) exceed the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(args.steps, args.warmup)
        line = {"impl": "reference", "metric": "Mpaths/s", "value": r["mpaths"], "unit": "Mpaths/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["mpaths"], "unit": "Mpaths/s", "cores": r["cores"], "kind": "port",
                                 "sample": r["sample"]},
                "e2e": {"value": r["mpaths"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    import numpy as np
    import torch
    import rs_pathtracing_b200 as rt
    from rs_pathtracing_b200 import api, _ffi
    from rs_pathtracing_b200.distributed import DistributedRenderer
    import ctypes as C

    if not torch.cuda.is_available() or rt.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the core has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    t_up0 = time.perf_counter()
    sc = rt.Scene.from_file(SCENE, random_spheres_seed=SCENE_SEED)
    cam = sc.camera()
    dr = DistributedRenderer(sc, DEPTH, seed=RNG_SEED, tile=32, device=local_rank)
    scene_upload_ms = (time.perf_counter() - t_up0) * 1e3
    n_paths = WIDTH * HEIGHT * SPP

    # --- kernel-side measurement: device-resident inputs, frame left on the device ---------------
    for _ in range(args.warmup):
        dr.render_device(cam, WIDTH, HEIGHT, SPP)
    barrier()
    sc.reset_stats(local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    frame_ms = []
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dr.render_device(cam, WIDTH, HEIGHT, SPP)
        frame_ms.append(sc.stats(local_rank).last_frame_ms)   # CUDA events on the library's stream
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = sc.stats(local_rank).kernel_launches
    # per-kernel device time: the same K steps again with one CUDA event pair around every launch on the
    # library's stream (the event pairs cost a few percent, so they stay out of the pass `value` comes from)
    sc.set_kernel_timing(True, local_rank)
    sc.reset_stats(local_rank)
    timed_ms = []
    for _ in range(args.steps):
        dr.render_device(cam, WIDTH, HEIGHT, SPP)
        timed_ms.append(sc.stats(local_rank).last_frame_ms)
    barrier()
    kst = sc.stats(local_rank)
    kernel_ms = {"k_raygen": kst.ms_raygen / args.steps, "k_extend": kst.ms_extend / args.steps,
                 "k_march+k_replay": kst.ms_march / args.steps, "k_shade": kst.ms_shade / args.steps,
                 "k_resolve": kst.ms_resolve / args.steps}
    extend_launches = kst.launches_extend / args.steps
    sc.set_kernel_timing(False, local_rank)
    step_ms = wall_s * 1e3 / args.steps
    t = torch.tensor([step_ms, float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        step_ms, launches = float(mx[0]), int(sm[1])
    value = n_paths / (step_ms * 1e-3) / 1e6

    # --- end to end through the public API: host buffers, D2H inside the timed region ------------
    for _ in range(1):
        dr.render(cam, WIDTH, HEIGHT, SPP)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_frame = dr.render(cam, WIDTH, HEIGHT, SPP)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    e2e_value = n_paths / (e2e_ms * 1e-3) / 1e6
    h2d = C.sizeof(_ffi.Camera) + C.sizeof(_ffi.RenderParams)
    d2h = WIDTH * HEIGHT * 24

    # --- frame check (untimed): consecutive bench frames are identical, which would hide a frame assembled from
    # stale shard buffers; so one more frame with ANOTHER seed and fewer samples goes through the same
    # DistributedRenderer and is compared, on rank 0, with the unsharded render of that seed -----------------
    frame_check = None
    try:
        dr.seed = RNG_SEED + 1
        check_spp = 16
        chk = dr.render(cam, WIDTH, HEIGHT, check_spp)        # collective: every rank takes part
        dr.seed = RNG_SEED
        if rank == 0:
            sc2 = rt.Scene.from_file(SCENE, random_spheres_seed=SCENE_SEED)
            ref = np.full((HEIGHT, WIDTH, 3), -1.0)
            d2 = sc2.device_scene(local_rank)
            api.render_start(d2, cam, api.render_params(WIDTH, HEIGHT, check_spp, DEPTH, RNG_SEED + 1))
            api.render_wait(d2, ref)
            # the exchanged accumulators hold float sums: equal to the f64 frame up to float rounding
            bad = (np.abs(chk - ref) > 3e-7 * np.maximum(np.abs(ref), 1e-9)).any(axis=2)
            frame_check = {"what": f"DistributedRenderer frame (seed {RNG_SEED + 1}, {check_spp} spp) against the unsharded "
                                   "render on rank 0, tolerance 3e-7 relative", "deviating_pixels": int(bad.sum())}
    except Exception as e:   # the check must never cost the bench line
        frame_check = {"error": repr(e)}
    barrier()

    line = None
    if rank == 0:
        assert host_frame is not None and np.isfinite(host_frame).all() and host_frame.mean() > 0.01
        # --- roofline of the dominant kernel (k_extend), rank 0's shard --------------------------------
        # work counters from one instrumented frame at reduced spp (counts scale linearly with spp)
        count_spp = 4
        sc.set_counters(True, local_rank)
        sc.reset_stats(local_rank)
        p = api.render_params(WIDTH, HEIGHT, count_spp, DEPTH, RNG_SEED, world, 0, tile=32)
        api.render_start(dr.dev_scene, cam, p)
        api.render_wait(dr.dev_scene, None)
        st = sc.stats(local_rank)
        sc.set_counters(False, local_rank)
        scale = SPP / count_spp
        segs, exact, culls, msteps = (st.segments * scale, st.shape_tests * scale, st.cull_tests * scale,
                                      st.march_steps * scale)
        n_shapes = sc.shape_count
        shard_ms = sum(timed_ms) / len(timed_ms)   # the pass the per-kernel times come from
        fp64_peak, fp32_peak = rt.measure_peaks(local_rank)
        # ALGORITHMIC work of k_extend: the reference tests every shape for every segment
        # (ShapeCollection::ray_intersect), 52 flop per (segment, shape) pair -- whether or not we cull it
        ext_ms = kernel_ms["k_extend"]
        ext_flops = segs * n_shapes * FLOPS_PER_SHAPE_TEST
        achieved = ext_flops / (ext_ms * 1e-3) / 1e12
        # EXECUTED work of k_extend: FP32 pre-tests (node + leaf ball tests) and exact FP64 tests
        cull_tflops = culls * FLOPS_PER_CULL_TEST / (ext_ms * 1e-3) / 1e12
        exact_tflops = exact * FLOPS_PER_SHAPE_TEST / (ext_ms * 1e-3) / 1e12
        frame_flops = ext_flops + msteps * FLOPS_PER_MARCH_STEP + segs * FLOPS_PER_SEGMENT
        hbm_bytes = segs * BYTES_PER_SEGMENT + (n_paths / world) * BYTES_PER_PATH
        hbm_gbs = hbm_bytes / (shard_ms * 1e-3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        try:   # dram bytes per k_extend launch from the committed ncu --set full capture (profiles/)
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["k_extend"]["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {
            "bound": "issue (fp64 reference arithmetic; no dense contraction, tensor cores unused)",
            "kernel": "k_extend", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved / fp64_peak,
            "peak_source": "rt_measure_peaks: DFMA micro-kernel timed live on this GPU (FMA = 2 flop). The bit-exact "
                           "contract forbids FMA contraction, so the literal brute-force loop tops out at 0.5; the "
                           "conservative FP32 pre-tests skip most of the algorithmic work, so frac may exceed 1",
            "kernel_ms_per_step": ext_ms, "launches_per_step": extend_launches,
            "ms_per_launch": ext_ms / max(extend_launches, 1), "share_of_step": ext_ms / shard_ms,
            "algorithmic_flops_per_step": ext_flops,
            "executed": {"fp32_pretest_tflops": cull_tflops, "fp32_peak_tflops": fp32_peak,
                         "fp32_frac": cull_tflops / fp32_peak, "fp64_exact_tflops": exact_tflops,
                         "fp64_frac": exact_tflops / fp64_peak, "pretests_per_segment": culls / max(segs, 1),
                         "exact_tests_per_segment": exact / max(segs, 1)},
            "flops_model": {"per_shape_test": FLOPS_PER_SHAPE_TEST, "per_pretest_fp32": FLOPS_PER_CULL_TEST,
                            "per_march_step": FLOPS_PER_MARCH_STEP, "per_segment": FLOPS_PER_SEGMENT,
                            "segments": segs, "shapes": n_shapes, "march_steps": msteps,
                            "frame_algorithmic_flops": frame_flops,
                            "frame_achieved_tflops": frame_flops / (shard_ms * 1e-3) / 1e12},
            "kernel_ms": kernel_ms,
            "traffic": traffic,
            "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                    "algorithmic_bytes_per_step": hbm_bytes},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(steps=1, warmup=0, sample=((4, 4), 32))
            cpu = {"value": r["mpaths"], "unit": "Mpaths/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                   "ms": r["ms"]}
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "device_ms_per_step_rank0": sum(frame_ms) / len(frame_ms),
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "scene_upload_ms_once": scene_upload_ms},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "frame_check": frame_check,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
