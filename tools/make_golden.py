"""Generates tests/golden/*: the fixtures the parity tests replay on the CPU (oracle) and on the GPU.

  python tools/make_golden.py

1. reference_kat.json — the reference's own asserting tests on this path, as data (inputs, expected
   values and tolerance, each citing the reference test it restates).  These PIN the oracle.
2. intersect_<scene>.npz — seeded ray batches and the oracle's nearest hit for them (index, t, normal,
   point, uv, front face).  The reference is Rust and cannot be run in this image, and its tests hold
   no golden vector for nearest hit / t / normal, so these vectors are ORACLE-generated ("parity
   unpinned" by the reference; they freeze the restatement so that oracle and CUDA path cannot drift
   together unnoticed).
3. render_<scene>.npz — small frames of the oracle with the Philox stream (path-for-path comparable).

Only numpy + the oracle are used; nothing under /root/reference is read.
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rs_pathtracing_b200 as rt  # host mirror only: scene loading / flattening (no GPU needed)
from oracle import pyoracle as po

GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES = ["spheres.json", "cornell_box.json", "detached_materials.json", "dupin.json", "cube_test.json",
          "light_source.json"]
N_RAYS = 2048

KAT = {
    "comment": "Known-answer tests held by the reference for the hot path (SURVEY.md 8c). tol = approx_equal's 1e-15 "
               "(src/algebra/mod.rs:14-17) unless exact.",
    "rotate_matrix": {
        "cite": "src/algebra/transform.rs:637-646", "rotate_deg": [0.0, -90.0, 0.0], "point": [0.0, 0.0, -1.0],
        "expected": [1.0, 0.0, 0.0], "tol": 1e-15},
    "matrix_multiplication": {
        "cite": "src/algebra/transform.rs:665-691", "m1": list(range(1, 17)), "m2": list(range(17, 33)),
        "m1m2": {"0,0": 250.0, "1,0": 618.0, "2,3": 1112.0}, "m2m1": {"0,0": 538.0, "1,0": 650.0, "2,3": 1080.0},
        "tol": 0.0},
    "bound_transform": {
        "cite": "src/world/shapes/mod.rs:880-899", "translate": [-10.0, 5.0, 2.5], "rotate": [0.0, 0.0, 0.0],
        "scale": [2.0, 2.0, 2.0], "aabb": [[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]],
        "expected": [[-12.0, 3.0, 0.5], [-8.0, 7.0, 4.5]], "tol": 1e-15},
    "camera": {
        "cite": "src/camera/mod.rs:315-343", "position": [0.0, 0.0, 0.0], "direction": [0.0, 0.0, -1.0],
        "up": [0.0, 1.0, 0.0], "focal_length": 1.0, "fov_deg": 90.0, "expected_right": [1.0, 0.0, 0.0],
        "image": [1920, 1080], "expected_pixel_resolution": 2.0 / 1920, "tol": 1e-15},
    "torus_ray_input_only": {
        "cite": "src/world/shapes/mod.rs:853-860 (prints, asserts nothing)", "origin": [0.0, 0.0, -10.0],
        "direction": [0.0, 0.0, 1.0]},
}


def scene_rays(sc, n, seed):
    from test_gpu_intersect import scene_rays as f
    return f(sc, n, seed)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    with open(os.path.join(GOLDEN, "reference_kat.json"), "w") as f:
        json.dump(KAT, f, indent=1)
    from test_gpu_intersect import TRIO, bench_rays
    cases = [("trio", rt.Scene.from_json(json.dumps(TRIO), add_random_spheres=False), None)]
    for name in SCENES:
        cases.append((name.replace(".json", ""), rt.Scene.from_file(os.path.join(ROOT, "scenes", name), 1), name))
    for tag, sc, name in cases:
        rays = bench_rays(N_RAYS, seed=42) if name is None else scene_rays(sc, N_RAYS, seed=7)
        osc = po.OracleScene(sc.desc())
        want = osc.intersect_batch(rays)
        np.savez_compressed(os.path.join(GOLDEN, f"intersect_{tag}.npz"), rays=rays, index=want["index"], t=want["t"],
                            normal=want["normal"], point=want["point"], uv=want["uv"], front=want["front"],
                            n_shapes=np.int64(sc.shape_count))
        print(tag, "shapes", sc.shape_count, "hits", int((want["index"] >= 0).sum()), "of", len(rays))
    for name, w, h, spp, depth in [("spheres.json", 48, 36, 4, 8), ("cornell_box.json", 32, 32, 4, 8),
                                   ("detached_materials.json", 48, 27, 4, 8), ("dupin.json", 48, 27, 4, 8),
                                   ("light_source.json", 48, 27, 4, 8)]:
        sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", name), 1)
        osc = po.OracleScene(sc.desc())
        frame, info = osc.render(sc.camera(), w, h, spp, depth, seed=11, rng="philox")
        np.savez_compressed(os.path.join(GOLDEN, f"render_{name.replace('.json', '')}.npz"), frame=frame,
                            params=np.array([w, h, spp, depth, 11], dtype=np.int64))
        print(name, "frame mean", float(frame.mean()))


if __name__ == "__main__":
    main()
