"""Flatten the scenes of BASELINE.json's configurations once, with the host mirror (Scene::from_json + seeded
add_random_spheres), into oracle/scenes/cfg*.npz -- the files bench.py's reference arm loads so that it runs the
oracle WITHOUT importing the product package.  tests/test_host.py checks that the committed files still equal what
the loader produces.

  python tools/make_oracle_scenes.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rs_pathtracing_b200 as rt
from oracle import pyoracle as po

OUT = os.path.join(ROOT, "oracle", "scenes")
SCENE_SEED = 1


def config_scene(cfg):
    """(scene, camera, description) of a BASELINE configuration"""
    name = {"1": "spheres.json", "3": "cornell_box.json", "4a": "detached_materials.json",
            "4b": "detached_materials.json", "5": "dupin.json"}[cfg]
    sc = rt.Scene.from_file(os.path.join(ROOT, "scenes", name), random_spheres_seed=SCENE_SEED)
    cam = sc.camera()
    note = f"scenes/{name} + add_random_spheres(seed {SCENE_SEED})"
    if cfg == "4b":
        sc.assign_material(1, "EarthMap")        # Sphere1  -> Metal + ImageTexture
        sc.assign_material(2, "Glass")           # Cushion  -> Dielectric
        sc.assign_material(5, "Lambertian01")    # a random sphere -> Lambertian + UVChecker
        sc.assign_material(6, "WhiteMirror")
        pos = np.array(cam.position.tuple())
        cam = rt.camera_new(pos, -pos, (0, 1, 0), 1.0, cam.fov_rad)
        note += "; every material / texture kind assigned, camera looking at the origin"
    return sc, cam, note


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    import json
    from bench import TRIO_SCENE
    trio = rt.Scene.from_json(json.dumps(TRIO_SCENE), add_random_spheres=False)
    po.save_flat_scene(os.path.join(OUT, "cfg2.npz"), trio.desc(), trio.camera(),
                       "bench trio {unit Sphere, unit Cube, Heart}: benches/bench_intersections.rs:16-66")
    for cfg in ("1", "3", "4a", "4b", "5"):
        sc, cam, note = config_scene(cfg)
        path = os.path.join(OUT, f"cfg{cfg}.npz")
        po.save_flat_scene(path, sc.desc(), cam, note)
        print(path, os.path.getsize(path), "bytes,", sc.shape_count, "shapes")
