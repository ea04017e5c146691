"""Host-side logic on CPU: the C-ABI libraries load and export every declared symbol, the JSON loader
mirrors the reference's schema and error behaviour, and the flattened transforms equal the oracle's
restatement bit for bit."""
import ctypes as C
import json
import math
import os
import re

import numpy as np
import pytest

import rs_pathtracing_b200 as rt
from rs_pathtracing_b200 import _ffi
from oracle import pyoracle as po

from conftest import ROOT, scene_path


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(rth?_[a-z0-9_]+)\s*\(", text))


def test_core_exports_every_declared_symbol():
    lib = _ffi.core()
    names = _declared("rt_b200.h")
    assert names == set(_ffi.CORE_SYMBOLS), names ^ set(_ffi.CORE_SYMBOLS)
    for n in names:
        assert hasattr(lib, n)
    assert lib.rt_abi_version() == 5


def test_host_exports_every_declared_symbol():
    lib = _ffi.host()
    names = _declared("rt_b200_host.h") - {"rth_image_loader"}
    assert names == set(_ffi.HOST_SYMBOLS), names ^ set(_ffi.HOST_SYMBOLS)
    for n in names:
        assert hasattr(lib, n)


def test_pod_layouts():
    assert C.sizeof(_ffi.Vec3) == 24 and C.sizeof(_ffi.Ray) == 48 and C.sizeof(_ffi.Camera) == 112
    assert C.sizeof(_ffi.Material) == 16 and C.sizeof(_ffi.Texture) == 40 and C.sizeof(_ffi.RenderParams) == 40
    assert C.sizeof(_ffi.Perlin) == 3 * 1024 + 256 * 24 and C.sizeof(_ffi.SceneDesc) == 120


@pytest.mark.parametrize("name,json_shapes", [("spheres.json", 5), ("cornell_box.json", 9),
                                               ("detached_materials.json", 5), ("dupin.json", 3),
                                               ("cube_test.json", 3), ("empty.json", 0), ("light_source.json", 3)])
def test_scene_loads_and_adds_random_spheres(name, json_shapes):
    bare = rt.Scene.from_file(scene_path(name), add_random_spheres=False)
    assert bare.shape_count == json_shapes
    full = rt.Scene.from_file(scene_path(name), random_spheres_seed=1)
    extra = full.shape_count - json_shapes
    assert 470 <= extra <= 484          # 22*22 grid minus those within 0.9 of (4, 0.2, 0); SURVEY §0.1
    d = full.desc()
    kinds = np.ctypeslib.as_array(d.kind, shape=(d.n_shapes,))
    assert (kinds[json_shapes:] == _ffi.RT_SHAPE_SPHERE).all()
    inv = np.ctypeslib.as_array(d.inverse, shape=(d.n_shapes, 12))
    dirm = np.ctypeslib.as_array(d.direct, shape=(d.n_shapes, 12))
    # radius 0.2, centre (a + 0.9u, 0.2, b + 0.9u) appended in grid order, a outer / b inner
    assert np.allclose(dirm[json_shapes:, [0, 5, 10]], 0.2)
    assert (inv[json_shapes:, [0, 5, 10]] == 5.0).all()
    cx, cy, cz = dirm[json_shapes:, 3], dirm[json_shapes:, 7], dirm[json_shapes:, 11]
    assert (cy == 0.2).all() and (cx >= -11).all() and (cx < 11).all() and (cz >= -11).all() and (cz < 11).all()
    assert (np.sqrt((cx - 4) ** 2 + cz ** 2) > 0.9).all()
    assert (np.diff(np.floor(cx)) >= 0).all()
    # reproducible for a seed, different across seeds
    again = rt.Scene.from_file(scene_path(name), random_spheres_seed=1).desc()
    assert np.array_equal(np.ctypeslib.as_array(again.direct, shape=(d.n_shapes, 12)), dirm)
    other = rt.Scene.from_file(scene_path(name), random_spheres_seed=2)
    od = other.desc()
    assert not np.array_equal(np.ctypeslib.as_array(od.direct, shape=(od.n_shapes, 12))[json_shapes:json_shapes + 5],
                              dirm[json_shapes:json_shapes + 5])


def test_random_sphere_material_mix():
    """80 % Lambertian / 15 % Metal / 5 % Dielectric (json_models.rs:86-112)"""
    counts = np.zeros(3)
    for seed in range(1, 9):
        sc = rt.Scene.from_file(scene_path("empty.json"), random_spheres_seed=seed)
        d = sc.desc()
        mats = [d.materials[d.material[i]] for i in range(d.n_shapes)]
        for m in mats:
            counts[m.kind] += 1
            if m.kind == _ffi.RT_MAT_METAL:
                assert 0.0 <= m.scalar < 0.5
                c = d.textures[m.texture].color
                assert all(0.0 <= v <= 0.5 for v in c.tuple())
            elif m.kind == _ffi.RT_MAT_DIELECTRIC:
                assert m.scalar == 1.5
    frac = counts / counts.sum()
    assert abs(frac[0] - 0.8) < 0.03 and abs(frac[1] - 0.15) < 0.03 and abs(frac[2] - 0.05) < 0.02


def test_flat_transforms_equal_oracle_bitwise():
    """InversableTransform::new in the host mirror vs the oracle's independent restatement"""
    text = json.load(open(scene_path("cornell_box.json")))
    sc = rt.Scene.from_file(scene_path("cornell_box.json"), add_random_spheres=False)
    d = sc.desc()
    inv = np.ctypeslib.as_array(d.inverse, shape=(d.n_shapes, 12))
    dirm = np.ctypeslib.as_array(d.direct, shape=(d.n_shapes, 12))
    for i, s in enumerate(text["shapes"]):
        t = s["transform"]
        od, oi = po.transform_new(t["translate"], t["rotate"], t["scale"])
        assert np.array_equal(od[:3].reshape(12), dirm[i])
        assert np.array_equal(oi[:3].reshape(12), inv[i])
    rng = np.random.default_rng(0)
    lib = _ffi.host()
    for _ in range(50):
        tr, ro, scl = rng.uniform(-100, 100, 3), rng.uniform(-180, 180, 3), rng.uniform(0.1, 50, 3)
        hd, hi = np.empty(16), np.empty(16)
        lib.rth_transform_new(_ffi.Vec3(*tr), _ffi.Vec3(*ro), _ffi.Vec3(*scl), hd.ctypes.data_as(C.POINTER(C.c_double)),
                              hi.ctypes.data_as(C.POINTER(C.c_double)))
        od, oi = po.transform_new(tr, ro, scl)
        assert np.array_equal(hd.reshape(4, 4), od) and np.array_equal(hi.reshape(4, 4), oi)


def test_camera_matches_oracle_and_reference_kat():
    cam = rt.camera_new((0, 0, 0), (0, 0, -1), (0, 1, 0), 1.0, math.radians(90.0))
    assert cam.right.tuple() == (1.0, 0.0, 0.0)   # src/camera/mod.rs:333
    rng = np.random.default_rng(2)
    for _ in range(20):
        p, dvec, up = rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
        a = rt.camera_new(p, dvec, up, 1.3, 0.7)
        b = po.camera_new(p, dvec, up, 1.3, 0.7)
        assert bytes(a) == bytes(b)
    sc = rt.Scene.from_file(scene_path("cornell_box.json"), add_random_spheres=False)
    c = sc.camera()
    assert c.fov_rad == 40.0 * (math.pi / 180.0) and c.position.tuple() == (278.0, 278.0, -800.0)


def test_loader_error_behaviour():
    base = json.load(open(scene_path("cube_test.json")))
    def load(d):
        return rt.Scene.from_json(json.dumps(d), add_random_spheres=False)
    for key in ("camera", "shapes", "materials", "background"):   # SceneJson: all four required
        d = dict(base); del d[key]
        with pytest.raises(rt.RtError, match="missing field"):
            load(d)
    d = json.loads(json.dumps(base)); d["shapes"][0]["material"] = "Nope"
    with pytest.raises(rt.RtError, match="not found"):
        load(d)
    d = json.loads(json.dumps(base)); d["shapes"][0]["type"] = "TransformedSphere"   # dupin.json's old tag
    with pytest.raises(rt.RtError, match="unknown variant"):
        load(d)
    with pytest.raises(rt.RtError):
        rt.Scene.from_json("{ not json", add_random_spheres=False)
    # both Vector3d spellings, unknown keys ignored, depth defaults to 4
    d = json.loads(json.dumps(base))
    d["shapes"][0]["transform"]["translate"] = {"x": 5, "y": 6, "z": 0}
    d["shapes"][0]["bogus"] = 1
    d["shapes"].append({"type": "BruteForsableShape", "shape": {"type": "Heart"}, "step": 0.01,
                        "transform": {"translate": [0, 0, 0], "rotate": [0, 0, 0], "scale": [1, 1, 1]}, "material": "Red"})
    sc = load(d)
    dd = sc.desc()
    params = np.ctypeslib.as_array(dd.params, shape=(dd.n_shapes, 8))
    assert params[-1][0] == _ffi.RT_SURF_HEART and params[-1][1] == 0.01 and params[-1][2] == 4.0
    ref = load(base).desc()
    assert np.array_equal(np.ctypeslib.as_array(ref.direct, shape=(ref.n_shapes, 12))[0],
                          np.ctypeslib.as_array(dd.direct, shape=(dd.n_shapes, 12))[0])


def test_image_texture_loaded_and_material_reassign():
    sc = rt.Scene.from_file(scene_path("detached_materials.json"), add_random_spheres=False)
    d = sc.desc()
    assert d.n_images == 1 and d.images[0].width == 1024 and d.images[0].height == 512
    names = [_ffi.host().rth_scene_material_name(sc._h, i).decode() for i in range(d.n_materials)]
    assert "EarthMap" in names
    sc.assign_material(1, "EarthMap")
    assert sc.desc().material[1] == names.index("EarthMap")
    with pytest.raises(rt.RtError):
        sc.assign_material(1, "Nope")


def test_compute_entry_points_fail_loudly_without_gpu():
    if rt.device_count() > 0:
        pytest.skip("GPU present")
    sc = rt.Scene.from_file(scene_path("cube_test.json"), add_random_spheres=False)
    with pytest.raises(rt.RtError, match="no CUDA device"):
        sc.closest_hit(np.zeros((1, 6)))
    with pytest.raises(rt.RtError, match="no CUDA device"):
        rt.GpuRenderer(sc, 12, 8)
    with pytest.raises(rt.RtError, match="no CUDA device"):
        rt.measure_peaks()


def test_shard_float4_count():
    lib = _ffi.core()
    p = _ffi.RenderParams()
    p.image = _ffi.ImageParams(100, 70)
    p.shard_count = 1
    assert lib.rt_shard_float4_count(C.byref(p), 0) == 7000
    p.shard_count, p.tile_width, p.tile_height = 4, 32, 32      # 4 x 3 = 12 tiles -> 3 per shard
    assert [lib.rt_shard_float4_count(C.byref(p), s) for s in range(4)] == [3 * 1024] * 4
    p.shard_count = 5                                            # 12 tiles over 5 shards: 3,3,2,2,2
    assert [lib.rt_shard_float4_count(C.byref(p), s) for s in range(5)] == [3072, 3072, 2048, 2048, 2048]


def test_march_region_bounds_dominate_sampled_derivatives():
    """the interval bounds behind the exact-skip marcher (csrc/march_bounds.hpp): G and H must dominate
    the gradient norm and the second directional derivative of every surface polynomial anywhere in its
    marching region"""
    lib = _ffi.core()
    orc = po.lib()
    rng = np.random.default_rng(0)
    cases = {0: [0, 0.01, 4, 0, 0, 0, 0, 0], 1: [1, 0.01, 4, 0.7, 0, 0, 0, 2.0], 2: [2, 0.01, 4, 1.3, 0, 0, 0, 2.0],
             3: [3, 0.01, 4, 1.11, 0.99, 0.5, 0.1, 2.5], 4: [4, 0.01, 4, 0, 0, 0, 0, 5.0], 5: [5, 0.01, 4, 0, 0, 0, 0, 1.5]}
    for k, q in cases.items():
        qa = (C.c_double * 8)(*q)
        G, H = C.c_double(), C.c_double()
        assert lib.rt_march_region_bounds(qa, C.byref(G), C.byref(H)) == 0
        G, H = G.value, H.value
        assert math.isfinite(H) and H > 0 and math.isfinite(G) and G > 0
        radii = np.array([1.45, 1.45 / 2.05, 1.45]) if k == 0 else np.array([q[7]] * 3)
        f = lambda p: orc.orc_surface_func(qa, po.Vec3(*p))
        worst2 = worst1 = 0.0
        for _ in range(400):
            v = rng.normal(size=3); v /= np.linalg.norm(v)
            p = v * rng.uniform(0, 1) ** (1 / 3) * radii * 1.05
            u = rng.normal(size=3); u /= np.linalg.norm(u)
            h = 1e-4 * radii.min()
            worst2 = max(worst2, abs((f(p + h * u) - 2 * f(p) + f(p - h * u)) / (h * h)))
            grad = np.array([(f(p + h * e) - f(p - h * e)) / (2 * h) for e in np.eye(3)])
            worst1 = max(worst1, np.linalg.norm(grad))
        assert worst2 <= H * (1 + 1e-6) + 1e-6, (k, worst2, H)
        assert worst1 <= G * (1 + 1e-6) + 1e-6, (k, worst1, G)
        assert worst2 > H / 200 and worst1 > G / 200, (k, worst1, G, worst2, H)   # and not absurdly loose


def test_rust_sys_crate_covers_the_abi():
    """rust/ray_tracing_b200-sys/src/lib.rs (uncompiled: no Rust toolchain here) must declare every symbol
    of include/rt_b200.h, the ABI version the header states, and mirror rt_stats / rt_scene_desc field
    for field."""
    src = open(os.path.join(ROOT, "rust", "ray_tracing_b200-sys", "src", "lib.rs")).read()
    hdr = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    for name in _declared("rt_b200.h"):
        assert re.search(r"\bpub fn %s\(" % name, src), name
    version = re.search(r"#define RT_B200_ABI_VERSION (\d+)", hdr).group(1)
    assert f"RT_B200_ABI_VERSION: c_int = {version};" in src
    for struct, cls in (("rt_stats", _ffi.Stats), ("rt_scene_desc", _ffi.SceneDesc)):
        body = re.search(r"pub struct %s \{(.*?)\}" % struct, src, re.S).group(1)
        rust_fields = re.findall(r"pub (\w+):", body)
        assert rust_fields == [f[0] for f in cls._fields_], struct


def test_save_png_round_trips(tmp_path):
    """rth_save_png (image::save_buffer of the bins): a valid PNG that decodes to the same RGBA bytes"""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for w, h in ((7, 5), (300, 260)):          # 300*260*4 + 260 > 65535: several stored deflate blocks
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        path = str(tmp_path / f"t{w}.png")
        rt.save_png(path, rgba)
        with Image.open(path) as im:
            assert im.mode == "RGBA" and im.size == (w, h)
            assert np.array_equal(np.asarray(im), rgba)
    with pytest.raises(rt.RtError):
        rt.save_png(str(tmp_path / "no_such_dir" / "x.png"), rgba)


@pytest.mark.parametrize("name", ["spheres.json", "cornell_box.json", "detached_materials.json", "dupin.json",
                                  "light_source.json"])
def test_cull_tree_nodes_enclose_their_members(name):
    """host-side invariant of the conservative cull tree (csrc/rt_cull.cuh): with the centre taken from the
    shape's own direct transform and the radius from its inverse, every leaf ball lies inside its group
    ball and inside its root ball, every group ball inside its root ball (FP64, no device needed)"""
    sc = rt.Scene.from_file(scene_path(name), random_spheres_seed=1)
    d = sc.desc()
    roots, groups, tree, flat = (C.c_uint32() for _ in range(4))
    worst = C.c_double()
    rc = _ffi.core().rt_cull_tree_check(C.byref(d), C.byref(roots), C.byref(groups), C.byref(tree), C.byref(flat),
                                       C.byref(worst))
    assert rc == 0
    kinds = sc.shape_kinds()
    assert tree.value + flat.value == int((kinds != _ffi.RT_SHAPE_MARCH).sum())
    assert tree.value >= 470 and roots.value == 1 and groups.value % 8 == 0 and groups.value * 16 >= tree.value
    assert flat.value <= 12                      # Rectangles, cubes next to them, the ground / sun spheres
    assert worst.value <= 1e-7, worst.value      # relative reach of a member beyond its node's radius (header)


def test_cull_tree_invariant_on_a_large_random_scene():
    """several roots (> 512 shapes under the tree), rotated / anisotropic shapes, a cluster far from the origin"""
    import json
    rng = np.random.default_rng(11)
    shapes = []
    for k in range(1400):
        c = rng.uniform(-20, 20, 3) + (np.array([3e5, -1e5, 2e5]) if k % 5 == 0 else 0.0)
        r = float(rng.uniform(0.05, 0.5))
        shapes.append({"type": "Cube" if k % 3 == 0 else "Sphere", "name": f"s{k}", "material": "M",
                       "transform": {"translate": c.tolist(), "rotate": rng.uniform(-90, 90, 3).tolist(),
                                     "scale": [r, r * float(rng.uniform(0.6, 1.6)), r * float(rng.uniform(0.6, 1.6))]}})
    scene = {"camera": {"position": [0, 0, -10], "direction": [0, 0, 1], "up": [0, 1, 0], "fov": 40.0, "focal_length": 1.0},
             "background": [0, 0, 0],
             "materials": {"M": {"type": "Lambertian", "albedo": {"type": "SolidColor", "color": [0.9, 0.1, 0.1]}}},
             "shapes": shapes}
    sc = rt.Scene.from_json(json.dumps(scene), add_random_spheres=False)
    d = sc.desc()
    roots, groups, tree, flat = (C.c_uint32() for _ in range(4))
    worst = C.c_double()
    assert _ffi.core().rt_cull_tree_check(C.byref(d), C.byref(roots), C.byref(groups), C.byref(tree), C.byref(flat),
                                         C.byref(worst)) == 0
    assert tree.value + flat.value == 1400 and tree.value > 1300 and roots.value >= 3
    assert worst.value <= 1e-7, worst.value


def test_oracle_abi_types_match_the_product_bindings():
    """oracle/abi_types.py is a second copy of the plain-data C-ABI structs (so that bench.py's reference arm runs
    the oracle without importing the product): field names, ctypes and sizes must agree with _ffi.py"""
    import ctypes as C
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle_abi_types", os.path.join(ROOT, "oracle", "abi_types.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from rs_pathtracing_b200 import _ffi

    def flat(t):
        out = []
        for name, ct in t._fields_:
            if isinstance(ct, type) and issubclass(ct, C.Structure):
                out.append((name, flat(ct)))
            elif hasattr(ct, "_type_") and isinstance(ct._type_, type) and issubclass(ct._type_, C.Structure):
                out.append((name, getattr(ct, "_length_", "ptr"), flat(ct._type_)))
            else:
                out.append((name, C.sizeof(ct), getattr(ct, "_type_", None) if not hasattr(ct, "contents") else "ptr"))
        return out

    for name in ("Vec3", "Ray", "Camera", "ImageParams", "Material", "Texture", "Image", "Perlin", "SceneDesc"):
        a, b = getattr(mod, name), getattr(_ffi, name)
        assert C.sizeof(a) == C.sizeof(b), name
        assert [f[0] for f in a._fields_] == [f[0] for f in b._fields_], name
        assert flat(a) == flat(b), name


def test_committed_flat_scenes_equal_the_loader_output(tmp_path):
    """oracle/scenes/cfg*.npz (what bench.py --impl reference renders) against a fresh flattening by the host mirror"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_oracle_scenes import config_scene
    from oracle import pyoracle as po
    for cfg in ("1", "3", "4a", "4b", "5"):
        sc, cam, note = config_scene(cfg)
        fresh = os.path.join(tmp_path, f"cfg{cfg}.npz")
        po.save_flat_scene(fresh, sc.desc(), cam, note)
        a, b = np.load(fresh), np.load(os.path.join(ROOT, "oracle", "scenes", f"cfg{cfg}.npz"))
        assert sorted(a.files) == sorted(b.files), cfg
        for k in a.files:
            assert np.array_equal(a[k], b[k]), (cfg, k)


def test_bench_reference_arm_runs_the_oracle_alone():
    """`bench.py --impl reference` prints the contract's JSON line and loads nothing of the product package (it
    asserts that itself); config 2 is the quick one"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
