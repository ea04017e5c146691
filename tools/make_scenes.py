"""Re-author the reference's scene fixtures into scenes/ (run in the authoring container only: it
reads /root/reference/scenes, which does not exist on the GPU box).

The numbers are the reference's (they define the named workloads); the files are normalised:
vectors become [x, y, z], keys the loader ignores (serde drops unknown fields: "scale" on textures,
"k", "radius", "tube_radius", "shape"/"step" on plain Spheres, ...) are removed, key order is fixed.
dupin.json does not load with the reference's current loader (old schema); it is
restated in the current schema with a camera of our choosing (stated in DESIGN.md).
"""
import json
import os
import sys

REF = "/root/reference/scenes"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes")


def vec(v):
    if isinstance(v, dict):
        return [v["x"], v["y"], v["z"]]
    return list(v)


def tex(t):
    k = t["type"]
    if k == "SolidColor":
        return {"type": k, "color": vec(t["color"])}
    if k == "CheckerTexture":
        return {"type": k, "odd": tex(t["odd"]), "even": tex(t["even"]), "multipliers": vec(t["multipliers"])}
    if k == "UVChecker":
        return {"type": k, "odd": tex(t["odd"]), "even": tex(t["even"]), "multipliers": list(t["multipliers"])}
    if k == "ImageTexture":
        return {"type": k, "image_filename": t["image_filename"]}
    if k == "NoiseTexture":
        return {"type": k, "scale": t["scale"]}
    raise ValueError(k)


def mat(m):
    k = m["type"]
    if k == "Lambertian":
        return {"type": k, "albedo": tex(m["albedo"])}
    if k == "Metal":
        return {"type": k, "albedo": tex(m["albedo"]), "fuzz": m["fuzz"]}
    if k == "Dielectric":
        return {"type": k, "index_of_refraction": m["index_of_refraction"]}
    if k == "DiffuseLight":
        return {"type": k, "emit": tex(m["emit"])}
    return {"type": k}


def xf(t):
    return {"translate": vec(t["translate"]), "rotate": vec(t["rotate"]), "scale": vec(t["scale"])}


def shape(s):
    k = s["type"]
    if k == "Sphere":
        o = {"type": k, "name": s["name"], "transform": xf(s["transform"]), "material": s["material"]}
        if s.get("inverse_normal"):
            o["inverse_normal"] = True
        return o
    if k == "Cube":
        return {"type": k, "name": s["name"], "transform": xf(s["transform"]), "material": s["material"]}
    if k == "Rectangle":
        return {"type": k, "x0": s["x0"], "y0": s["y0"], "x1": s["x1"], "y1": s["y1"],
                "transform": xf(s["transform"]), "material": s["material"]}
    if k == "BruteForsableShape":
        sh = s["shape"]
        keep = {"Heart": [], "Sine": ["a", "sphere_radius"], "Star": ["a", "sphere_radius"],
                "DupinCyclide": ["a", "b", "c", "d", "sphere_radius"], "HuntsSurface": ["sphere_radius"],
                "Cushion": ["sphere_radius"]}[sh["type"]]
        o = {"type": k, "name": s.get("name", sh["type"]), "shape": {"type": sh["type"], **{q: sh[q] for q in keep}},
             "step": s["step"], "transform": xf(s["transform"]), "material": s["material"]}
        if "depth" in s:
            o["depth"] = s["depth"]
        return o
    raise ValueError(k)


def camera(c):
    return {"position": vec(c["position"]), "direction": vec(c["direction"]), "up": vec(c["up"]),
            "fov": c["fov"], "focal_length": c["focal_length"]}


def convert(name):
    d = json.load(open(os.path.join(REF, name)))
    return {"camera": camera(d["camera"]), "background": vec(d["background"]),
            "materials": {k: mat(m) for k, m in d["materials"].items()},
            "shapes": [shape(s) for s in d["shapes"]]}


def dupin():
    """scenes/dupin.json restated in the loader's current schema: the three shapes with
    their transforms / material parameters, inline materials hoisted into the materials map, the two
    `TransformedSphere`s as `Sphere`s.  The old file has no camera: position (0,8,-20) looking at
    (0,3,0), fov 40 is OUR choice."""
    d = json.load(open(os.path.join(REF, "dupin.json")))
    shapes, mats = [], {}
    names = ["Dupin", "Disc", "Ground"]
    for s, nm in zip(d["shapes"], names):
        m = s["material"]
        inner = m["material"]
        albedo = {"type": "SolidColor", "color": vec(inner["albedo"]["color"])}
        if m["type"] == "Metal":
            mats[nm] = {"type": "Metal", "albedo": albedo, "fuzz": inner["fuzz"]}
        elif m["type"] == "Lambertian":
            mats[nm] = {"type": "Lambertian", "albedo": albedo}
        else:
            raise ValueError(m["type"])
        if s["type"] == "BruteForsableShape":
            sh = s["shape"]
            shapes.append({"type": "BruteForsableShape", "name": nm,
                           "shape": {"type": sh["type"], **{q: sh[q] for q in ["a", "b", "c", "d", "sphere_radius"]}},
                           "step": s["step"], "transform": xf(s["transform"]), "material": nm})
        else:
            shapes.append({"type": "Sphere", "name": nm, "transform": xf(s["transform"]), "material": nm})
    cam = {"position": [0.0, 8.0, -20.0], "direction": [0.0, -5.0, 20.0], "up": [0.0, 1.0, 0.0], "fov": 40.0,
           "focal_length": 1.0}
    return {"camera": cam, "background": [0.0, 0.0, 0.0], "materials": mats, "shapes": shapes}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ["spheres.json", "cornell_box.json", "detached_materials.json", "cube_test.json", "empty.json",
                 "light_source.json"]:
        with open(os.path.join(OUT, name), "w") as f:
            json.dump(convert(name), f, indent=1)
            f.write("\n")
    with open(os.path.join(OUT, "dupin.json"), "w") as f:
        json.dump(dupin(), f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    sys.exit(main())
