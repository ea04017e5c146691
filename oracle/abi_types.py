"""ORACLE — TEST INFRASTRUCTURE ONLY.  The plain-data types of include/rt_b200.h as ctypes structures.

A second copy of the declarations in rs_pathtracing_b200/_ffi.py (data only, no code), so that the reference arm
of bench.py (`--impl reference`) can run the oracle on a flattened scene WITHOUT importing the product package or
loading its libraries.  tests/test_host.py checks that the two copies agree field by field.
"""
import ctypes as C


class Vec3(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]

    def __init__(self, x=0.0, y=0.0, z=0.0):
        super().__init__(float(x), float(y), float(z))

    def tuple(self):
        return (self.x, self.y, self.z)


class Ray(C.Structure):
    _fields_ = [("origin", Vec3), ("direction", Vec3)]


class Camera(C.Structure):
    _fields_ = [("position", Vec3), ("direction", Vec3), ("up", Vec3), ("right", Vec3),
                ("fov_rad", C.c_double), ("focal_length", C.c_double)]


class ImageParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("texture", C.c_uint32), ("scalar", C.c_double)]


class Texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("odd", C.c_uint32), ("even", C.c_uint32), ("image", C.c_uint32),
                ("color", Vec3)]


class Image(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgba", C.POINTER(C.c_uint8))]


class Perlin(C.Structure):
    _fields_ = [("perm_x", C.c_uint32 * 256), ("perm_y", C.c_uint32 * 256), ("perm_z", C.c_uint32 * 256),
                ("ranvec", Vec3 * 256)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_shapes", C.c_uint32),
        ("kind", C.POINTER(C.c_uint8)),
        ("flags", C.POINTER(C.c_uint8)),
        ("inverse", C.POINTER(C.c_double)),
        ("direct", C.POINTER(C.c_double)),
        ("params", C.POINTER(C.c_double)),
        ("material", C.POINTER(C.c_uint32)),
        ("n_materials", C.c_uint32),
        ("materials", C.POINTER(Material)),
        ("n_textures", C.c_uint32),
        ("textures", C.POINTER(Texture)),
        ("n_images", C.c_uint32),
        ("images", C.POINTER(Image)),
        ("n_noise", C.c_uint32),
        ("noise", C.POINTER(Perlin)),
    ]
