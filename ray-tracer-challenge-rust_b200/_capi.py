"""ctypes declarations for a library that exports the scene-construction C API of include/rtc.h.

`BuilderApi(lib, prefix)` binds the construction subset (matrices, materials, shapes, worlds, cameras) of a shared
library whose symbols are `<prefix>translation`, `<prefix>shape_new`, ...  The product library uses prefix `rtc_`;
the CPU oracle (oracle/, test infrastructure) deliberately exports the same shapes under `orc_`, so the scene
descriptions in scenes.py can be replayed into it by tests/ and bench.py's cpu_baseline leg.  This module never loads
the oracle itself.
"""
import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_u64_p = C.POINTER(C.c_uint64)


class Material(C.Structure):
    """rtc_material (include/rtc.h) == material.rs:4-14 with its Option<Pattern> (pattern.rs:14-19) flattened in."""
    _fields_ = [
        ("color", C.c_double * 3),
        ("ambient", C.c_double), ("diffuse", C.c_double), ("specular", C.c_double), ("shininess", C.c_double),
        ("reflective", C.c_double), ("transparency", C.c_double), ("refractive_index", C.c_double),
        ("pattern_kind", C.c_int32), ("_pad", C.c_int32),
        ("pattern_a", C.c_double * 3), ("pattern_b", C.c_double * 3),
        ("pattern_transform", C.c_double * 16), ("pattern_inverse", C.c_double * 16),
    ]


class CameraDesc(C.Structure):
    """rtc_camera_desc"""
    _fields_ = [("hsize", C.c_uint32), ("vsize", C.c_uint32), ("inverse", C.c_double * 16),
                ("half_width", C.c_double), ("half_height", C.c_double), ("pixel_size", C.c_double)]


class Rows(C.Structure):
    """rtc_rows"""
    _fields_ = [("band_rows", C.c_uint32), ("band_first", C.c_uint32), ("band_stride", C.c_uint32),
                ("layout", C.c_uint32)]
    COMPACT, FRAME = 0, 1


class Stats(C.Structure):
    """rtc_stats"""
    _fields_ = [("primary_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("reflect_rays", C.c_uint64),
                ("refract_rays", C.c_uint64), ("kernel_launches", C.c_uint64), ("device_ms", C.c_double)]

    @property
    def total_rays(self):
        return self.primary_rays + self.shadow_rays + self.reflect_rays + self.refract_rays


class Computations(C.Structure):
    """rtc_computations == Computations (intersection.rs:88-100) + Computations::schlick of the hit of one ray"""
    _fields_ = [("hit", C.c_int32), ("leaf", C.c_int32), ("inside", C.c_int32), ("_pad", C.c_int32), ("t", C.c_double),
                ("point", C.c_double * 3), ("eyev", C.c_double * 3), ("normalv", C.c_double * 3),
                ("reflectv", C.c_double * 3), ("over_point", C.c_double * 3), ("under_point", C.c_double * 3),
                ("n1", C.c_double), ("n2", C.c_double), ("reflectance", C.c_double)]


def as_f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} doubles, got {a.size}")
    return a


def dptr(a):
    return a.ctypes.data_as(c_double_p)


class BuilderApi:
    """The construction subset shared by librtc_b200.so (`rtc_`) and the oracle (`orc_`)."""

    SPHERE, PLANE, CUBE, CYLINDER, CONE, GROUP, TRIANGLE, SMOOTH_TRIANGLE = range(8)
    PATTERN_NONE, PATTERN_STRIPE, PATTERN_GRADIENT, PATTERN_RING, PATTERN_CHECKERS, PATTERN_TEST = -1, 0, 1, 2, 3, 4

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        f = self._fn
        vp = C.c_void_p
        f("last_error", C.c_char_p)
        f("translation", None, C.c_double, C.c_double, C.c_double, c_double_p)
        f("scaling", None, C.c_double, C.c_double, C.c_double, c_double_p)
        f("rotation_x", None, C.c_double, c_double_p)
        f("rotation_y", None, C.c_double, c_double_p)
        f("rotation_z", None, C.c_double, c_double_p)
        f("shearing", None, *([C.c_double] * 6), c_double_p)
        f("view_transform", C.c_int, c_double_p, c_double_p, c_double_p, c_double_p)
        f("matrix_mul", None, c_double_p, c_double_p, c_double_p)
        f("matrix_transpose", None, c_double_p, c_double_p)
        f("matrix_inverse", C.c_int, c_double_p, c_double_p)
        f("matrix_mul_tuple", None, c_double_p, c_double_p, c_double_p)
        f("material_default", None, C.POINTER(Material))
        f("material_set_pattern_transform", C.c_int, C.POINTER(Material), c_double_p)
        f("shape_new", vp, C.c_int, C.c_double, C.c_double, C.c_int)
        f("shape_triangle", vp, c_double_p, c_double_p, c_double_p)
        f("shape_smooth_triangle", vp, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p)
        f("shape_free", None, vp)
        f("shape_set_transform", C.c_int, vp, c_double_p)
        f("shape_set_material", C.c_int, vp, C.POINTER(Material))
        f("shape_push_shape", C.c_int, vp, vp)
        f("shape_leaf_count", C.c_uint64, vp)
        f("obj_parse_file", vp, C.c_char_p, c_u64_p)
        f("obj_parse_str", vp, C.c_char_p, C.c_uint64, c_u64_p)
        f("mesh_from_arrays", vp, c_double_p, C.c_uint64, C.POINTER(C.c_int32), C.c_uint64)
        f("smooth_mesh_from_arrays", vp, c_double_p, C.c_uint64, c_double_p, C.c_uint64, C.POINTER(C.c_int32),
          C.POINTER(C.c_int32), C.c_uint64)
        f("world_new", vp, c_double_p, c_double_p)
        f("world_default", vp)
        f("world_free", None, vp)
        f("world_push", C.c_int, vp, vp)
        f("world_set_recursion_limit", C.c_int, vp, C.c_uint32)
        f("camera_new", vp, C.c_uint64, C.c_uint64, C.c_double)
        f("camera_free", None, vp)
        f("camera_set_transform", C.c_int, vp, c_double_p)

    def _fn(self, name, restype, *argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        fn.argtypes = list(argtypes)
        setattr(self, name, fn)
        return fn

    def error(self):
        m = self.last_error()
        return m.decode("utf-8", "replace") if m else ""


# ---- layer-1 description structs (rtc_scene_desc and friends) ---------------------------------------------------------
class TransformDesc(C.Structure):
    _fields_ = [("transform", C.c_double * 16), ("inverse", C.c_double * 16)]


class TriangleDesc(C.Structure):
    _fields_ = [("p1", C.c_double * 3), ("p2", C.c_double * 3), ("p3", C.c_double * 3), ("e1", C.c_double * 3),
                ("e2", C.c_double * 3), ("normal", C.c_double * 3)]


class VertexNormals(C.Structure):
    _fields_ = [("n1", C.c_double * 3), ("n2", C.c_double * 3), ("n3", C.c_double * 3)]


class ShapeDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("material", C.c_int32), ("transform", C.c_int32), ("capped", C.c_int32),
                ("minimum", C.c_double), ("maximum", C.c_double), ("child_count", C.c_int32), ("triangle", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("shapes", C.POINTER(ShapeDesc)), ("shape_count", C.c_uint32), ("root_count", C.c_uint32),
                ("transforms", C.POINTER(TransformDesc)), ("transform_count", C.c_uint32),
                ("materials", C.POINTER(Material)), ("material_count", C.c_uint32),
                ("triangles", C.POINTER(TriangleDesc)), ("triangle_count", C.c_uint32),
                ("light_position", C.c_double * 3), ("light_intensity", C.c_double * 3),
                ("vertex_normals", C.POINTER(VertexNormals)), ("recursion_limit", C.c_uint32), ("_reserved", C.c_uint32)]
