"""The reference's host API (Matrix / Material / Pattern / Shape / World / Camera), generic over a BuilderApi.

Class and method names follow the Rust sources (shape.rs, material.rs, pattern.rs, world.rs, camera.rs,
transformations.rs) so scene descriptions read like the reference's own `main.rs`.  All arithmetic is done by the bound C
library (the product's host mirror, or — from tests/ and bench.py's CPU leg only — the oracle), never by Python/numpy,
so both sides see bit-identical inputs.
"""
import ctypes as C

import numpy as np

from ._capi import BuilderApi, Material as CMaterial, as_f64, dptr

BLACK, WHITE = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
RED, GREEN, BLUE = (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)


class Matrix:
    """Matrix<4> (matrix.rs:6-227), row-major."""

    def __init__(self, api, values):
        self.api = api
        self.v = as_f64(values, 16).copy()

    def __mul__(self, other):  # matrix.rs:187-205
        out = np.empty(16)
        self.api.matrix_mul(dptr(self.v), dptr(other.v), dptr(out))
        return Matrix(self.api, out)

    __matmul__ = __mul__

    def inverse(self):  # matrix.rs:138-157
        out = np.empty(16)
        if self.api.matrix_inverse(dptr(self.v), dptr(out)) != 0:
            raise ValueError(self.api.error())
        return Matrix(self.api, out)

    def transpose(self):
        out = np.empty(16)
        self.api.matrix_transpose(dptr(self.v), dptr(out))
        return Matrix(self.api, out)

    def mul_tuple(self, t4):  # matrix.rs:207-227
        t = as_f64(t4, 4)
        out = np.empty(4)
        self.api.matrix_mul_tuple(dptr(self.v), dptr(t), dptr(out))
        return out

    def array(self):
        return self.v.reshape(4, 4).copy()


class Transformations:
    """transformations.rs:4-93 bound to one library."""

    def __init__(self, api):
        self.api = api

    def _m(self, fn, *args):
        out = np.empty(16)
        fn(*args, dptr(out))
        return Matrix(self.api, out)

    def identity(self):
        return Matrix(self.api, np.eye(4).ravel())

    def translation(self, x, y, z):
        return self._m(self.api.translation, x, y, z)

    def scaling(self, x, y, z):
        return self._m(self.api.scaling, x, y, z)

    def rotation_x(self, r):
        return self._m(self.api.rotation_x, r)

    def rotation_y(self, r):
        return self._m(self.api.rotation_y, r)

    def rotation_z(self, r):
        return self._m(self.api.rotation_z, r)

    def shearing(self, xy, xz, yx, yz, zx, zy):
        return self._m(self.api.shearing, xy, xz, yx, yz, zx, zy)

    def view_transform(self, frm, to, up):
        out = np.empty(16)
        a, b, c = as_f64(frm, 3), as_f64(to, 3), as_f64(up, 3)
        if self.api.view_transform(dptr(a), dptr(b), dptr(c), dptr(out)) != 0:
            raise ValueError(self.api.error())
        return Matrix(self.api, out)


class Pattern:
    """pattern.rs:14-66"""

    def __init__(self, kind, a=BLACK, b=BLACK):
        self.kind, self.a, self.b, self.transform = kind, tuple(a), tuple(b), None

    @staticmethod
    def stripe(a, b):
        return Pattern(BuilderApi.PATTERN_STRIPE, a, b)

    @staticmethod
    def gradient(a, b):
        return Pattern(BuilderApi.PATTERN_GRADIENT, a, b)

    @staticmethod
    def ring(a, b):
        return Pattern(BuilderApi.PATTERN_RING, a, b)

    @staticmethod
    def checkers(a, b):
        return Pattern(BuilderApi.PATTERN_CHECKERS, a, b)

    @staticmethod
    def test_pattern():
        return Pattern(BuilderApi.PATTERN_TEST)

    def set_transform(self, m):
        self.transform = m


class Material:
    """material.rs:4-29"""

    def __init__(self):
        self.color = WHITE
        self.ambient, self.diffuse, self.specular, self.shininess = 0.1, 0.9, 0.9, 200.0
        self.reflective, self.transparency, self.refractive_index = 0.0, 0.0, 1.0
        self.pattern = None

    def to_c(self, api):
        m = CMaterial()
        api.material_default(C.byref(m))
        m.color[:] = self.color
        m.ambient, m.diffuse, m.specular, m.shininess = self.ambient, self.diffuse, self.specular, self.shininess
        m.reflective, m.transparency, m.refractive_index = self.reflective, self.transparency, self.refractive_index
        if self.pattern is not None:
            m.pattern_kind = self.pattern.kind
            m.pattern_a[:] = self.pattern.a
            m.pattern_b[:] = self.pattern.b
            if self.pattern.transform is not None:
                if api.material_set_pattern_transform(C.byref(m), dptr(self.pattern.transform.v)) != 0:
                    raise ValueError(api.error())
        return m


class Shape:
    """shape.rs:42-245.  Owns a C handle until pushed into a group or a world (which consume it)."""

    def __init__(self, api, handle, kind):
        if not handle:
            raise ValueError(api.error())
        self.api, self.h, self.kind = api, handle, kind
        self.material = Material()  # get_material_mut(): edited in place, sent at push time for leaves

    def _live(self):
        if self.h is None:
            raise ValueError("shape was moved into a group or world")
        return self.h

    def set_transform(self, m):  # shape.rs:196-218
        if self.api.shape_set_transform(self._live(), dptr(m.v)) != 0:
            raise ValueError(self.api.error())

    def set_material(self, material):  # shape.rs:220-229 (push-down through groups)
        self.material = material
        cm = material.to_c(self.api)
        if self.api.shape_set_material(self._live(), C.byref(cm)) != 0:
            raise ValueError(self.api.error())

    def get_material_mut(self):
        return self.material

    def push_shape(self, child):  # shape.rs:528-535
        child._commit_material()
        if self.api.shape_push_shape(self._live(), child._live()) != 0:
            raise ValueError(self.api.error())
        child.h = None

    def _commit_material(self):
        if self.kind != BuilderApi.GROUP:
            cm = self.material.to_c(self.api)
            self.api.shape_set_material(self._live(), C.byref(cm))

    def leaf_count(self):
        return self.api.shape_leaf_count(self._live())

    def __del__(self):
        if getattr(self, "h", None) is not None:
            self.api.shape_free(self.h)
            self.h = None


class Shapes:
    """Shape constructors (shape.rs:52-193, obj_file.rs:23-128) bound to one library."""

    def __init__(self, api):
        self.api = api

    def _new(self, kind, mn=0.0, mx=0.0, capped=False):
        return Shape(self.api, self.api.shape_new(kind, mn, mx, int(capped)), kind)

    def sphere(self):
        return self._new(BuilderApi.SPHERE)

    def glass_sphere(self):  # shape.rs:63-76
        s = self.sphere()
        s.material.transparency, s.material.refractive_index = 1.0, 1.5
        return s

    def plane(self):
        return self._new(BuilderApi.PLANE)

    def cube(self):
        return self._new(BuilderApi.CUBE)

    def cylinder(self, minimum=-np.inf, maximum=np.inf, capped=False):
        return self._new(BuilderApi.CYLINDER, minimum, maximum, capped)

    def cone(self, minimum=-np.inf, maximum=np.inf, capped=False):
        return self._new(BuilderApi.CONE, minimum, maximum, capped)

    def group(self):
        return self._new(BuilderApi.GROUP)

    def triangle(self, p1, p2, p3):
        a, b, c = as_f64(p1, 3), as_f64(p2, 3), as_f64(p3, 3)
        return Shape(self.api, self.api.shape_triangle(dptr(a), dptr(b), dptr(c)), BuilderApi.TRIANGLE)

    def smooth_triangle(self, p1, p2, p3, n1, n2, n3):
        """The book's smooth_triangle (not in the reference; rtc.h RTC_SMOOTH_TRIANGLE)."""
        a = [as_f64(x, 3) for x in (p1, p2, p3, n1, n2, n3)]
        return Shape(self.api, self.api.shape_smooth_triangle(*[dptr(x) for x in a]), BuilderApi.SMOOTH_TRIANGLE)

    def obj_file(self, path):
        """Parser::from_obj_file(path).obj_to_group()"""
        ign = C.c_uint64(0)
        s = Shape(self.api, self.api.obj_parse_file(str(path).encode(), C.byref(ign)), BuilderApi.GROUP)
        s.ignored_lines = ign.value
        return s

    def obj_str(self, text):
        b = text.encode() if isinstance(text, str) else bytes(text)
        ign = C.c_uint64(0)
        s = Shape(self.api, self.api.obj_parse_str(b, len(b), C.byref(ign)), BuilderApi.GROUP)
        s.ignored_lines = ign.value
        return s

    def mesh(self, vertices, faces):
        """group{ default_group{ triangles } } — what obj_to_group() returns for a `v`/`f`-only OBJ."""
        v = as_f64(vertices)
        f = np.ascontiguousarray(faces, dtype=np.int32)
        h = self.api.mesh_from_arrays(dptr(v), v.size // 3, f.ctypes.data_as(C.POINTER(C.c_int32)), f.size // 3)
        return Shape(self.api, h, BuilderApi.GROUP)


    def smooth_mesh(self, vertices, normals, faces, face_normals=None):
        """The same for `vn` records and `f v//n` faces (face_normals defaults to faces: one normal per vertex)."""
        v, n = as_f64(vertices), as_f64(normals)
        f = np.ascontiguousarray(faces, dtype=np.int32)
        fn = f if face_normals is None else np.ascontiguousarray(face_normals, dtype=np.int32)
        i32p = C.POINTER(C.c_int32)
        h = self.api.smooth_mesh_from_arrays(dptr(v), v.size // 3, dptr(n), n.size // 3, f.ctypes.data_as(i32p),
                                             fn.ctypes.data_as(i32p), f.size // 3)
        return Shape(self.api, h, BuilderApi.GROUP)


class Light:
    def __init__(self, position, intensity):
        self.position, self.intensity = tuple(position), tuple(intensity)


class WorldHandle:
    """World (world.rs:13-24): `objects.push` consumes the shape."""

    def __init__(self, api, light=None, handle=None):
        self.api = api
        if handle is None:
            p, i = as_f64(light.position, 3), as_f64(light.intensity, 3)
            handle = api.world_new(dptr(p), dptr(i))
        self.h = handle
        self.light = light

    def push(self, shape):
        shape._commit_material()
        if self.api.world_push(self.h, shape._live()) != 0:
            raise ValueError(self.api.error())
        shape.h = None

    def set_recursion_limit(self, limit):
        """world.rs:11 RECURSION_LIMIT for this world (0 = the reference's 5)."""
        if self.api.world_set_recursion_limit(self.h, int(limit)) != 0:
            raise ValueError(self.api.error())

    def __del__(self):
        if getattr(self, "h", None) is not None:
            self.api.world_free(self.h)
            self.h = None


class CameraHandle:
    """Camera (camera.rs:5-46)."""

    def __init__(self, api, hsize, vsize, field_of_view):
        self.api, self.hsize, self.vsize, self.field_of_view = api, int(hsize), int(vsize), float(field_of_view)
        self.h = api.camera_new(self.hsize, self.vsize, self.field_of_view)

    def set_transform(self, m):
        if self.api.camera_set_transform(self.h, dptr(m.v)) != 0:
            raise ValueError(self.api.error())

    def __del__(self):
        if getattr(self, "h", None) is not None:
            self.api.camera_free(self.h)
            self.h = None
