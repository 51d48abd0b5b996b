"""rtc_b200 — B200-native per-pixel render path for antoinehebert/ray-tracer-challenge-rust.

The product is librtc_b200.so (CUDA sm_100a behind the C ABI of include/rtc.h); this package is its Python host: the
reference's `World / Camera / Canvas / to_ppm` API (camera.rs, world.rs, canvas.rs) bound with ctypes.  There is no CPU
rendering path: without the built library the import fails, and without a CUDA device every render call raises
RtcError(RTC_ERR_CUDA).

The directory name contains hyphens, so import it with importlib (tests/conftest.py and bench.py do):
    rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
"""
import ctypes as C

import numpy as np

from . import scenes  # noqa: F401
from ._capi import BuilderApi, CameraDesc, Rows, Stats, as_f64, dptr
from ._lib import (LIB_PATH, RTC_BUILD_DEVICE_LBVH, RTC_BUILD_HOST_SAH, RTC_ERR_CUDA, RTC_ERR_INVALID, RTC_ERR_PANIC, RTC_ERR_TIMEOUT, RTC_ERR_UNSUPPORTED, RTC_OK, RtcError,
                   api)
from .scene_api import (BLACK, BLUE, GREEN, RED, WHITE, CameraHandle, Light, Material, Matrix, Pattern, Shape, Shapes,
                        Transformations, WorldHandle)

__all__ = ["api", "World", "Camera", "Canvas", "Light", "Material", "Pattern", "Shapes", "Transformations", "Matrix",
           "Rows", "Stats", "RtcError", "scenes", "device_count", "measure_fp64_peak", "ppm_from_rgba8", "ppm_from_device", "MultiRenderer", "render_multi"]


def device_count():
    return api().device_count()


def measure_fp64_peak(device=0):
    """Self-measured FP64 issue peaks in Gflop/s: (DMUL+DADD chains without FMA, DFMA chains counted as 2)."""
    a, b = C.c_double(0), C.c_double(0)
    api().check(api().measure_fp64_peak(device, C.byref(a), C.byref(b)))
    return a.value, b.value


def ppm_from_rgba8(rgba8, width, height):
    """Canvas::to_ppm (canvas.rs:28-58) over an RGBA8 frame -> bytes."""
    a = np.ascontiguousarray(rgba8, dtype=np.uint8)
    if a.size != width * height * 4:
        raise ValueError("rgba8 has the wrong size")
    n = C.c_uint64(0)
    p = api().ppm_from_rgba8(a.ctypes.data_as(C.c_void_p), width, height, C.byref(n))
    try:
        return C.string_at(p, n.value)
    finally:
        api().free(p)


def ppm_from_device(d_rgba8, width, height, device=0, stream=0):
    """Canvas::to_ppm on the GPU for an RGBA8 frame that lives in device memory (d_rgba8: device pointer as int, e.g.
    tensor.data_ptr()) -> bytes.  Only the text crosses PCIe."""
    a = api()
    cap = a.ppm_max_bytes(width, height)
    buf = a.pinned_alloc(cap)
    if not buf:
        raise RtcError(RTC_ERR_CUDA, "cannot allocate pinned memory: " + a.error())
    try:
        n = C.c_uint64(0)
        a.check(a.ppm_encode_device(device, C.c_void_p(d_rgba8), width, height, C.c_void_p(stream) if stream else None,
                                    C.c_void_p(buf), cap, C.byref(n)))
        return C.string_at(buf, n.value)
    finally:
        a.pinned_free(buf)


class Canvas:
    """Canvas (canvas.rs:5-63)."""

    def __init__(self, width=None, height=None, _handle=None):
        self.api = api()
        self.h = _handle if _handle is not None else self.api.canvas_new(int(width), int(height))
        self.width, self.height = self.api.canvas_width(self.h), self.api.canvas_height(self.h)

    def get_pixel(self, x, y):
        out = np.empty(3)
        self.api.check(self.api.canvas_get_pixel(self.h, x, y, dptr(out)))
        return out

    def set_pixel(self, x, y, color):
        c = as_f64(color, 3)
        self.api.check(self.api.canvas_set_pixel(self.h, x, y, dptr(c)))

    def pixels_f64(self):
        """(height, width, 3) float64 copy of the Canvas colours, or None (rendered with want_f64=False)."""
        p = self.api.canvas_pixels_f64(self.h)
        if not p:
            return None
        return np.ctypeslib.as_array(p, shape=(self.height, self.width, 3)).copy()

    def pixels_rgba8(self):
        p = self.api.canvas_pixels_rgba8(self.h)
        return np.ctypeslib.as_array(p, shape=(self.height, self.width, 4)).copy()

    def to_ppm(self):
        n = C.c_uint64(0)
        p = self.api.canvas_to_ppm(self.h, C.byref(n))
        try:
            return C.string_at(p, n.value)
        finally:
            self.api.free(p)

    def __del__(self):
        if getattr(self, "h", None) is not None:
            self.api.canvas_free(self.h)
            self.h = None


class World(WorldHandle):
    """World (world.rs:13-98) on the GPU."""

    def __init__(self, light=None, _handle=None):
        super().__init__(api(), light, _handle)

    @staticmethod
    def default_world():  # world.rs:26-41
        return World(_handle=api().world_default())

    def color_at(self, rays):
        """World::color_at (world.rs:80-82) for an (n, 6) array of rays (origin xyz, direction xyz) -> (n, 3)."""
        r = as_f64(rays).reshape(-1, 6)
        out = np.empty((r.shape[0], 3))
        self.api.check(self.api.world_color_at(self.h, dptr(r), r.shape[0], dptr(out)))
        return out

    def intersect(self, rays, cap=16, device=0):
        """World::intersect (world.rs:43-54) per ray -> list of [(t, leaf), ...] in the reference's sorted order."""
        r = as_f64(rays).reshape(-1, 6)
        n = r.shape[0]
        t = np.empty((n, cap))
        leaf = np.empty((n, cap), dtype=np.int32)
        cnt = np.empty(n, dtype=np.uint32)
        self.api.check(self.api.intersect(self.scene(device), dptr(r), n, cap, dptr(t),
                                          leaf.ctypes.data_as(C.POINTER(C.c_int32)),
                                          cnt.ctypes.data_as(C.POINTER(C.c_uint32))))
        if n and int(cnt.max()) > cap:
            return self.intersect(rays, int(cnt.max()), device)
        return [[(float(t[i, k]), int(leaf[i, k])) for k in range(int(cnt[i]))] for i in range(n)]

    def prepare_computations(self, rays, device=0):
        """Intersection::hit + prepare_computations + schlick (intersection.rs:17-128) for the hit of each ray."""
        from ._capi import Computations
        r = as_f64(rays).reshape(-1, 6)
        out = (Computations * r.shape[0])()
        self.api.check(self.api.prepare_computations(self.scene(device), dptr(r), r.shape[0], out))
        return list(out)

    def normal_at(self, leaf, points, device=0):
        """Shape::normal_at (shape.rs:466-519) of the leaf-th leaf (pre-order) at world points -> (n, 3)."""
        p = as_f64(points).reshape(-1, 3)
        out = np.empty_like(p)
        self.api.check(self.api.normal_at(self.scene(device), leaf, dptr(p), p.shape[0], dptr(out)))
        return out

    def set_build(self, build):
        """Mesh build of this world's scene: "host" (binned SAH, the default — best for many frames of one scene) or
        "device" (linear BVH built on the GPU — best when a scene is built, rendered once and dropped)."""
        flags = {"host": RTC_BUILD_HOST_SAH, "device": RTC_BUILD_DEVICE_LBVH}[build]
        self.api.check(self.api.world_set_build(self.h, flags))

    def set_recursion_limit(self, limit):
        """World's RECURSION_LIMIT (world.rs:11): 0 / 5 / 6 one bounce (the reference), 2-3 none, 8-9 two, ..."""
        self.api.check(self.api.world_set_recursion_limit(self.h, int(limit)))

    def drop_scenes(self):
        """Forget the uploaded scenes: the next render marshals, flattens and uploads again."""
        self.api.world_drop_scenes(self.h)

    def scene(self, device=0):
        """The layer-1 rtc_scene handle (flattened + uploaded on first use; owned by the world)."""
        s = C.c_void_p()
        self.api.check(self.api.world_scene(self.h, device, C.byref(s)))
        return s

    def scene_info(self, device=0):
        n = (C.c_uint64 * 6)()
        self.api.check(self.api.scene_info(self.scene(device), n))
        return dict(zip(("leaves", "gates", "meshes", "mesh_triangles", "bvh_nodes", "device_bytes"), list(n)))

    def flatten_info(self, want_gates=False):
        """Host-only flattening facts (no device needed)."""
        n = (C.c_uint64 * 8)()
        self.api.check(self.api.world_flatten_info(self.h, n, None, 0))
        info = dict(zip(("leaves", "gates", "meshes", "mesh_triangles", "bvh_nodes", "bvh_max_depth", "program_nodes",
                         "transforms"), list(n)))
        if want_gates:
            g = np.empty((info["gates"], 6))
            self.api.check(self.api.world_flatten_info(self.h, n, dptr(g), info["gates"]))
            info["gate_boxes"] = g
        return info

    def kernel_features(self):
        """(feature mask of the flattened world, feature mask of the render kernel instantiation chosen for it)."""
        f = (C.c_uint32 * 2)()
        self.api.check(self.api.world_kernel_features(self.h, f))
        return int(f[0]), int(f[1])


class Camera(CameraHandle):
    """Camera (camera.rs:5-79); render() is the drop-in for camera.rs:67-79."""

    def __init__(self, hsize, vsize, field_of_view):
        super().__init__(api(), hsize, vsize, field_of_view)

    def desc(self):
        d = CameraDesc()
        self.api.camera_desc_get(self.h, C.byref(d))
        return d

    def render(self, world, want_f64=True, stats=None, device=0):
        """Camera::render(&World) -> Canvas.  `stats` (a Stats) receives ray counts and the kernel's device time.
        device=ALL_DEVICES (-1, RTC_DEVICE_ALL): the frame is sharded over every CUDA device of this process."""
        out = C.c_void_p()
        st = stats if stats is not None else None
        self.api.check(self.api.camera_render(self.h, world.h, device, int(bool(want_f64)), C.byref(out),
                                              C.byref(st) if st is not None else None))
        return Canvas(_handle=out)

    def render_into(self, world, rgba8=None, rgb_f64=None, rows=None, stats=None, device=0):
        """rtc_render with caller-owned HOST arrays (numpy; pinned torch tensors work through .numpy()):
        rgba8 (local_rows, hsize, 4) uint8 and/or rgb_f64 (local_rows, hsize, 3) float64."""
        d = self.desc()
        p8 = rgba8.ctypes.data_as(C.c_void_p) if rgba8 is not None else None
        p64 = rgb_f64.ctypes.data_as(C.c_void_p) if rgb_f64 is not None else None
        self.api.check(self.api.render(world.scene(device), C.byref(d), C.byref(rows) if rows is not None else None,
                                       p8, p64, C.byref(stats) if stats is not None else None))

    def render_device(self, world, d_rgba8=None, d_rgb_f64=None, rows=None, stream=0, stats=None, device=0):
        """rtc_render_device: DEVICE pointers (ints, e.g. torch tensor .data_ptr()), asynchronous on `stream` (a
        cudaStream_t as int).  With `stats` the call synchronises the stream and fills it."""
        d = self.desc()
        scene = world.scene(device) if hasattr(world, "scene") else world  # a World, or a raw rtc_scene handle
        self.api.check(self.api.render_device(scene, C.byref(d),
                                              C.byref(rows) if rows is not None else None,
                                              C.c_void_p(d_rgba8) if d_rgba8 else None,
                                              C.c_void_p(d_rgb_f64) if d_rgb_f64 else None,
                                              C.c_void_p(stream) if stream else None,
                                              1 if stats is not None else 0,
                                              C.byref(stats) if stats is not None else None))

    def rows_count(self, rows=None):
        d = self.desc()
        return self.api.rows_count(C.byref(d), C.byref(rows) if rows is not None else None)


ALL_DEVICES = -1  # RTC_DEVICE_ALL


class MultiRenderer:
    """rtc_multi_*: one frame sharded over GPUs 0..ngpus-1 of THIS process (no torch, no NCCL).  where="host": every device
    copies its own row bands into one pinned host frame over its own PCIe link; where="device": kernels store straight into
    a frame on device 0 over NVLink peer mappings."""

    def __init__(self, world, ngpus, build="host"):
        self.api = api()
        m = C.c_void_p()
        self.api.check(self.api.world_marshal(world.h, C.byref(m)))
        try:
            self.h = C.c_void_p()
            flags = {"host": RTC_BUILD_HOST_SAH, "device": RTC_BUILD_DEVICE_LBVH}[build]
            self.api.check(self.api.multi_create(self.api.marshalled_desc(m), ngpus, flags, C.byref(self.h)))
        finally:
            self.api.marshalled_free(m)
        self.ngpus = ngpus

    def render(self, camera, where="host", stats=None, copy=True):
        """-> (vsize, hsize, 4) uint8 numpy array (where="host": a copy of the pinned frame, or a view of it with
        copy=False, valid until the next render) or the device-0 pointer of the frame (where="device")."""
        d = camera.desc()
        code = {"host": 0, "device": 1}[where]
        self.api.check(self.api.multi_render(self.h, C.byref(d), code, None, C.byref(stats) if stats is not None else None))
        if where == "device":
            return self.api.multi_device_frame(self.h)
        p = C.cast(self.api.multi_host_frame(self.h), C.POINTER(C.c_uint8))
        a = np.ctypeslib.as_array(p, shape=(camera.vsize, camera.hsize, 4))
        return a.copy() if copy else a

    def render_into(self, camera, rgba8=None, rgb_f64=None, stats=None):
        """rtc_multi_render_host: the sharded frame straight into caller-owned HOST arrays — rgba8 (vsize, hsize, 4) uint8
        and / or rgb_f64 (vsize, hsize, 3) float64; every device copies its own bands to their frame positions."""
        d = camera.desc()
        p8 = rgba8.ctypes.data_as(C.c_void_p) if rgba8 is not None else None
        p64 = rgb_f64.ctypes.data_as(C.c_void_p) if rgb_f64 is not None else None
        self.api.check(self.api.multi_render_host(self.h, C.byref(d), p8, p64, C.byref(stats) if stats is not None else None))

    def close(self):
        if getattr(self, "h", None):
            self.api.multi_destroy(self.h)
            self.h = None

    __del__ = close


def render_multi(world, camera, ngpus, build="device", stats=None):
    """rtc_render_multi: marshal, upload to every device, render sharded, frame to host, drop — one call."""
    a = api()
    m = C.c_void_p()
    a.check(a.world_marshal(world.h, C.byref(m)))
    try:
        out = np.empty((camera.vsize, camera.hsize, 4), dtype=np.uint8)
        d = camera.desc()
        flags = {"host": RTC_BUILD_HOST_SAH, "device": RTC_BUILD_DEVICE_LBVH}[build]
        a.check(a.render_multi(a.marshalled_desc(m), C.byref(d), ngpus, flags, out.ctypes.data_as(C.c_void_p),
                               C.byref(stats) if stats is not None else None))
        return out
    finally:
        a.marshalled_free(m)


def build_scene(name, hsize=None, vsize=None):
    """A named BASELINE config as (World, Camera) on the product library."""
    w, c = scenes.build(api(), name, hsize, vsize)
    world = World(_handle=w.h)
    w.h = None
    world.light = w.light
    cam = Camera.__new__(Camera)
    cam.api, cam.hsize, cam.vsize, cam.field_of_view, cam.h = c.api, c.hsize, c.vsize, c.field_of_view, c.h
    c.h = None
    return world, cam
