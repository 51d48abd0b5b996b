"""Test-side bindings: the CPU oracle (oracle/, the parity checker) and the host simulation of the device program
(tests/hostsim).  Both are TEST INFRASTRUCTURE; the product package never imports this module."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
SIM_SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
SIM_LIB = os.path.join(ROOT, "tests", "_build", "librtc_hostsim.so")

_pkg = importlib.import_module("ray-tracer-challenge-rust_b200")
_capi = importlib.import_module("ray-tracer-challenge-rust_b200._capi")
scenes = _pkg.scenes


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


class OracleCounters(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("shadow", C.c_uint64), ("reflect", C.c_uint64), ("refract", C.c_uint64),
                ("leaf_tests", C.c_uint64), ("seconds", C.c_double)]

    @property
    def total_rays(self):
        return self.primary + self.shadow + self.reflect + self.refract


class OracleApi(_capi.BuilderApi):
    FAITHFUL, CACHED = 0, 1

    def __init__(self, path):
        super().__init__(C.CDLL(path), "orc_")
        f, vp = self._fn, C.c_void_p
        f("world_color_at", C.c_int, vp, C.c_int, _capi.c_double_p, C.c_uint64, _capi.c_double_p)
        f("camera_params", None, vp, _capi.c_double_p, _capi.c_double_p, _capi.c_double_p, _capi.c_double_p)
        f("camera_ray_for_pixel", None, vp, C.c_uint64, C.c_uint64, _capi.c_double_p, _capi.c_double_p)
        f("camera_render", C.c_int, vp, vp, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_uint64, _capi.c_double_p,
          C.POINTER(OracleCounters))
        f("quantise", None, _capi.c_double_p, C.c_uint64, C.POINTER(C.c_uint8))
        f("to_ppm", vp, _capi.c_double_p, C.c_uint64, C.c_uint64, _capi.c_u64_p)
        f("free", None, vp)

    def render(self, world, cam, mode=1, nthreads=None, pixels=None):
        """Camera::render on the oracle -> (rgb f64 [n,3], counters).  pixels: (n,2) uint32 x,y list or None."""
        nthreads = nthreads or (os.cpu_count() or 1)
        if pixels is None:
            n, pp = cam.hsize * cam.vsize, None
        else:
            pixels = np.ascontiguousarray(pixels, dtype=np.uint32)
            n, pp = pixels.shape[0], pixels.ctypes.data_as(C.POINTER(C.c_uint32))
        out = np.empty((n, 3))
        cnt = OracleCounters()
        if self.camera_render(cam.h, world.h, mode, nthreads, pp, n, _capi.dptr(out), C.byref(cnt)) != 0:
            raise RuntimeError("oracle: " + self.error())
        return out, cnt

    def color_at(self, world, rays, mode=1):
        r = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        out = np.empty((r.shape[0], 3))
        if self.world_color_at(world.h, mode, _capi.dptr(r), r.shape[0], _capi.dptr(out)) != 0:
            raise RuntimeError("oracle: " + self.error())
        return out

    def quantise_rgba8(self, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.float64).reshape(-1, 3)
        out = np.empty((rgb.shape[0], 4), dtype=np.uint8)
        self.quantise(_capi.dptr(rgb), rgb.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def ppm(self, rgb, width, height):
        rgb = np.ascontiguousarray(rgb, dtype=np.float64).reshape(-1, 3)
        n = C.c_uint64(0)
        p = self.to_ppm(_capi.dptr(rgb), width, height, C.byref(n))
        try:
            return C.string_at(p, n.value)
        finally:
            self.free(p)


def load_oracle():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle.hpp", "oracle_capi.cpp", "Makefile")]
    if _newer(ORACLE_LIB, srcs):
        subprocess.run(["make", "-C", ORACLE_DIR, "_build/liboracle.so"], check=True, capture_output=True)
    return OracleApi(ORACLE_LIB)


class HostSim:
    """The product's device program compiled for the host (see tests/hostsim/hostsim.cpp)."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        vp = C.c_void_p
        self.lib.sim_last_error.restype = C.c_char_p
        self.lib.sim_scene_create.argtypes = [vp, C.POINTER(vp)]
        self.lib.sim_scene_create_ex.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int)]
        self.lib.sim_scene_tables.argtypes = [vp, _capi.c_u64_p]
        self.lib.sim_scene_gates.argtypes = [vp, _capi.c_double_p, C.c_uint64]
        self.lib.sim_scene_gates.restype = C.c_uint64
        self.lib.sim_scene_destroy.argtypes = [vp]
        self.lib.sim_scene_cluster_entries.argtypes = [vp, C.POINTER(C.c_int32), C.c_uint64]
        self.lib.sim_scene_cluster_entries.restype = C.c_uint64
        self.lib.sim_render.argtypes = [vp, C.POINTER(_capi.CameraDesc), C.POINTER(C.c_uint32), C.c_uint64, C.c_int,
                                        _capi.c_double_p, C.POINTER(C.c_uint8), _capi.c_u64_p]
        self.lib.sim_color_at.argtypes = [vp, _capi.c_double_p, C.c_uint64, _capi.c_double_p]

    def scene(self, world, device_build=False, clusters=True):
        """Marshal a product World exactly as rtc_world_scene does and flatten it for the simulation.  device_build:
        meshes go through the simulated device build (csrc/lbvh.cuh run as loops) instead of the host SAH builder."""
        api = world.api
        m = C.c_void_p()
        api.check(api.world_marshal(world.h, C.byref(m)))
        try:
            s = C.c_void_p()
            depth = C.c_int(0)
            rc = self.lib.sim_scene_create_ex(api.marshalled_desc(m), int(device_build) | (0 if clusters else 2),
                                              C.byref(s), C.byref(depth))
            if rc != 0:
                raise RuntimeError(f"hostsim {rc}: " + self.lib.sim_last_error().decode())
        finally:
            api.marshalled_free(m)
        sc = SimScene(self, s)
        sc.bvh_depth = depth.value
        return sc


class SimScene:
    def __init__(self, sim, handle):
        self.sim, self.h = sim, handle

    def gates(self):
        """(n, 6) gate boxes: lo xyz, hi xyz."""
        buf = np.zeros((64, 6))
        n = self.sim.lib.sim_scene_gates(self.h, buf.ctypes.data_as(_capi.c_double_p), 64)
        return buf[:n]

    def cluster_entries(self):
        """[(skip, prim), ...] of every LIST cluster's skip list (device_scene.h DBox32)."""
        buf = np.zeros((4096, 2), dtype=np.int32)
        n = self.sim.lib.sim_scene_cluster_entries(self.h, buf.ctypes.data_as(C.POINTER(C.c_int32)), 4096)
        return [tuple(int(x) for x in row) for row in buf[:n]]

    def tables(self):
        """(bvh nodes, triangles, meshes, content hash of the mesh tables)."""
        n = (C.c_uint64 * 4)()
        self.sim.lib.sim_scene_tables(self.h, n)
        return tuple(n)

    def render(self, cam, pixels=None, nthreads=None):
        d = cam.desc()
        nthreads = nthreads or (os.cpu_count() or 1)
        if pixels is None:
            n, pp = cam.hsize * cam.vsize, None
        else:
            pixels = np.ascontiguousarray(pixels, dtype=np.uint32)
            n, pp = pixels.shape[0], pixels.ctypes.data_as(C.POINTER(C.c_uint32))
        rgb = np.empty((n, 3))
        rgba = np.empty((n, 4), dtype=np.uint8)
        cnt = (C.c_uint64 * 4)()
        self.sim.lib.sim_render(self.h, C.byref(d), pp, n, nthreads, _capi.dptr(rgb),
                                rgba.ctypes.data_as(C.POINTER(C.c_uint8)), cnt)
        return rgb, rgba, list(cnt)

    def color_at(self, rays):
        r = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        out = np.empty((r.shape[0], 3))
        self.sim.lib.sim_color_at(self.h, _capi.dptr(r), r.shape[0], _capi.dptr(out))
        return out

    def __del__(self):
        if self.h:
            self.sim.lib.sim_scene_destroy(self.h)
            self.h = None


import contextlib


@contextlib.contextmanager
def scattered_description(world, seed=5):
    """(desc, scattered): the marshalled rtc_scene_desc of a product World and a copy whose triangle payloads are stored
    in a random order (shape.triangle indices follow) — the same world, described differently."""
    api = world.api
    m = C.c_void_p()
    api.check(api.world_marshal(world.h, C.byref(m)))
    try:
        desc = C.cast(api.marshalled_desc(m), C.POINTER(_capi.SceneDesc)).contents
        nt, ns = desc.triangle_count, desc.shape_count
        perm = np.random.default_rng(seed).permutation(nt)  # new position of each payload
        tris = (_capi.TriangleDesc * nt)()
        for old in range(nt):
            tris[int(perm[old])] = desc.triangles[old]
        shapes = (_capi.ShapeDesc * ns)()
        for i in range(ns):
            shapes[i] = desc.shapes[i]
            if shapes[i].kind == 6:
                shapes[i].triangle = int(perm[shapes[i].triangle])
        d2 = _capi.SceneDesc()
        C.memmove(C.byref(d2), C.byref(desc), C.sizeof(_capi.SceneDesc))
        d2.triangles, d2.shapes = tris, shapes
        yield desc, d2
    finally:
        api.marshalled_free(m)


def load_hostsim():
    csrc = os.path.join(ROOT, "ray-tracer-challenge-rust_b200", "csrc")
    srcs = [SIM_SRC, os.path.join(ROOT, "include", "rtc.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if _newer(SIM_LIB, srcs):
        os.makedirs(os.path.dirname(SIM_LIB), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-x", "c++",
                        SIM_SRC, "-o", SIM_LIB, "-lpthread"], check=True, capture_output=True)
    return HostSim(SIM_LIB)


def subset_pixels(hsize, vsize, step=16, offset=8):
    """The deterministic 1/step^2 pixel subset of a full-resolution camera."""
    xs = np.arange(offset, hsize, step, dtype=np.uint32)
    ys = np.arange(offset, vsize, step, dtype=np.uint32)
    gx, gy = np.meshgrid(xs, ys)
    return np.stack([gx.ravel(), gy.ravel()], axis=1).astype(np.uint32)


def compare_rgba(a, b):
    """-> (fraction of pixels exactly equal, max abs channel difference) over RGBA8 arrays."""
    a = np.asarray(a).reshape(-1, 4).astype(np.int32)
    b = np.asarray(b).reshape(-1, 4).astype(np.int32)
    exact = np.all(a == b, axis=1).mean() if len(a) else 1.0
    md = int(np.abs(a - b).max()) if len(a) else 0
    return float(exact), md
