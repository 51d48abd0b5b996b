"""Algorithmic work model of one frame: per-unit f64 flop costs x exact unit counts.

One flop = one f64 add/sub/mul/div/sqrt/min/max/compare-select as the REFERENCE's arithmetic needs it (pow counted as 1);
the padded BVH box tests are our own structure and are reported separately, never added to the algorithmic total.  The
unit counts come from rtc_render_tally (the same per-ray program with counters compiled in) plus the pixel count.
"""
import ctypes as C

from ._lib import api

# TallyIndex order in csrc/rt_core.cuh
TALLY_NAMES = ["xform_ray", "gate", "sphere", "plane", "cube", "cylinder", "cone", "tri_det", "tri_u", "tri_v",
               "tri_full", "bvh_box", "shade", "normal_sphere", "normal_plane", "normal_cube", "normal_cylinder",
               "normal_cone", "pattern", "pow", "refract", "schlick", "container_walk"]

# flops per unit, with the reference lines they count
FLOPS = {
    "ray_gen": 36,        # camera.rs:50-62: 4 + 2 + 18 (3x4 mat*point) + 3 + 9 (normalize)
    "xform_ray": 33,      # ray.rs:19-24: 18 (point) + 15 (vector)
    "gate": 20,           # shape.rs:403-425: 3 x check_axis (5) + 4 min/max + 1 compare
    "sphere": 27,         # shape.rs:260-271
    "plane": 2,           # shape.rs:276-280
    "cube": 20,           # shape.rs:305-318
    "cylinder": 58,       # shape.rs:323-351 walls 31 + caps 27 (:560-585)
    "cone": 63,           # shape.rs:359-397 + caps
    "tri_det": 15,        # shape.rs:439-443: cross 9 + dot 5 + compare 1
    "tri_u": 27,          # ... + div 1 + sub 3 + dot 5 + mul 1 + 2 compares
    "tri_v": 44,          # ... + cross 9 + dot 5 + mul 1 + add 1 + 2 compares
    "tri_full": 51,       # ... + dot 5 + mul 1  (shape.rs:439-455)
    "shade": 130,         # intersection.rs:18-27,68-69 + material.rs:48-74 + world.rs:102-104 (no pattern, no normal)
    "normal_sphere": 51,  # shape.rs:466-519: 18 (world->object) + 15 (inverse-transpose) + 2 x 9 (normalised twice)
    "normal_plane": 51,   # same transforms; the local normal is constant
    "normal_cube": 56,    # + 3 abs, 2 max
    "normal_cylinder": 56,
    "normal_cone": 56,
    "pattern": 41,        # pattern.rs:99-100 (18 + 18) + floors
    "pow": 1,
    "refract": 20,        # world.rs:141-152
    "schlick": 20,        # intersection.rs:109-127
    "quantise": 12,       # canvas.rs:61-63 x 3 channels, per pixel
}
BVH_BOX_FLOPS = 20        # our padded child-box test: 6 fma-able mul-adds + 8 min/max, ...; NOT reference arithmetic


def frame_tally(world, cam, rows=None, device=0):
    """-> dict name -> count for one frame of `cam` over `world` (runs the tally kernel; untimed)."""
    a = api()
    n = a.lib.rtc_tally_count()
    assert n == len(TALLY_NAMES), (n, len(TALLY_NAMES))
    counts = (C.c_uint64 * n)()
    d = cam.desc()
    a.lib.rtc_render_tally.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    a.lib.rtc_render_tally.restype = C.c_int
    a.check(a.lib.rtc_render_tally(world.scene(device), C.byref(d), C.byref(rows) if rows is not None else None,
                                   counts))
    t = dict(zip(TALLY_NAMES, list(counts)))
    npx = cam.hsize * (cam.rows_count(rows) if rows is not None else cam.vsize)
    t["ray_gen"] = npx
    t["quantise"] = npx
    return t


def algorithmic_flops(tally):
    """(reference-arithmetic flops, BVH box-test flops) of a frame."""
    f = sum(FLOPS[k] * tally.get(k, 0) for k in FLOPS)
    return f, BVH_BOX_FLOPS * tally.get("bvh_box", 0)
