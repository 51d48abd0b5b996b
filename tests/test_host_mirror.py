"""The C ABI and the host mirror of the reference API (no GPU needed): exported symbols, transformations/matrix KATs
against the oracle bit for bit, construction semantics (push-down, panics), marshalling, flattening and gate boxes."""
import ctypes as C
import importlib
import math
import os
import re

import numpy as np
import pytest

import helpers

capi = importlib.import_module("ray-tracer-challenge-rust_b200._capi")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(rtc):
    header = open(os.path.join(ROOT, "include", "rtc.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(rtc_[a-z0-9_]+)\s*\(", header))
    assert len(names) > 50
    lib = C.CDLL(rtc.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(rtc, has_gpu):
    """Without a CUDA device every rendering entry point fails loudly with RTC_ERR_CUDA."""
    if has_gpu:
        pytest.skip("a GPU is present")
    w, c = rtc.build_scene("hexagon", 16, 8)
    with pytest.raises(rtc.RtcError) as e:
        c.render(w)
    assert e.value.code == rtc.RTC_ERR_CUDA
    with pytest.raises(rtc.RtcError) as e:
        w.color_at([[0, 0, -5, 0, 0, 1]])
    assert e.value.code == rtc.RTC_ERR_CUDA


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "ray-tracer-challenge-rust_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(d, f)).read()
                assert "liboracle" not in text and "oracle.hpp" not in text and "oracle_capi" not in text, f


def _mat(api, fn, *args):
    out = np.empty(16)
    getattr(api, fn)(*args, capi.dptr(out))
    return out


def test_transformations_match_oracle_bitwise(rtc, oracle):
    rng = np.random.default_rng(1)
    a, o = rtc.api(), oracle
    for _ in range(50):
        x, y, z = rng.uniform(-5, 5, 3)
        for fn, args in (("translation", (x, y, z)), ("scaling", (x, y, z)), ("rotation_x", (x,)), ("rotation_y", (y,)),
                         ("rotation_z", (z,)), ("shearing", tuple(rng.uniform(-1, 1, 6)))):
            assert np.array_equal(_mat(a, fn, *args).view(np.uint64), _mat(o, fn, *args).view(np.uint64)), fn
    Ta, To = rtc.Transformations(a), rtc.Transformations(o)
    for _ in range(50):
        frm, to = rng.uniform(-9, 9, 3), rng.uniform(-2, 2, 3)
        mo = To.view_transform(frm, to, (0, 1, 0)) * To.rotation_y(0.3)
        ma = Ta.view_transform(frm, to, (0, 1, 0)) * Ta.rotation_y(0.3)
        s = rng.uniform(0.2, 3, 3)
        ma, mo = ma * Ta.scaling(*s), mo * To.scaling(*s)
        assert np.array_equal(ma.v.view(np.uint64), mo.v.view(np.uint64))
        assert np.array_equal(ma.inverse().v.view(np.uint64), mo.inverse().v.view(np.uint64))
        assert np.array_equal(ma.transpose().v.view(np.uint64), mo.transpose().v.view(np.uint64))
        t = rng.uniform(-3, 3, 4)
        assert np.array_equal(ma.mul_tuple(t).view(np.uint64), mo.mul_tuple(t).view(np.uint64))


def test_matrix_known_answers(rtc):
    """matrix.rs:483-559, transformations.rs:278-319 (5-decimal KATs of the reference)."""
    T = rtc.Transformations(rtc.api())
    a = rtc.Matrix(rtc.api(), [-5, 2, 6, -8, 1, -5, 1, 8, 7, 7, -6, -7, 1, -3, 7, 4])
    b = a.inverse().array()
    np.testing.assert_allclose(b, [[0.21805, 0.45113, 0.24060, -0.04511], [-0.80827, -1.45677, -0.44361, 0.52068],
                                   [-0.07895, -0.22368, -0.05263, 0.19737], [-0.52256, -0.81391, -0.30075, 0.30639]],
                               atol=1e-5)
    with pytest.raises(ValueError):  # |det| < 1e-5 (matrix.rs:140)
        rtc.Matrix(rtc.api(), [-4, 2, -2, -3, 9, 6, 2, 6, 0, -5, 1, -5, 0, 0, 0, 0]).inverse()
    with pytest.raises(ValueError):  # scaling(0.01)^3 has det 1e-6: the reference refuses it
        T.scaling(0.01, 0.01, 0.01).inverse()
    v = T.view_transform((1, 3, 2), (4, -2, 8), (1, 1, 0)).array()
    np.testing.assert_allclose(v, [[-0.50709, 0.50709, 0.67612, -2.36643], [0.76772, 0.60609, 0.12122, -2.82843],
                                   [-0.35857, 0.59761, -0.71714, 0.0], [0, 0, 0, 1]], atol=1e-5)
    np.testing.assert_allclose(T.view_transform((0, 0, 0), (0, 0, 1), (0, 1, 0)).array(), T.scaling(-1, 1, -1).array())


def test_camera_known_answers(rtc):
    """camera.rs:84-142"""
    c = rtc.Camera(200, 125, math.pi / 2).desc()
    assert abs(c.pixel_size - 0.01) < 1e-5
    c = rtc.Camera(125, 200, math.pi / 2).desc()
    assert abs(c.pixel_size - 0.01) < 1e-5


def _desc(rtc, world):
    api = rtc.api()
    m = C.c_void_p()
    api.check(api.world_marshal(world.h, C.byref(m)))
    d = C.cast(api.marshalled_desc(m), C.POINTER(capi.SceneDesc)).contents
    return m, d


def test_set_transform_pushes_down_and_panics_like_the_reference(rtc):
    """shape.rs:196-218: groups keep the identity, leaves get T * own; a second call panics; singular matrices panic."""
    T, S = rtc.Transformations(rtc.api()), rtc.Shapes(rtc.api())
    g = S.group()
    s = S.sphere()
    s.set_transform(T.translation(5, 0, 0))
    g.push_shape(s)
    g.set_transform(T.scaling(2, 2, 2))
    with pytest.raises(ValueError, match="more than once"):
        g.set_transform(T.scaling(2, 2, 2))
    w = rtc.World(rtc.Light((0, 0, 0), (1, 1, 1)))
    w.push(g)
    m, d = _desc(rtc, w)
    try:
        assert d.shape_count == 2 and d.root_count == 1
        assert d.shapes[0].kind == rtc.api().GROUP and d.shapes[0].child_count == 1
        gt = np.array(d.transforms[d.shapes[0].transform].transform[:]).reshape(4, 4)
        assert np.array_equal(gt, np.eye(4))
        lt = np.array(d.transforms[d.shapes[1].transform].transform[:]).reshape(4, 4)
        expect = (T.scaling(2, 2, 2) * T.translation(5, 0, 0)).array()
        assert np.array_equal(lt, expect)
        li = np.array(d.transforms[d.shapes[1].transform].inverse[:])
        assert np.array_equal(li, (T.scaling(2, 2, 2) * T.translation(5, 0, 0)).inverse().v)
    finally:
        rtc.api().marshalled_free(m)
    s2 = S.sphere()
    with pytest.raises(ValueError, match="invertible"):
        s2.set_transform(T.scaling(0.01, 0.01, 0.01))
    with pytest.raises(ValueError, match="group"):
        S.sphere().push_shape(S.sphere())


def test_uncapped_cylinder_in_group_is_a_reference_panic(rtc):
    """bounds.rs:143: Bounds::add asserts is_point(); 0 * inf = NaN in w makes the reference panic on the first ray."""
    S = rtc.Shapes(rtc.api())
    g = S.group()
    g.push_shape(S.cylinder())
    w = rtc.World(rtc.Light((0, 0, 0), (1, 1, 1)))
    w.push(g)
    with pytest.raises(rtc.RtcError) as e:
        w.flatten_info()
    assert e.value.code == rtc.RTC_ERR_PANIC and "bounds.rs:143" in e.value.message


def test_flattening_of_the_configs(rtc):
    w, _ = rtc.build_scene("hexagon", 8, 4)
    info = w.flatten_info(want_gates=True)
    assert (info["leaves"], info["gates"], info["meshes"]) == (12, 7, 0)
    # the hexagon's outer gate contains the origin and spans the ring (bounds.rs:50-125)
    lo, hi = info["gate_boxes"][0][:3], info["gate_boxes"][0][3:]
    assert (lo <= 0).all() and (hi >= 0).all() and hi[0] > 2.5 and lo[0] < -2.5
    w, _ = rtc.build_scene("teapot", 8, 4)
    info = w.flatten_info(want_gates=True)
    # obj_to_group nests group{default_group{...}}: both groups have the same box, so the flattener
    # lets one gate stand for both
    assert info["leaves"] == 6320 and info["mesh_triangles"] == 6320 and info["meshes"] == 1 and info["gates"] == 1
    assert info["bvh_max_depth"] <= 40 and info["transforms"] == 1
    w, _ = rtc.build_scene("pumpkin", 8, 4)
    g = w.flatten_info(want_gates=True)["gate_boxes"][0]
    assert g[0] <= 0 <= g[3] and g[1] <= 0 <= g[4] and g[2] <= 0 <= g[5]  # origin-seeded even though the mesh is far away


def test_scene_create_rejects_malformed_descriptions(rtc):
    api = rtc.api()
    shapes = (capi.ShapeDesc * 1)()
    shapes[0].kind, shapes[0].material, shapes[0].transform = 0, 0, 0
    tr = (capi.TransformDesc * 1)()
    tr[0].transform[:] = list(np.eye(4).ravel())
    tr[0].inverse[:] = list(np.eye(4).ravel())
    mats = (capi.Material * 1)()
    api.material_default(C.byref(mats[0]))
    d = capi.SceneDesc()
    d.shapes, d.shape_count, d.root_count = shapes, 1, 1
    d.transforms, d.transform_count, d.materials, d.material_count = tr, 1, mats, 1
    out = C.c_void_p()

    def create():
        return api.scene_create(C.cast(C.byref(d), C.c_void_p), 0, C.byref(out))

    shapes[0].material = 3
    assert create() == rtc.RTC_ERR_INVALID and "material index" in api.error()
    shapes[0].material = 0
    shapes[0].kind = 42
    assert create() == rtc.RTC_ERR_INVALID
    shapes[0].kind = 0
    d.root_count = 2
    assert create() == rtc.RTC_ERR_INVALID
    d.root_count = 1
    tr[0].transform[12] = 0.5  # projective row: not representable with the reference's point/vector asserts
    assert create() == rtc.RTC_ERR_UNSUPPORTED
    tr[0].transform[12] = 0.0
    rc = create()  # valid now: succeeds on a GPU box, RTC_ERR_CUDA here
    assert rc in (rtc.RTC_OK, rtc.RTC_ERR_CUDA)
    if rc == rtc.RTC_OK:
        api.scene_destroy(out)


def _fold_group_box(vertices, faces, m):
    """bounds.rs:50-151 restated with elementwise numpy (IEEE, no FMA): every triangle's origin-seeded box, its eight
    corners through the triangle's transform as ((m0*x + m1*y) + m2*z) + m3*1, folded into an origin-seeded box."""
    tri = vertices[faces - 1]                                  # (n, 3 points, 3)
    lo = np.minimum(tri.min(axis=1), 0.0)
    hi = np.maximum(tri.max(axis=1), 0.0)
    out_lo, out_hi = np.zeros(3), np.zeros(3)
    for c in range(8):
        p = np.stack([(hi if (c >> (2 - a)) & 1 else lo)[:, a] for a in range(3)], axis=1)
        for r in range(3):
            v = ((m[r, 0] * p[:, 0] + m[r, 1] * p[:, 1]) + m[r, 2] * p[:, 2]) + m[r, 3] * 1.0
            out_lo[r] = min(out_lo[r], v.min())
            out_hi[r] = max(out_hi[r], v.max())
    return np.concatenate([out_lo, out_hi])


def test_big_mesh_gate_box_is_the_reference_fold_bit_for_bit(rtc):
    """A mesh group of 10 000 triangles takes the flattener's sliced, multi-threaded fold (flatten.hpp): its gate must
    still be exactly what Bounds::new folds — compared bit for bit with an independent numpy restatement."""
    from importlib import import_module
    scenes = import_module("ray-tracer-challenge-rust_b200.scenes")
    v, f = scenes.load_mesh("pumpkin")
    S, T = rtc.Shapes(rtc.api()), rtc.Transformations(rtc.api())
    g = S.mesh(v, f)
    m = T.translation(0.5, -1.25, 3.0) * T.rotation_y(0.4) * T.scaling(0.3, 0.31, 0.29)
    g.set_transform(m)
    w = rtc.World(rtc.Light((0, 0, 0), (1, 1, 1)))
    w.push(g)
    info = w.flatten_info(want_gates=True)
    assert info["gates"] == 1 and info["mesh_triangles"] == 10000
    # set_transform on a group multiplies into each child (shape.rs:203-217): outer -> inner group -> triangle
    eye = type(m)(rtc.api(), np.eye(4))
    mm = ((m * eye) * eye).v.reshape(4, 4)
    want = _fold_group_box(np.asarray(v, dtype=np.float64), np.asarray(f), mm)
    got = np.asarray(info["gate_boxes"][0], dtype=np.float64)
    assert np.array_equal(got, want), (got, want)


def test_long_leaf_run_with_an_unbounded_child_still_panics(rtc):
    """The sliced fold must not swallow the reference's panic (bounds.rs:143) for a group of thousands of leaves."""
    S = rtc.Shapes(rtc.api())
    g = S.group()
    for i in range(4500):
        g.push_shape(S.cylinder() if i == 3333 else S.sphere())
    w = rtc.World(rtc.Light((0, 0, 0), (1, 1, 1)))
    w.push(g)
    with pytest.raises(rtc.RtcError) as e:
        w.flatten_info()
    assert e.value.code == rtc.RTC_ERR_PANIC and "bounds.rs:143" in e.value.message


def test_build_flags_are_validated_before_any_device_work(rtc):
    api = rtc.api()
    w = rtc.World(rtc.Light((0, 0, 0), (1, 1, 1)))
    assert api.world_set_build(w.h, 7) == rtc.RTC_ERR_INVALID and "build flag" in api.error()
    assert api.world_set_build(w.h, rtc.RTC_BUILD_DEVICE_LBVH) == rtc.RTC_OK
    assert api.world_set_build(w.h, rtc.RTC_BUILD_HOST_SAH) == rtc.RTC_OK
    with pytest.raises(KeyError):
        w.set_build("fastest")
    m = C.c_void_p()
    api.check(api.world_marshal(w.h, C.byref(m)))
    try:
        out = C.c_void_p()
        assert api.scene_create_ex(api.marshalled_desc(m), 0, 0x10, C.byref(out)) == rtc.RTC_ERR_INVALID
    finally:
        api.marshalled_free(m)


def test_non_transitive_shape_equality_is_refused(rtc):
    """Three glass spheres a == b == c, a != c (shape.rs:638-646 with its 1e-5 tolerance): the reference's n1/n2 then
    depend on which of them sits in the container list; the library says so instead of guessing."""
    import worldgen
    w, _ = worldgen.chained_equal_world(rtc.api())
    world = rtc.World(_handle=w.h)
    w.h = None
    with pytest.raises(rtc.RtcError) as e:
        world.flatten_info()
    assert e.value.code == rtc.RTC_ERR_UNSUPPORTED and "not each other" in e.value.message


@pytest.mark.parametrize("name,mask", [("table", 388), ("hexagon", 73), ("teapot", 96), ("cow_teddy", 98), ("pumpkin", 226)])
def test_each_benchmark_world_has_a_kernel_of_its_own(rtc, name, mask):
    """The five BASELINE configs are rendered by instantiations compiled for exactly what they contain (render_launch.cuh):
    a cluster of cubes must not count as a triangle mesh, a mesh scene must not carry the cluster code, ..."""
    world, _ = rtc.build_scene(name, 64, 36)
    assert world.kernel_features() == (mask, mask)
