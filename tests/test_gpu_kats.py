"""The reference's own per-shape and per-intersection unit tests (src/shape.rs:692-1652, src/intersection.rs:203-379,
src/world.rs:200-209), asked of the CUDA path through the C-ABI probes rtc_intersect / rtc_normal_at /
rtc_prepare_computations (-m gpu).  The expected values are the ones the reference's tests hold (5 decimals, as its
assert_almost_eq! / approximate Tuple equality compare them)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ATOL = 1e-5  # utils.rs:2 — what the reference's assert_eq! on tuples and assert_almost_eq! tolerate
R2, R3 = math.sqrt(2.0), math.sqrt(3.0)


def _world(rtc, *shapes):
    w = rtc.World(rtc.Light((-10.0, 10.0, -10.0), (1.0, 1.0, 1.0)))
    for s in shapes:
        w.push(s)
    return w


def _norm(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


def _ts(world, origin, direction):
    return [t for t, _ in world.intersect([list(origin) + list(direction)])[0]]


@pytest.fixture()
def kit(rtc):
    return rtc, rtc.Shapes(rtc.api()), rtc.Transformations(rtc.api())


# ---------------------------------------------------------------------------------------------- spheres shape.rs:692-874
@pytest.mark.parametrize("origin,expected", [
    ((0, 0, -5), [4.0, 6.0]),    # shape.rs:692
    ((0, 1, -5), [5.0, 5.0]),    # :704 tangent
    ((0, 2, -5), []),            # :716 miss
    ((0, 0, 0), [-1.0, 1.0]),    # :726 inside
    ((0, 0, 5), [-6.0, -4.0]),   # :738 behind
])
def test_sphere_intersections(kit, origin, expected):
    rtc, S, T = kit
    got = _ts(_world(rtc, S.sphere()), origin, (0, 0, 1))
    assert len(got) == len(expected)
    np.testing.assert_allclose(got, expected, atol=ATOL)


def test_transformed_sphere_intersections(kit):
    rtc, S, T = kit
    s = S.sphere()
    s.set_transform(T.scaling(2, 2, 2))  # shape.rs:776
    np.testing.assert_allclose(_ts(_world(rtc, s), (0, 0, -5), (0, 0, 1)), [3.0, 7.0], atol=ATOL)
    s = S.sphere()
    s.set_transform(T.translation(5, 0, 0))  # shape.rs:790
    assert _ts(_world(rtc, s), (0, 0, -5), (0, 0, 1)) == []


def test_sphere_normals(kit):
    rtc, S, T = kit
    w = _world(rtc, S.sphere())
    pts = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (R3 / 3, R3 / 3, R3 / 3)]  # shape.rs:800-846
    n = w.normal_at(0, pts)
    np.testing.assert_allclose(n, pts, atol=ATOL)
    np.testing.assert_allclose(np.linalg.norm(n, axis=1), 1.0, atol=ATOL)  # :848 the normal is normalised
    s = S.sphere()
    s.set_transform(T.translation(0, 1, 0))  # shape.rs:861
    np.testing.assert_allclose(_world(rtc, s).normal_at(0, [(0, 1.70711, -0.70711)])[0], [0, 0.70711, -0.70711], atol=ATOL)
    s = S.sphere()
    s.set_transform(T.scaling(1, 0.5, 1) * T.rotation_z(math.pi / 5))  # shape.rs:869
    np.testing.assert_allclose(_world(rtc, s).normal_at(0, [(0, R2 / 2, -R2 / 2)])[0], [0, 0.97014, -0.24254], atol=ATOL)


def test_normal_on_a_child_object(kit):
    """shape.rs:955-974: transforms pushed down through two groups."""
    rtc, S, T = kit
    s = S.sphere()
    s.set_transform(T.translation(5, 0, 0))
    g2 = S.group()
    g2.push_shape(s)
    g2.set_transform(T.scaling(1, 2, 3))
    g1 = S.group()
    g1.push_shape(g2)
    g1.set_transform(T.rotation_y(math.pi / 2))
    n = _world(rtc, g1).normal_at(0, [(1.7321, 1.1547, -5.5774)])[0]
    np.testing.assert_allclose(n, [0.28570, 0.42854, -0.85716], atol=ATOL)


# ---------------------------------------------------------------------------------------------- planes shape.rs:980-1026
def test_plane(kit):
    rtc, S, T = kit
    w = _world(rtc, S.plane())
    np.testing.assert_allclose(w.normal_at(0, [(0, 0, 0), (10, 0, -10), (-5, 0, 150)]), [[0, 1, 0]] * 3, atol=ATOL)
    assert _ts(w, (0, 10, 0), (0, 0, 1)) == []  # parallel
    assert _ts(w, (0, 0, 0), (0, 0, 1)) == []   # coplanar
    assert _ts(w, (0, 1, 0), (0, -1, 0)) == [1.0]
    assert _ts(w, (0, -1, 0), (0, 1, 0)) == [1.0]


# ---------------------------------------------------------------------------------------------- cubes shape.rs:1033-1161
@pytest.mark.parametrize("origin,direction,t1,t2", [
    ((5, 0.5, 0), (-1, 0, 0), 4, 6), ((-5, 0.5, 0), (1, 0, 0), 4, 6), ((0.5, 5, 0), (0, -1, 0), 4, 6),
    ((0.5, -5, 0), (0, 1, 0), 4, 6), ((0.5, 0, 5), (0, 0, -1), 4, 6), ((0.5, 0, -5), (0, 0, 1), 4, 6),
    ((0, 0.5, 0), (0, 0, 1), -1, 1),
])
def test_ray_intersects_a_cube(kit, origin, direction, t1, t2):
    rtc, S, T = kit
    got = _ts(_world(rtc, S.cube()), origin, direction)
    assert len(got) == 2
    np.testing.assert_allclose(got, [t1, t2], atol=ATOL)


@pytest.mark.parametrize("origin,direction", [
    ((-2, 0, 0), (0.2673, 0.5345, 0.8018)), ((0, -2, 0), (0.8018, 0.2673, 0.5345)),
    ((0, 0, -2), (0.5345, 0.8018, 0.2673)), ((2, 0, 2), (0, 0, -1)), ((0, 2, 2), (0, -1, 0)), ((2, 2, 0), (-1, 0, 0)),
])
def test_ray_misses_a_cube(kit, origin, direction):
    rtc, S, T = kit
    assert _ts(_world(rtc, S.cube()), origin, direction) == []


def test_cube_normals(kit):
    rtc, S, T = kit
    cases = [((1, 0.5, -0.8), (1, 0, 0)), ((-1, -0.2, 0.9), (-1, 0, 0)), ((-0.4, 1, -0.1), (0, 1, 0)),
             ((0.3, -1, -0.7), (0, -1, 0)), ((-0.6, 0.3, 1), (0, 0, 1)), ((0.4, 0.4, -1), (0, 0, -1)),
             ((1, 1, 1), (1, 0, 0)), ((-1, -1, -1), (-1, 0, 0))]  # corners go to x: shape.rs:1153-1160
    n = _world(rtc, S.cube()).normal_at(0, [p for p, _ in cases])
    np.testing.assert_allclose(n, [e for _, e in cases], atol=ATOL)


# ------------------------------------------------------------------------------------------ cylinders shape.rs:1167-1380
def test_cylinder_intersections(kit):
    rtc, S, T = kit
    w = _world(rtc, S.cylinder())
    for origin, direction in (((1, 0, 0), (0, 1, 0)), ((0, 0, 0), (0, 1, 0)), ((0, 0, -5), (1, 1, 1))):
        assert _ts(w, origin, _norm(direction)) == []
    for origin, direction, t0, t1 in (((1, 0, -5), (0, 0, 1), 5, 5), ((0, 0, -5), (0, 0, 1), 4, 6),
                                      ((0.5, 0, -5), (0.1, 1, 1), 6.80798, 7.08872)):
        np.testing.assert_allclose(_ts(w, origin, _norm(direction)), [t0, t1], atol=ATOL)
    n = w.normal_at(0, [(1, 0, 0), (0, 5, -1), (0, -2, 1), (-1, 1, 0)])
    np.testing.assert_allclose(n, [(1, 0, 0), (0, 0, -1), (0, 0, 1), (-1, 0, 0)], atol=ATOL)


def test_constrained_and_capped_cylinders(kit):
    rtc, S, T = kit
    w = _world(rtc, S.cylinder(1.0, 2.0, False))  # shape.rs:1253-1312 (un-normalised directions, as the reference passes)
    for origin, direction, count in (((0, 1.5, 0), (0.1, 1, 0), 0), ((0, 3, -5), (0, 0, 1), 0), ((0, 0, -5), (0, 0, 1), 0),
                                     ((0, 2, -5), (0, 0, 1), 0), ((0, 1, -5), (0, 0, 1), 0), ((0, 1.5, -2), (0, 0, 1), 2)):
        assert len(_ts(w, origin, direction)) == count
    w = _world(rtc, S.cylinder(1.0, 2.0, True))  # shape.rs:1324-1362
    for origin, direction, count in (((0, 3, 0), (0, -1, 0), 2), ((0, 3, -2), (0, -1, 2), 2), ((0, 4, -2), (0, -1, 1), 2),
                                     ((0, 0, -2), (0, 1, 2), 2), ((0, -1, -2), (0, 1, 1), 2)):
        assert len(_ts(w, origin, _norm(direction))) == count
    pts = [(0, 1, 0), (0.5, 1, 0), (0, 1, 0.5), (0, 2, 0), (0.5, 2, 0), (0, 2, 0.5)]  # shape.rs:1364-1380
    np.testing.assert_allclose(w.normal_at(0, pts), [(0, -1, 0)] * 3 + [(0, 1, 0)] * 3, atol=ATOL)


# ---------------------------------------------------------------------------------------------- cones shape.rs:1387-1471
def test_cones(kit):
    rtc, S, T = kit
    w = _world(rtc, S.cone())
    for origin, direction, t0, t1 in (((0, 0, -5), (0, 0, 1), 5, 5), ((0, 0, -5), (1, 1, 1), 8.66025, 8.66025),
                                      ((1, 1, -5), (-0.5, -1, 1), 4.55006, 49.44994)):
        np.testing.assert_allclose(_ts(w, origin, _norm(direction)), [t0, t1], atol=ATOL)
    np.testing.assert_allclose(_ts(w, (0, 0, -1), _norm((0, 1, 1))), [0.35355], atol=ATOL)  # parallel to one half
    capped = _world(rtc, S.cone(-0.5, 0.5, True))
    for origin, direction, count in (((0, 0, -5), (0, 1, 0), 0), ((0, 0, -0.25), (0, 1, 1), 2), ((0, 0, -0.25), (0, 1, 0), 4)):
        assert len(_ts(capped, origin, _norm(direction))) == count
    n = w.normal_at(0, [(0, 0, 0), (1, 1, 1), (-1, -1, 0)])  # the reference normalises the book's local normals
    np.testing.assert_allclose(n, [(0, 0, 0), _norm((1, -R2, 1)), _norm((-1, 1, 0))], atol=ATOL)


# ------------------------------------------------------------------------------------ groups shape.rs:1478-1538, :913-974
def test_groups(kit):
    rtc, S, T = kit
    assert _ts(_world(rtc, S.group()), (0, 0, 0), (0, 0, 0)) == []  # empty group
    s1, s2, s3 = S.sphere(), S.sphere(), S.sphere()
    s2.set_transform(T.translation(0, 0, -3))
    s3.set_transform(T.translation(5, 0, 0))
    g = S.group()
    for s in (s1, s2, s3):
        g.push_shape(s)
    xs = _world(rtc, g).intersect([[0, 0, -5, 0, 0, 1]])[0]
    assert [leaf for _, leaf in xs] == [1, 1, 0, 0]  # s2, s2, s1, s1 (shape.rs:1519-1522)
    s = S.sphere()
    s.set_transform(T.translation(5, 0, 0))
    g = S.group()
    g.push_shape(s)
    g.set_transform(T.scaling(2, 2, 2))
    assert len(_ts(_world(rtc, g), (10, 0, -10), (0, 0, 1))) == 2  # shape.rs:1525-1538


# ------------------------------------------------------------------------------------------ triangles shape.rs:1545-1652
def test_triangles(kit):
    rtc, S, T = kit
    w = _world(rtc, S.triangle((0, 1, 0), (-1, 0, 0), (1, 0, 0)))
    assert _ts(w, (0, -1, -2), (0, 1, 0)) == []  # parallel
    assert _ts(w, (1, 1, -2), (0, 0, 1)) == []   # p1-p3 edge
    assert _ts(w, (-1, 1, -2), (0, 0, 1)) == []  # p1-p2 edge
    assert _ts(w, (0, -1, -2), (0, 0, 1)) == []  # p2-p3 edge
    assert _ts(w, (0, 0.5, -2), (0, 0, 1)) == [2.0]
    n = w.normal_at(0, [(0, 0.5, 0), (-0.5, 0.75, 0), (0.5, 0.25, 0)])
    np.testing.assert_allclose(n, [(0, 0, -1)] * 3, atol=ATOL)


# --------------------------------------------------------------------------------------------- world.rs:200-209
def test_default_world_intersections(kit):
    rtc, S, T = kit
    xs = rtc.World.default_world().intersect([[0, 0, -5, 0, 0, 1]])[0]
    np.testing.assert_allclose([t for t, _ in xs], [4, 4.5, 5.5, 6], atol=ATOL)
    assert [leaf for _, leaf in xs] == [0, 1, 1, 0]


# ------------------------------------------------------------------------------------- intersection.rs:203-264, 328-337
def test_prepare_computations(kit):
    rtc, S, T = kit
    w = _world(rtc, S.sphere())
    c = w.prepare_computations([[0, 0, -5, 0, 0, 1]])[0]  # intersection.rs:203-216, 234-241
    assert c.hit == 1 and c.leaf == 0 and c.inside == 0
    np.testing.assert_allclose([c.t], [4.0], atol=ATOL)
    np.testing.assert_allclose(list(c.point), [0, 0, -1], atol=ATOL)
    np.testing.assert_allclose(list(c.eyev), [0, 0, -1], atol=ATOL)
    np.testing.assert_allclose(list(c.normalv), [0, 0, -1], atol=ATOL)
    c = w.prepare_computations([[0, 0, 0, 0, 0, 1]])[0]   # intersection.rs:243-259: inside, normal inverted
    assert c.inside == 1
    np.testing.assert_allclose(list(c.point), [0, 0, 1], atol=ATOL)
    np.testing.assert_allclose(list(c.normalv), [0, 0, -1], atol=ATOL)
    c = _world(rtc, S.plane()).prepare_computations([[0, 1, -1, 0, -R2 / 2, R2 / 2]])[0]  # intersection.rs:218-232
    np.testing.assert_allclose(list(c.reflectv), [0, R2 / 2, R2 / 2], atol=ATOL)
    s = S.sphere()
    s.set_transform(T.translation(0, 0, 1))
    c = _world(rtc, s).prepare_computations([[0, 0, -5, 0, 0, 1]])[0]  # intersection.rs:261-264, 328-337
    assert c.over_point[2] < -ATOL / 2 and c.point[2] > c.over_point[2]
    assert c.under_point[2] > ATOL / 2 and c.point[2] < c.under_point[2]
    assert w.prepare_computations([[0, 2, -5, 0, 0, 1]])[0].hit == 0


def test_n1_and_n2_at_various_intersections(kit):
    """intersection.rs:288-325.  The reference asks prepare_computations about each of the six intersections of one ray;
    here the ray starts 0.01 before each of them, which makes it the hit and leaves the sorted list — hence n1 and n2 —
    what it was (intersections behind the origin still count, intersection.rs:32)."""
    rtc, S, T = kit
    a, b, c = S.glass_sphere(), S.glass_sphere(), S.glass_sphere()
    a.set_transform(T.scaling(2, 2, 2))
    a.get_material_mut().refractive_index = 1.5
    b.set_transform(T.translation(0, 0, -0.25))
    b.get_material_mut().refractive_index = 2.0
    c.set_transform(T.translation(0, 0, 0.25))
    c.get_material_mut().refractive_index = 2.5
    w = _world(rtc, a, b, c)
    xs = w.intersect([[0, 0, -4, 0, 0, 1]])[0]
    np.testing.assert_allclose([t for t, _ in xs], [2, 2.75, 3.25, 4.75, 5.25, 6], atol=ATOL)
    assert [leaf for _, leaf in xs] == [0, 1, 2, 1, 2, 0]
    expected = [(1.0, 1.5), (1.5, 2.0), (2.0, 2.5), (2.5, 2.5), (2.5, 1.5), (1.5, 1.0)]
    rays = [[0, 0, -4 + t - 0.01, 0, 0, 1] for t, _ in xs]
    for comps, (t, leaf), (n1, n2) in zip(w.prepare_computations(rays), xs, expected):
        assert comps.leaf == leaf and abs(comps.t - 0.01) < 1e-9
        assert (comps.n1, comps.n2) == (n1, n2)


def test_schlick(kit):
    rtc, S, T = kit
    w = _world(rtc, S.glass_sphere())
    tir, perpendicular, grazing = w.prepare_computations([[0, 0, R2 / 2, 0, 1, 0], [0, 0, 0, 0, 1, 0],
                                                          [0, 0.99, -2, 0, 0, 1]])
    assert tir.reflectance == 1.0                                   # intersection.rs:340-353
    assert abs(perpendicular.reflectance - 0.04) < ATOL             # :355-366
    # :368-379 builds its intersection by hand at the rounded t = 1.8589; the real hit is at 1.85893264, which moves the
    # point on the sphere and with it the reflectance by 8e-5
    assert abs(grazing.t - 1.8589) < 1e-4 and abs(grazing.reflectance - 0.48873) < 2e-4


def test_probes_agree_with_the_oracle_on_a_mesh(rtc, oracle):
    """rtc_intersect on a BVH mesh behind a gate: the sorted list of every ray of a small frame against brute force over
    the same triangles (numpy Moller-Trumbore in the mesh's object space is not bit-exact, so compare counts and 1e-9)."""
    import helpers
    world, cam = rtc.build_scene("teapot", 24, 12)
    ow, oc = helpers.scenes.build(oracle, "teapot", 24, 12)
    ref, _ = oracle.render(ow, oc, mode=oracle.CACHED)
    d = cam.desc()
    rays = []
    o = np.empty(4)   # orc_camera_ray_for_pixel writes 4-tuples (x, y, z, w)
    dr = np.empty(4)
    for y in range(12):
        for x in range(24):
            oracle.camera_ray_for_pixel(oc.h, x, y, o.ctypes.data_as(helpers._capi.c_double_p),
                                        dr.ctypes.data_as(helpers._capi.c_double_p))
            rays.append(list(o[:3]) + list(dr[:3]))
    xs = world.intersect(rays)
    hits = np.array([any(t >= 0 for t, _ in x) for x in xs])
    lit = (ref.reshape(-1, 3) != 0).any(axis=1)
    assert np.array_equal(hits, lit)  # a pixel is non-black exactly where World::intersect has a hit
    for x in xs:
        assert [t for t, _ in x] == sorted(t for t, _ in x)


# ---------------------------------------------------------------------------------------------- smooth triangles
# NOT reference tests: the reference quotes these scenarios commented out (intersection.rs:381-386, obj_file.rs:295-335);
# the expected values are the book's.  Parity of this kind with the reference is unpinned (it does not implement it).
def _smooth_tri(S):
    return S.smooth_triangle((0, 1, 0), (-1, 0, 0), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0))


def test_smooth_triangle_intersection_and_normal(kit):
    rtc, S, T = kit
    w = _world(rtc, _smooth_tri(S))
    ray = [-0.2, 0.3, -2, 0, 0, 1]
    np.testing.assert_allclose(_ts(w, ray[:3], ray[3:]), [2.0], atol=ATOL)
    # "A smooth triangle uses u/v to interpolate the normal": normal_at(tri, point(0, 0, 0), intersection_with_uv(1, tri, 0.45, 0.25))
    np.testing.assert_allclose(w.normal_at(0, [[0.45, 0.25, 0.0]])[0], [-0.5547, 0.83205, 0.0], atol=ATOL)
    # "An intersection with a smooth triangle stores u/v" (u = 0.45, v = 0.25 for this ray) and "Preparing the normal on a
    # smooth triangle": the normal prepare_computations reports is the interpolated one
    comps = w.prepare_computations([ray])[0]
    assert comps.hit == 1 and comps.leaf == 0 and abs(comps.t - 2.0) < ATOL
    np.testing.assert_allclose(comps.normalv, [-0.5547, 0.83205, 0.0], atol=ATOL)


def test_obj_faces_with_normals(kit):
    rtc, S, T = kit
    g = S.obj_str("v 0 1 0\nv -1 0 0\nv 1 0 0\n\nvn -1 0 0\nvn 1 0 0\nvn 0 1 0\n\nf 1//3 2//1 3//2\nf 1/0/3 2/102/1 3/14/2\n")
    assert g.ignored_lines == 0 and g.leaf_count() == 2
    w = _world(rtc, g)
    # t1.n1 = normals[3], t1.n2 = normals[1], t1.n3 = normals[2]; t2 = t1: both leaves interpolate the same normals
    for leaf in (0, 1):
        np.testing.assert_allclose(w.normal_at(leaf, [[0.45, 0.25, 0.0]])[0], [-0.5547, 0.83205, 0.0], atol=ATOL)
        np.testing.assert_allclose(w.normal_at(leaf, [[0.0, 0.0, 0.0]])[0], [0.0, 1.0, 0.0], atol=ATOL)   # n1 at p1
        np.testing.assert_allclose(w.normal_at(leaf, [[1.0, 0.0, 0.0]])[0], [-1.0, 0.0, 0.0], atol=ATOL)  # n2 at p2


def test_shared_divisor(rtc):
    """rt_core.cuh SharedDivisor: x / m, y / m, z / m (and check_axis's two quotients, the quadratics' two roots) with the
    reciprocal refinement of the compiler's own division sequence done once — every quotient must be bit for bit the
    compiler's a / d.  10^9 operand pairs per seed: raw bit patterns, ordinary magnitudes, special values."""
    import ctypes as C
    api = rtc.api()
    for seed in (1, 2, 3):
        bad = C.c_uint64(123)
        api.check(api.selftest_shared_divisor(0, 10 ** 9, seed, C.byref(bad)))
        assert bad.value == 0
