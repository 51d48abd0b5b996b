"""Seeded random worlds for parity fuzzing, written against scene_api so they replay into the product and the oracle."""
import importlib
import math

import numpy as np

_pkg = importlib.import_module("ray-tracer-challenge-rust_b200")
sa = importlib.import_module("ray-tracer-challenge-rust_b200.scene_api")


def _rand_transform(T, rng, scale=(0.3, 1.5), spread=3.0):
    m = T.translation(*(rng.uniform(-spread, spread, 3)))
    m = m * T.rotation_y(rng.uniform(-math.pi, math.pi)) * T.rotation_x(rng.uniform(-1.0, 1.0))
    if rng.random() < 0.3:
        m = m * T.rotation_z(rng.uniform(-1.0, 1.0))
    if rng.random() < 0.2:
        m = m * T.shearing(*(rng.uniform(-0.3, 0.3, 6)))
    return m * T.scaling(*(rng.uniform(scale[0], scale[1], 3)))


def _rand_material(T, rng, allow_glass=True):
    m = sa.Material()
    m.color = tuple(rng.uniform(0.1, 1.0, 3))
    m.ambient, m.diffuse, m.specular = rng.uniform(0.05, 0.3), rng.uniform(0.3, 0.9), rng.uniform(0.0, 0.9)
    m.shininess = float(rng.choice([10.0, 50.0, 200.0, 300.0]))
    r = rng.random()
    if r < 0.25:
        m.reflective = rng.uniform(0.1, 0.9)
    elif r < 0.5 and allow_glass:
        m.transparency, m.refractive_index = rng.uniform(0.3, 1.0), float(rng.choice([1.0, 1.33, 1.5, 2.4]))
        if rng.random() < 0.6:
            m.reflective = rng.uniform(0.1, 0.9)
    if rng.random() < 0.4:
        kind = rng.integers(0, 5)
        a, b = tuple(rng.uniform(0, 1, 3)), tuple(rng.uniform(0, 1, 3))
        p = [sa.Pattern.stripe, sa.Pattern.gradient, sa.Pattern.ring, sa.Pattern.checkers][kind % 4](a, b) \
            if kind < 4 else sa.Pattern.test_pattern()
        if rng.random() < 0.7:
            p.set_transform(T.rotation_y(rng.uniform(0, 3)) * T.scaling(*(rng.uniform(0.2, 1.0, 3))))
        m.pattern = p
    return m


def _rand_leaf(S, T, rng, force_capped=False):
    k = rng.integers(0, 6)
    if k == 0:
        s = S.sphere()
    elif k == 1:
        s = S.cube()
    elif k == 2:
        lo = rng.uniform(-1.5, 0.0)
        s = S.cylinder(lo, lo + rng.uniform(0.5, 2.0), bool(rng.integers(0, 2)) or force_capped)
    elif k == 3:
        lo = rng.uniform(-1.0, 0.0)
        s = S.cone(lo, lo + rng.uniform(0.5, 1.5), bool(rng.integers(0, 2)) or force_capped)
    elif k == 4:
        p = rng.uniform(-1.5, 1.5, (3, 3))
        s = S.triangle(p[0], p[1], p[2])
    else:
        s = S.sphere()
    return s, k


def random_world(api, seed, hsize=48, vsize=32, nobjects=10, groups=True):
    """-> (world, camera).  Mixed primitives, nested groups with pushed-down transforms, glass-in-glass, a floor plane,
    small triangle fans (so MESH runs, BVHs and tie-breaks on shared edges are exercised)."""
    rng = np.random.default_rng(seed)
    T, S = sa.Transformations(api), sa.Shapes(api)
    cam = sa.CameraHandle(api, hsize, vsize, rng.uniform(0.6, 1.2))
    frm = rng.uniform(-1, 1, 3) * np.array([6, 2, 6]) + np.array([0, 3.5, 0])
    if abs(frm[0]) + abs(frm[2]) < 3:
        frm[2] -= 6
    cam.set_transform(T.view_transform(frm, rng.uniform(-0.5, 0.5, 3), (0.0, 1.0, 0.0)))
    world = sa.WorldHandle(api, sa.Light(tuple(rng.uniform(-6, 6, 3) + np.array([0, 8, 0])), tuple(rng.uniform(0.6, 1, 3))))

    floor = S.plane()
    floor.set_transform(T.translation(0, -2.0, 0))
    floor.material = _rand_material(T, rng, allow_glass=False)
    world.push(floor)

    for i in range(nobjects):
        r = rng.random()
        if groups and r < 0.25:
            g = S.group()
            for _ in range(rng.integers(1, 4)):
                # an uncapped cylinder/cone inside a group makes the reference panic (bounds.rs:143)
                leaf, k = _rand_leaf(S, T, rng, force_capped=True)
                leaf.set_transform(_rand_transform(T, rng, spread=1.0))
                leaf.material = _rand_material(T, rng)
                g.push_shape(leaf)
            if rng.random() < 0.25:  # a plane inside a group: Bounds::new gives it the (sic) box (-1,-1,0)..(1,1,0)
                pl = S.plane()
                pl.set_transform(T.translation(0, rng.uniform(-0.5, 0.5), 0) * T.rotation_x(rng.uniform(-0.4, 0.4)))
                pl.material = _rand_material(T, rng)
                g.push_shape(pl)
            if rng.random() < 0.5:  # nested group
                inner = S.group()
                leaf = S.sphere()
                leaf.set_transform(T.scaling(0.4, 0.4, 0.4))
                leaf.material = _rand_material(T, rng)
                inner.push_shape(leaf)
                inner.set_transform(T.translation(*(rng.uniform(-1, 1, 3))))
                g.push_shape(inner)
            g.set_transform(_rand_transform(T, rng, scale=(0.5, 1.2)))
            world.push(g)
        elif r < 0.4:  # a small closed-ish mesh: an octahedron through the OBJ path (shared edges and vertices)
            v = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64)
            v = v + rng.uniform(-0.1, 0.1, v.shape)
            f = np.array([[1, 3, 5], [3, 2, 5], [2, 4, 5], [4, 1, 5], [3, 1, 6], [2, 3, 6], [4, 2, 6], [1, 4, 6]],
                         dtype=np.int32)
            m = S.mesh(v, f)
            m.set_transform(_rand_transform(T, rng))
            m.set_material(_rand_material(T, rng))
            world.push(m)
        elif r < 0.5:  # glass inside glass: the n1/n2 container walk
            outer = S.glass_sphere()
            tr = _rand_transform(T, rng, scale=(1.0, 1.4))
            outer.set_transform(tr)
            outer.material.reflective = rng.uniform(0.0, 0.5)
            outer.material.diffuse, outer.material.ambient = 0.1, 0.05
            world.push(outer)
            inner = S.glass_sphere()
            inner.set_transform(tr * T.scaling(0.5, 0.5, 0.5))
            inner.material.refractive_index = 1.0 + rng.uniform(0.0, 1.0)
            inner.material.reflective = rng.uniform(0.0, 0.5)
            world.push(inner)
        else:
            leaf, k = _rand_leaf(S, T, rng)
            leaf.set_transform(_rand_transform(T, rng))
            leaf.material = _rand_material(T, rng)
            world.push(leaf)
    return world, cam


def random_soup_world(api, seed, hsize=40, vsize=28):
    """-> (world, camera).  Triangle soups big enough for the device mesh build (>= 256 triangles each): a lumpy closed
    shell with shared vertices and edges, exact duplicates of some of its triangles (equal-t hits: the tie goes to the
    lower DFS leaf whatever order a BVH visits them in), slivers, one soup of glass (the n1/n2 container walk crosses
    every triangle up to the hit), reflective floor, a few primitives in front and behind."""
    rng = np.random.default_rng(1000 + seed)
    T, S = sa.Transformations(api), sa.Shapes(api)
    cam = sa.CameraHandle(api, hsize, vsize, rng.uniform(0.7, 1.0))
    cam.set_transform(T.view_transform((rng.uniform(-2, 2), rng.uniform(1, 3), -7.0), (0.0, 0.3, 0.0), (0.0, 1.0, 0.0)))
    world = sa.WorldHandle(api, sa.Light((rng.uniform(-5, 5), 7.0, -6.0), (1.0, 1.0, 0.95)))

    floor = S.plane()
    floor.set_transform(T.translation(0, -1.6, 0))
    fm = sa.Material()
    fm.reflective = 0.35
    fm.pattern = sa.Pattern.checkers((0.2, 0.2, 0.2), (0.9, 0.9, 0.9))
    floor.material = fm
    world.push(floor)

    def shell(nu, nv, radius, lump):
        us = np.linspace(0, 2 * math.pi, nu, endpoint=False)
        vs = np.linspace(0.15, math.pi - 0.15, nv)
        verts = []
        for v in vs:
            for u in us:
                r = radius * (1 + lump * math.sin(3 * u) * math.sin(2 * v))
                verts.append((r * math.sin(v) * math.cos(u), r * math.cos(v), r * math.sin(v) * math.sin(u)))
        faces = []
        for j in range(nv - 1):
            for i in range(nu):
                a, b = j * nu + i, j * nu + (i + 1) % nu
                c, d = a + nu, b + nu
                faces += [(a + 1, b + 1, c + 1), (b + 1, d + 1, c + 1)]
        return np.array(verts, dtype=np.float64), np.array(faces, dtype=np.int32)

    for k in range(2):
        v, f = shell(int(rng.integers(12, 18)), int(rng.integers(10, 14)), rng.uniform(0.8, 1.3), rng.uniform(0.0, 0.3))
        v = v + rng.normal(0, 0.01, v.shape)
        dup = f[rng.integers(0, len(f), 24)]                       # exact duplicates of existing triangles
        sliver = f[rng.integers(0, len(f), 8)].copy()
        sliver[:, 2] = sliver[:, 1]                                 # degenerate (zero-area) triangles
        f = np.concatenate([f, dup, sliver]).astype(np.int32)
        g = S.mesh(v, f)
        m = sa.Material()
        m.color = tuple(rng.uniform(0.2, 1.0, 3))
        if k == 0:
            m.transparency, m.refractive_index, m.reflective = 0.8, 1.5, 0.3
            m.diffuse, m.ambient = 0.2, 0.05
        else:
            m.reflective = rng.uniform(0.0, 0.5)
        g.set_material(m)
        g.set_transform(T.translation(-1.4 + 2.8 * k, rng.uniform(-0.2, 0.4), rng.uniform(-0.5, 0.5)) *
                        T.rotation_y(rng.uniform(0, 3)) * T.scaling(1.0, rng.uniform(0.7, 1.2), 1.0))
        world.push(g)

    for _ in range(3):
        leaf, _k = _rand_leaf(S, T, rng, force_capped=True)
        leaf.set_transform(_rand_transform(T, rng, scale=(0.3, 0.7), spread=2.5))
        leaf.material = _rand_material(T, rng)
        world.push(leaf)
    return world, cam


def duplicate_glass_world(api, variant=0, hsize=64, vsize=48):
    """-> (world, camera).  The camera sits INSIDE shapes that the reference's Shape equality (shape.rs:638-646: kind,
    transform and material, 1e-5 tolerance) cannot tell apart, so the n1/n2 container walk of prepare_computations
    (intersection.rs:29-62) toggles ONE container where an identity comparison would see two:
      variant 0  two bit-identical glass spheres around the camera, a smaller glass sphere and a wall ahead
      variant 1  the twins differ by 3e-6 in their transforms (equal within EPSILON, different bits), one of them sits in
                 a group (behind a gate), plus twin glass cubes
      variant 2  a glass mesh around the camera whose OBJ lists some faces twice (value-equal triangles in one run)
    """
    T, S = sa.Transformations(api), sa.Shapes(api)
    cam = sa.CameraHandle(api, hsize, vsize, 1.0)
    cam.set_transform(T.view_transform((0.1, 0.2, -0.6), (0.0, 0.1, 4.0), (0.0, 1.0, 0.0)))
    world = sa.WorldHandle(api, sa.Light((-3.0, 5.0, -4.0), (1.0, 1.0, 0.9)))

    wall = S.plane()
    wall.set_transform(T.translation(0, 0, 6.0) * T.rotation_x(math.pi / 2))
    wm = sa.Material()
    wm.pattern = sa.Pattern.checkers((0.1, 0.1, 0.1), (0.9, 0.9, 0.9))
    wall.material = wm
    world.push(wall)

    def glass(shape, index, reflective=0.3, transparency=0.9):
        shape.material.transparency, shape.material.refractive_index = transparency, index
        shape.material.reflective = reflective
        shape.material.diffuse, shape.material.ambient = 0.1, 0.05
        return shape

    if variant == 0:
        for _ in range(2):
            s = glass(S.sphere(), 1.5)
            s.set_transform(T.scaling(2.0, 2.0, 2.0))
            world.push(s)
        inner = glass(S.sphere(), 2.0)
        inner.set_transform(T.translation(0.2, 0.1, 1.0) * T.scaling(0.5, 0.5, 0.5))
        world.push(inner)
    elif variant == 1:
        a = glass(S.sphere(), 1.5)
        a.set_transform(T.scaling(2.0, 2.0, 2.0))
        world.push(a)
        g = S.group()
        b = glass(S.sphere(), 1.5)
        b.set_transform(T.translation(3e-6, -2e-6, 0.0) * T.scaling(2.0, 2.0, 2.0))
        g.push_shape(b)
        other = S.sphere()
        other.set_transform(T.translation(2.5, 0.0, 3.0))
        g.push_shape(other)
        world.push(g)
        for k in range(2):
            c = glass(S.cube(), 1.33, reflective=0.0)
            c.set_transform(T.translation(0.0, 0.0, 0.5 + 1e-6 * k) * T.rotation_y(0.3) * T.scaling(1.2, 1.2, 1.2))
            world.push(c)
        third = glass(S.cube(), 2.4)
        third.set_transform(T.translation(-0.3, 0.0, 1.4) * T.scaling(0.3, 0.3, 0.3))
        world.push(third)
    else:
        v = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64) * 1.8
        f = [[1, 3, 5], [3, 2, 5], [2, 4, 5], [4, 1, 5], [3, 1, 6], [2, 3, 6], [4, 2, 6], [1, 4, 6]]
        f = np.array(f + [f[4], f[5], f[6], f[7], f[4]], dtype=np.int32)  # back faces twice, one of them three times
        m = S.mesh(v, f)
        gm = sa.Material()
        gm.transparency, gm.refractive_index, gm.reflective, gm.diffuse, gm.ambient = 0.9, 1.5, 0.3, 0.1, 0.05
        m.set_material(gm)
        world.push(m)
        inner = glass(S.sphere(), 2.0)
        inner.set_transform(T.translation(0.0, 0.1, 0.9) * T.scaling(0.4, 0.4, 0.4))
        world.push(inner)
    return world, cam


def chained_equal_world(api):
    """Three glass spheres a == b == c but a != c (x translations 0, 8e-6, 1.6e-5): Shape equality is not an
    equivalence here, so no class structure reproduces the reference's container walk — the library must refuse it."""
    T, S = sa.Transformations(api), sa.Shapes(api)
    cam = sa.CameraHandle(api, 8, 8, 1.0)
    cam.set_transform(T.view_transform((0.0, 0.0, -5.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)))
    world = sa.WorldHandle(api, sa.Light((-3.0, 5.0, -4.0), (1.0, 1.0, 1.0)))
    for k in range(3):
        s = S.glass_sphere()
        s.set_transform(T.translation(8e-6 * k, 0.0, 0.0))
        world.push(s)
    return world, cam


def smooth_world(api, seed, hsize=48, vsize=32):
    """-> (world, camera).  Smooth triangles (the book's SmoothTriangle, absent from the reference): a lumpy shell whose
    every triangle carries vertex normals (one of glass: refraction bends along the interpolated normal, the n1/n2 walk
    crosses smooth triangles), a hand-built group that MIXES flat and smooth triangles in one run (one mesh on the
    device), a reflective floor that mirrors them, and a non-uniformly scaled transform (normal_to_world matters)."""
    rng = np.random.default_rng(3000 + seed)
    T, S = sa.Transformations(api), sa.Shapes(api)
    cam = sa.CameraHandle(api, hsize, vsize, rng.uniform(0.7, 1.0))
    cam.set_transform(T.view_transform((rng.uniform(-2, 2), rng.uniform(1, 3), -7.0), (0.0, 0.3, 0.0), (0.0, 1.0, 0.0)))
    world = sa.WorldHandle(api, sa.Light((rng.uniform(-5, 5), 7.0, -6.0), (1.0, 1.0, 0.95)))
    floor = S.plane()
    floor.set_transform(T.translation(0, -1.6, 0))
    fm = sa.Material()
    fm.reflective = 0.35
    fm.pattern = sa.Pattern.checkers((0.2, 0.2, 0.2), (0.9, 0.9, 0.9))
    floor.material = fm
    world.push(floor)

    def shell(nu, nv, radius, lump):
        us = np.linspace(0, 2 * math.pi, nu, endpoint=False)
        vs = np.linspace(0.15, math.pi - 0.15, nv)
        verts = []
        for v in vs:
            for u in us:
                r = radius * (1 + lump * math.sin(3 * u) * math.sin(2 * v))
                verts.append((r * math.sin(v) * math.cos(u), r * math.cos(v), r * math.sin(v) * math.sin(u)))
        faces = []
        for j in range(nv - 1):
            for i in range(nu):
                a, b = j * nu + i, j * nu + (i + 1) % nu
                c, d = a + nu, b + nu
                faces += [(a + 1, b + 1, c + 1), (b + 1, d + 1, c + 1)]
        return np.array(verts, dtype=np.float64), np.array(faces, dtype=np.int32)

    for k in range(2):
        v, f = shell(int(rng.integers(10, 16)), int(rng.integers(8, 12)), rng.uniform(0.8, 1.3), rng.uniform(0.0, 0.3))
        n = v / np.linalg.norm(v, axis=1, keepdims=True) + rng.normal(0, 0.05, v.shape)  # not unit length on purpose
        if k == 0:
            g = S.smooth_mesh(v, n, f)
        else:  # flat and smooth triangles interleaved in one run of siblings
            g = S.group()
            for j, (a, b, c) in enumerate(f):
                if j % 3 == 0:
                    g.push_shape(S.triangle(v[a - 1], v[b - 1], v[c - 1]))
                else:
                    g.push_shape(S.smooth_triangle(v[a - 1], v[b - 1], v[c - 1], n[a - 1], n[b - 1], n[c - 1]))
        m = sa.Material()
        m.color = tuple(rng.uniform(0.2, 1.0, 3))
        if k == 0 and seed % 2 == 0:
            m.transparency, m.refractive_index, m.reflective = 0.8, 1.5, 0.3
            m.diffuse, m.ambient = 0.2, 0.05
        else:
            m.reflective = rng.uniform(0.0, 0.5)
        g.set_material(m)
        g.set_transform(T.translation(-1.4 + 2.8 * k, rng.uniform(-0.2, 0.4), rng.uniform(-0.5, 0.5)) *
                        T.rotation_y(rng.uniform(0, 3)) * T.scaling(1.0, rng.uniform(0.5, 1.4), 0.8))
        world.push(g)
    return world, cam
