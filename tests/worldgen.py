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
