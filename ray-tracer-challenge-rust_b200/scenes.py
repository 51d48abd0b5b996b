"""The five benchmark worlds of BASELINE.json, written against scene_api (so the same description can be replayed into
the product library and, by the tests, into the oracle).

C1 hexagon, C2 table and C3 teapot transcribe the reference's own builders (main.rs:84-146, 151-323, 368-397); C4 and
C5 are the synthetic scenes SURVEY.md §8(d) specifies on the reference's cow / teddy / pumpkin meshes.  Mesh data comes
from assets/*.npz (tools/import_assets.py made them from /root/reference/objs/*.obj, which does not exist on the GPU
box).
"""
import math
import os

import numpy as np

from .scene_api import (BLACK, BLUE, GREEN, WHITE, CameraHandle, Light, Material, Pattern, Shapes, Transformations,
                        WorldHandle)

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
PI = math.pi

# name -> (default hsize, default vsize) as BASELINE.json names them
CONFIGS = {
    "hexagon": (400, 200),
    "table": (1920, 1080),
    "teapot": (1920, 1080),
    "cow_teddy": (3840, 2160),
    "pumpkin": (7680, 4320),
    "cow_teddy_smooth": (3840, 2160),  # C4 with smooth triangles: not a BASELINE config of its own
}


def load_mesh(name):
    d = np.load(os.path.join(ASSETS, name + ".npz"))
    return d["vertices"], d["faces"]


def vertex_normals(vertices, faces):
    """One normal per vertex for an asset that ships none (cow-nonormals.obj, teddy.obj): the sum of the adjacent faces'
    cross products (area-weighted), accumulated in face order, normalised.  Plain numpy, so the oracle and the product
    are handed the same bits."""
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64) - 1
    fn = np.cross(v[f[:, 2]] - v[f[:, 0]], v[f[:, 1]] - v[f[:, 0]])  # e2 x e1, the orientation of shape.rs:185
    n = np.zeros_like(v)
    for k in range(3):
        np.add.at(n, f[:, k], fn)
    length = np.linalg.norm(n, axis=1, keepdims=True)
    length[length == 0.0] = 1.0
    return n / length


def _light():
    return Light((0.0, 6.9, -5.0), (1.0, 1.0, 0.9))


def _camera(api, T, hsize, vsize, frm, to):
    cam = CameraHandle(api, hsize, vsize, 0.785)
    cam.set_transform(T.view_transform(frm, to, (0.0, 1.0, 0.0)))
    return cam


def hexagon(api, hsize=400, vsize=200):  # main.rs:84-146
    T, S = Transformations(api), Shapes(api)
    cam = _camera(api, T, hsize, vsize, (8.0, 6.0, -8.0), (0.0, 0.0, 0.0))
    world = WorldHandle(api, _light())

    def corner():
        c = S.sphere()
        c.set_transform(T.translation(0., 0., -1.) * T.scaling(0.25, 0.25, 0.25))
        return c

    def edge():
        e = S.cylinder(0., 1., True)
        e.set_transform(T.translation(0., 0., -1.) * T.rotation_y(-PI / 6.) * T.rotation_z(-PI / 2.)
                        * T.scaling(0.25, 1., 0.25))
        return e

    def side():
        g = S.group()
        g.push_shape(corner())
        g.push_shape(edge())
        return g

    hexa = S.group()
    for i in range(6):
        s = side()
        s.set_transform(T.rotation_y(float(i) * PI / 3.))
        hexa.push_shape(s)
    hexa.set_transform(T.scaling(2.5, 2.5, 2.5))
    world.push(hexa)
    return world, cam


def table(api, hsize=1920, vsize=1080):  # main.rs:151-323
    T, S = Transformations(api), Shapes(api)
    cam = _camera(api, T, hsize, vsize, (8.0, 6.0, -8.0), (0.0, 3.0, 0.0))
    world = WorldHandle(api, _light())

    def cube(transform, **mat):
        c = S.cube()
        c.set_transform(transform)
        m = c.get_material_mut()
        for k, v in mat.items():
            setattr(m, k, v)
        world.push(c)

    p = Pattern.checkers(BLACK, (0.25, 0.25, 0.25))
    p.set_transform(T.scaling(0.07, 0.07, 0.07))
    cube(T.scaling(20.0, 7.0, 20.0) * T.translation(0.0, 1.0, 0.1), pattern=p, ambient=0.25, diffuse=0.7,
         specular=0.9, shininess=300.0, reflective=0.1)
    p = Pattern.checkers((0.4863, 0.3765, 0.2941), (0.3725, 0.2902, 0.2275))
    p.set_transform(T.scaling(0.05, 20.0, 0.05))
    cube(T.scaling(10.0, 10.0, 10.0), pattern=p, ambient=0.1, diffuse=0.7, specular=0.9, shininess=300.0,
         reflective=0.1)
    p = Pattern.stripe((0.5529, 0.4235, 0.3255), (0.6588, 0.5098, 0.4000))
    p.set_transform(T.scaling(0.05, 0.05, 0.05) * T.rotation_y(0.1))
    cube(T.translation(0.0, 3.1, 0.0) * T.scaling(3.0, 0.1, 2.0), pattern=p, ambient=0.1, diffuse=0.7, specular=0.9,
         shininess=300.0, reflective=0.2)
    for x, z in ((2.7, -1.7), (2.7, 1.7), (-2.7, -1.7), (-2.7, 1.7)):
        cube(T.translation(x, 1.5, z) * T.scaling(0.1, 1.5, 0.1), color=(0.5529, 0.4235, 0.3255), ambient=0.2,
             diffuse=0.7)
    cube(T.translation(0.0, 3.45001, 0.0) * T.rotation_y(0.2) * T.scaling(0.25, 0.25, 0.25), color=(1.0, 1.0, 0.8),
         ambient=0.0, diffuse=0.3, specular=0.9, shininess=300.0, reflective=0.1, transparency=0.7,
         refractive_index=1.5)
    cube(T.translation(1.0, 3.35, -0.9) * T.rotation_y(-0.4) * T.scaling(0.15, 0.15, 0.15), color=(1.0, 0.5, 0.5),
         reflective=0.6, diffuse=0.4)
    cube(T.translation(-1.5, 3.27, 0.3) * T.rotation_y(0.4) * T.scaling(0.15, 0.7, 0.15), color=(1.0, 1.0, 0.5))
    cube(T.translation(0.0, 3.25, 1.0) * T.rotation_y(0.4) * T.scaling(0.2, 0.05, 0.05), color=(0.5, 1.0, 0.5))
    cube(T.translation(-0.6, 3.4, -1.0) * T.rotation_y(0.8) * T.scaling(0.05, 0.2, 0.05), color=(0.5, 0.5, 1.0))
    cube(T.translation(2.0, 3.4, 1.0) * T.rotation_y(0.8) * T.scaling(0.05, 0.2, 0.05), color=(0.5, 1.0, 1.0))
    cube(T.translation(-10.0, 4.0, 1.0) * T.scaling(0.05, 1.0, 1.0), color=(0.7098, 0.2471, 0.2196), diffuse=0.6)
    cube(T.translation(-10.0, 3.4, 2.7) * T.scaling(0.05, 0.4, 0.4), color=(0.2667, 0.2706, 0.6902), diffuse=0.6)
    cube(T.translation(-10.0, 4.6, 2.7) * T.scaling(0.05, 0.4, 0.4), color=(0.3098, 0.5961, 0.3098), diffuse=0.6)
    cube(T.translation(-2.0, 3.5, 9.95) * T.scaling(5.0, 1.5, 0.05), color=(0.3882, 0.2627, 0.1882), diffuse=0.7)
    cube(T.translation(-2.0, 3.5, 9.95) * T.scaling(4.8, 1.4, 0.06), color=BLACK, diffuse=0.0, ambient=0.0,
         specular=0.0, shininess=300.0, reflective=1.0)
    return world, cam


def teapot(api, hsize=1920, vsize=1080):  # main.rs:368-397
    T, S = Transformations(api), Shapes(api)
    cam = _camera(api, T, hsize, vsize, (0.0, 4.0, -12.0), (0.0, 0.0, 0.0))
    world = WorldHandle(api, _light())
    pot = S.mesh(*load_mesh("teapot"))
    pot.set_transform(T.translation(0., -1.5, 0.))
    m = Material()
    m.pattern = Pattern.gradient(GREEN, BLUE)
    pot.set_material(m)
    world.push(pot)
    return world, cam


def _cow(api, T, S, cow=None):  # main.rs:340-351
    cow = cow if cow is not None else S.mesh(*load_mesh("cow"))
    cow.set_transform(T.translation(0., 3.5, 0.) * T.scaling(0.5, 0.5, 0.5))
    m = Material()
    m.color, m.ambient, m.diffuse, m.specular, m.shininess, m.reflective = WHITE, 0.1, 0.7, 0.9, 300.0, 0.2
    cow.set_material(m)
    return cow


def cow(api, hsize=400, vsize=200):  # main.rs:328-363 — what the shipped binary renders
    T, S = Transformations(api), Shapes(api)
    cam = _camera(api, T, hsize, vsize, (8.0, 6.0, -8.0), (0.0, 3.0, 0.0))
    world = WorldHandle(api, _light())
    world.push(_cow(api, T, S))
    return world, cam


def cow_teddy(api, hsize=3840, vsize=2160, smooth=False):
    """smooth=True: the same scene with every mesh triangle a smooth triangle over generated vertex normals (BASELINE's
    config 4 names smooth_triangle meshes; the assets carry no `vn` records and the reference has no smooth triangles)."""
    T, S = Transformations(api), Shapes(api)
    cam = _camera(api, T, hsize, vsize, (8.0, 6.0, -8.0), (0.0, 3.0, 0.0))
    world = WorldHandle(api, _light())

    def mesh(name):
        v, f = load_mesh(name)
        return S.smooth_mesh(v, vertex_normals(v, f), f) if smooth else S.mesh(v, f)

    world.push(_cow(api, T, S, mesh("cow")))
    teddy = mesh("teddy")
    teddy.set_transform(T.translation(-4.5, 2.6, 1.5) * T.scaling(0.12, 0.12, 0.12))
    m = Material()
    m.color = (1.0, 0.6, 0.3)
    teddy.set_material(m)
    world.push(teddy)
    floor = S.plane()
    fm = floor.get_material_mut()
    fm.color, fm.reflective = (0.8, 0.8, 0.8), 0.4
    world.push(floor)
    return world, cam


def pumpkin(api, hsize=7680, vsize=4320):
    T, S = Transformations(api), Shapes(api)
    cam = _camera(api, T, hsize, vsize, (0.0, 4.0, -12.0), (0.0, 1.75, 0.0))
    world = WorldHandle(api, _light())
    pk = S.mesh(*load_mesh("pumpkin"))
    pk.set_transform(T.translation(0., 7.3, 0.) * T.scaling(0.05, 0.05, 0.05) * T.rotation_x(-PI / 2.))
    m = Material()
    m.color, m.reflective, m.transparency, m.refractive_index = (1.0, 0.55, 0.1), 0.3, 0.5, 1.5
    pk.set_material(m)
    world.push(pk)
    floor = S.plane()
    fm = floor.get_material_mut()
    p = Pattern.checkers(BLACK, (0.25, 0.25, 0.25))
    p.set_transform(T.translation(0., 0.25, 0.) * T.scaling(0.5, 0.5, 0.5))
    fm.pattern, fm.reflective = p, 0.3
    world.push(floor)
    return world, cam


def cow_teddy_smooth(api, hsize=3840, vsize=2160):
    return cow_teddy(api, hsize, vsize, smooth=True)


BUILDERS = {"hexagon": hexagon, "table": table, "teapot": teapot, "cow": cow, "cow_teddy": cow_teddy,
            "pumpkin": pumpkin, "cow_teddy_smooth": cow_teddy_smooth}


def build(api, name, hsize=None, vsize=None):
    """-> (WorldHandle, CameraHandle) for a named config at its BASELINE resolution (or the one given)."""
    fn = BUILDERS[name]
    if hsize is None:
        hsize, vsize = CONFIGS.get(name, (400, 200))
    return fn(api, int(hsize), int(vsize))
