"""The product's per-ray program (csrc/rt_core.cuh) + flattener + BVH, compiled for the host by tests/hostsim, against
the oracle — bit for bit.  This is the CPU-box check of the logic the GPU runs (the GPU run of the same comparisons is
tests/test_gpu_parity.py); it is test infrastructure, not a product path."""
import importlib

import numpy as np
import pytest

import helpers
import worldgen


def _wrap(rtc, w, c):
    """scene_api handles built on the product api -> product World/Camera wrappers."""
    world = rtc.World(_handle=w.h)
    w.h = None
    cam = rtc.Camera.__new__(rtc.Camera)
    cam.api, cam.hsize, cam.vsize, cam.field_of_view, cam.h = c.api, c.hsize, c.vsize, c.field_of_view, c.h
    c.h = None
    return world, cam


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.mark.parametrize("name,w,h", [("hexagon", 200, 100), ("table", 160, 90), ("teapot", 64, 36), ("cow", 64, 32),
                                      ("cow_teddy", 64, 36), ("pumpkin", 64, 36)])
def test_configs_bit_exact(rtc, oracle, hostsim, name, w, h):
    world, cam = rtc.build_scene(name, w, h)
    ow, oc = helpers.scenes.build(oracle, name, w, h)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    rgb, rgba, scnt = hostsim.scene(world).render(cam)
    assert _bits_equal(ref, rgb)
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]
    assert np.array_equal(rgba, oracle.quantise_rgba8(ref))


@pytest.mark.parametrize("seed", range(12))
def test_random_worlds_bit_exact(rtc, oracle, hostsim, seed):
    world, cam = _wrap(rtc, *worldgen.random_world(rtc.api(), seed))
    ow, oc = worldgen.random_world(oracle, seed)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    rgb, _, scnt = hostsim.scene(world).render(cam)
    assert _bits_equal(ref, rgb), f"seed {seed}: {np.count_nonzero((ref != rgb).any(axis=1))} pixels differ"
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]


def test_faithful_and_cached_oracle_agree(oracle):
    """The timed CPU arm (faithful: the reference's algorithm and cost) and the golden generator (cached) are one
    arithmetic."""
    ow, oc = helpers.scenes.build(oracle, "hexagon", 80, 40)
    a, _ = oracle.render(ow, oc, mode=oracle.FAITHFUL, nthreads=1)
    b, _ = oracle.render(ow, oc, mode=oracle.CACHED)
    assert _bits_equal(a, b)
    ow, oc = helpers.scenes.build(oracle, "teapot", 1920, 1080)
    px = helpers.subset_pixels(1920, 1080, 240, 120)
    a, _ = oracle.render(ow, oc, mode=oracle.FAITHFUL, nthreads=1, pixels=px)
    b, _ = oracle.render(ow, oc, mode=oracle.CACHED, pixels=px)
    assert _bits_equal(a, b)


def test_full_resolution_subsets_bit_exact(rtc, oracle, hostsim):
    """BASELINE resolutions (the ulp-sensitive wall checkers of the table scene need the real 1920x1080 camera)."""
    for name, w, h, step in (("table", 1920, 1080, 24), ("teapot", 1920, 1080, 48), ("pumpkin", 7680, 4320, 160)):
        world, cam = rtc.build_scene(name, w, h)
        ow, oc = helpers.scenes.build(oracle, name, w, h)
        px = helpers.subset_pixels(w, h, step, step // 2)
        ref, _ = oracle.render(ow, oc, mode=oracle.CACHED, pixels=px)
        rgb, _, _ = hostsim.scene(world).render(cam, pixels=px)
        assert _bits_equal(ref, rgb), name


def test_explicit_rays_and_ties(rtc, oracle, hostsim):
    """color_at on explicit rays, including rays aimed exactly at shared mesh vertices/edges and cube edges."""
    world, cam = rtc.build_scene("teapot", 32, 16)
    ow, _ = helpers.scenes.build(oracle, "teapot", 32, 16)
    v, f = helpers.scenes.load_mesh("teapot")
    rng = np.random.default_rng(5)
    rays = []
    origin = np.array([0.0, 4.0, -12.0])
    for i in rng.integers(0, len(v), 200):
        target = v[i] + np.array([0.0, -1.5, 0.0])  # the teapot's translation
        d = target - origin
        rays.append(np.concatenate([origin, d / np.linalg.norm(d)]))
    for i in rng.integers(0, len(f), 200):  # edge midpoints
        a, b = v[f[i][0] - 1], v[f[i][1] - 1]
        d = (a + b) / 2 + np.array([0.0, -1.5, 0.0]) - origin
        rays.append(np.concatenate([origin, d / np.linalg.norm(d)]))
    rays = np.array(rays)
    assert _bits_equal(oracle.color_at(ow, rays), hostsim.scene(world).color_at(rays))


@pytest.mark.parametrize("name,w,h", [("teapot", 64, 36), ("cow", 64, 32), ("cow_teddy", 64, 36), ("pumpkin", 64, 36)])
def test_device_mesh_build_bit_exact(rtc, oracle, hostsim, name, w, h):
    """RTC_BUILD_DEVICE_LBVH (csrc/lbvh.cuh: Morton order, Karras hierarchy, fitted boxes), its kernel bodies run as loops
    by tests/hostsim: another tree, the same pixels and ray counts as the oracle — bit for bit."""
    world, cam = rtc.build_scene(name, w, h)
    ow, oc = helpers.scenes.build(oracle, name, w, h)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    scene = hostsim.scene(world, device_build=True)
    rgb, rgba, scnt = scene.render(cam)
    assert _bits_equal(ref, rgb)
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]
    assert 0 < scene.bvh_depth <= 46  # fits the device traversal stack (kBvhStackDepth - 2)
    host = hostsim.scene(world)
    assert scene.tables()[1] == host.tables()[1] and scene.tables()[3] != host.tables()[3]  # same triangles, other tree
    # the mesh groups' gate boxes were folded by the (simulated) device too: the values of the host fold (bounds.rs:50-151)
    assert scene.gates().shape == host.gates().shape and len(scene.gates()) >= 1
    assert np.array_equal(scene.gates(), host.gates())


def test_device_mesh_build_degenerate_meshes(rtc, oracle, hostsim):
    """Keys that collide: a mesh whose triangles all share one centroid (a fan of coincident copies) and a flat strip on
    one axis — the hierarchy must still split (equal keys are told apart by sorted position) and stay shallow."""
    n = 600
    verts = [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0)]
    faces = [(1, 2, 3)] * n                                   # n coincident triangles
    for i in range(n):                                          # a strip along x, all centroids on one line
        verts += [(2.0 + i, 0.0, 0.0), (3.0 + i, 0.0, 0.0), (2.5 + i, 0.0, 1.0)]
        faces.append((3 * i + 4, 3 * i + 5, 3 * i + 6))
    v, f = np.asarray(verts, dtype=np.float64), np.asarray(faces, dtype=np.int32)

    def build(api_like, mod):
        Sx, Tx = mod.Shapes(api_like), mod.Transformations(api_like)
        g = Sx.mesh(v, f)
        g.set_transform(Tx.scaling(0.01, 1.0, 1.0))
        w = mod.WorldHandle(api_like, mod.Light((0.0, 5.0, -5.0), (1.0, 1.0, 1.0)))
        w.push(g)
        cam = mod.CameraHandle(api_like, 48, 24, 0.9)
        cam.set_transform(Tx.view_transform((3.0, 2.0, -6.0), (3.0, 0.0, 0.0), (0.0, 1.0, 0.0)))
        return w, cam

    world, cam = _wrap(rtc, *build(rtc.api(), helpers.scenes))
    ow, oc = build(oracle, helpers.scenes)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    scene = hostsim.scene(world, device_build=True)
    rgb, _, scnt = scene.render(cam)
    assert _bits_equal(ref, rgb)
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]
    assert scene.bvh_depth <= 46


def test_device_mesh_build_gathers_scattered_triangles(rtc, hostsim):
    """A description whose triangle payloads are NOT stored in shape order (the boundary allows any indices) takes the
    gather path of the device build's input; the frame is the one of the in-order description."""
    import ctypes as C
    world, cam = rtc.build_scene("teapot", 48, 27)
    with helpers.scattered_description(world) as (desc, scattered):
        frames = []
        for d in (desc, scattered):
            s, depth = C.c_void_p(), C.c_int(0)
            assert hostsim.lib.sim_scene_create_ex(C.cast(C.byref(d), C.c_void_p), 1, C.byref(s), C.byref(depth)) == 0
            frames.append(helpers.SimScene(hostsim, s).render(cam)[0])
        assert _bits_equal(frames[0], frames[1])


def test_device_gate_fold_leaves_the_panic_to_the_host_build(rtc, hostsim):
    """A non-finite vertex makes Bounds::add panic in the reference (bounds.rs:143).  The device fold only flags it; the
    product then rebuilds on the host, which raises the reference's panic."""
    v, f = helpers.scenes.load_mesh("teapot")
    v = np.array(v, dtype=np.float64)
    v[17, 1] = np.inf
    S = rtc.Shapes(rtc.api())
    w = rtc.World(rtc.Light((0, 5, -5), (1, 1, 1)))
    w.push(S.mesh(v, f))
    scene = hostsim.scene(w, device_build=True)
    assert scene.bvh_depth >= 1 << 20          # the simulated device build says "rebuild on the host"
    with pytest.raises(RuntimeError, match="bounds.rs:143"):
        hostsim.scene(w)                        # ... and the host build panics like the reference


@pytest.mark.parametrize("seed", range(4))
def test_random_triangle_soups_bit_exact_under_both_builds(rtc, oracle, hostsim, seed):
    """Glass and mirror triangle soups with duplicate and degenerate triangles: the host SAH build and the simulated
    device build visit triangles in different orders and must both give the oracle's frame and ray counts."""
    world, cam = _wrap(rtc, *worldgen.random_soup_world(rtc.api(), seed))
    ow, oc = worldgen.random_soup_world(oracle, seed)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    for device_build in (False, True):
        scene = hostsim.scene(world, device_build=device_build)
        rgb, _, scnt = scene.render(cam)
        assert _bits_equal(ref, rgb), f"seed {seed} device_build={device_build}: " \
                                      f"{np.count_nonzero((ref != rgb).any(axis=1))} pixels differ"
        assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]
        assert cnt.refract > 0


@pytest.mark.parametrize("name,w,h", [("table", 160, 90), ("hexagon", 100, 50)])
def test_general_cube_precheck_bit_exact(rtc, oracle, hostsim, name, w, h):
    """With reject mode 3 switched off (RTC_B200_NO_DIAG_CUBE: every cube takes the nine-product EPSILON pre-check, as
    rotated ones always do) not a pixel or a ray count changes."""
    world, cam = rtc.build_scene(name, w, h)
    ow, oc = helpers.scenes.build(oracle, name, w, h)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    rgb, _, scnt = hostsim.scene(world, clusters=False).render(cam)
    assert _bits_equal(ref, rgb)
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]


@pytest.mark.parametrize("seed", range(12))
def test_random_worlds_bit_exact_without_clusters(rtc, oracle, hostsim, seed):
    world, cam = _wrap(rtc, *worldgen.random_world(rtc.api(), seed))
    ow, oc = worldgen.random_world(oracle, seed)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    rgb, _, scnt = hostsim.scene(world, clusters=False).render(cam)
    assert _bits_equal(ref, rgb), f"seed {seed}: {np.count_nonzero((ref != rgb).any(axis=1))} pixels differ"
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]


@pytest.mark.parametrize("ntri", [255, 256, 257, 511])
def test_device_mesh_build_threshold_sizes(rtc, oracle, hostsim, ntri):
    """Meshes just below and above the device build's threshold (256 triangles): below it the host builds the mesh even
    when the device build is requested; the frames are the oracle's either way."""
    rng = np.random.default_rng(ntri)
    nv = ntri + 2
    ang = np.linspace(0, 4 * np.pi, nv)
    v = np.stack([np.cos(ang) * (1 + 0.2 * np.arange(nv) / nv), np.linspace(-1, 1, nv) + rng.normal(0, 0.02, nv),
                  np.sin(ang) * (1 + 0.2 * np.arange(nv) / nv)], axis=1)
    f = np.array([(i + 1, i + 2, i + 3) for i in range(ntri)], dtype=np.int32)  # a ribbon spiralling upwards

    def build(api_like, mod):
        Sx, Tx = mod.Shapes(api_like), mod.Transformations(api_like)
        g = Sx.mesh(v, f)
        g.set_transform(Tx.rotation_x(0.3))
        w = mod.WorldHandle(api_like, mod.Light((2.0, 5.0, -5.0), (1.0, 1.0, 1.0)))
        w.push(g)
        cam = mod.CameraHandle(api_like, 40, 30, 0.9)
        cam.set_transform(Tx.view_transform((0.0, 1.0, -5.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)))
        return w, cam

    world, cam = _wrap(rtc, *build(rtc.api(), helpers.scenes))
    ow, oc = build(oracle, helpers.scenes)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    host = hostsim.scene(world)
    dev = hostsim.scene(world, device_build=True)
    for scene in (host, dev):
        rgb, _, scnt = scene.render(cam)
        assert _bits_equal(ref, rgb)
        assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]
    same_tables = host.tables()[3] == dev.tables()[3]
    assert same_tables == (ntri < 256)  # below the threshold the request is served by the host build


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_value_equal_shapes_are_one_container(rtc, oracle, hostsim, variant):
    """shape.rs:638-646 + intersection.rs:33,42: the container walk finds its containers by VALUE.  With the camera inside
    twin glass shapes an identity comparison gets other n1/n2 (and other pixels); the flattener groups value-equal leaves
    into classes and the walk toggles one container per class."""
    world, cam = _wrap(rtc, *worldgen.duplicate_glass_world(rtc.api(), variant))
    ow, oc = worldgen.duplicate_glass_world(oracle, variant)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    for device_build in (False, True):  # a class inside a device-built mesh sends the scene back to the host build
        rgb, _, scnt = hostsim.scene(world, device_build=device_build).render(cam)
        assert _bits_equal(ref, rgb), f"{np.count_nonzero((ref != rgb).any(axis=1))} pixels differ"
        assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]


def test_non_transitive_equality_is_refused(rtc, hostsim):
    world, cam = _wrap(rtc, *worldgen.chained_equal_world(rtc.api()))
    with pytest.raises(RuntimeError, match="not each other"):
        hostsim.scene(world)


def _drift_world(api):
    """Six cubes (three axis-aligned, three rotated) far apart (one cluster, radius ~ 150) and, for each, rays whose object-space x direction is
    below EPSILON: check_axis then calls the x slab 'parallel' and decides by the ORIGIN's x alone (shape.rs:593-599), so the
    reference reports hits for rays that have drifted out of the cube by the time they reach it."""
    T, S = worldgen.sa.Transformations(api), worldgen.sa.Shapes(api)
    world = worldgen.sa.WorldHandle(api, worldgen.sa.Light((0.0, 300.0, -300.0), (1.0, 1.0, 1.0)))
    mats = []
    for k in range(6):
        c = S.cube()
        m = T.translation(-150.0 + 60.0 * k, 3.0 * k, 10.0 * (k % 3))
        if k % 2:  # a rotated cube's world box has slack; an axis-aligned one's box IS the cube — the sharp case
            m = m * T.rotation_y(0.3 + 0.37 * k) * T.rotation_x(0.1 * k)
        else:
            m = m * T.scaling(1.0 + 0.5 * k, 1.0, 2.0)
        c.set_transform(m)
        c.material.color = (0.2 + 0.1 * k, 0.9 - 0.1 * k, 0.5)
        c.material.ambient = 0.4
        world.push(c)
        mats.append(m.array().reshape(4, 4))
    rays, drifted = [], []
    for m in mats:
        for dist in (150.0, 600.0, 1100.0):
            for ox in (0.9999, -0.9999, 0.995, -0.995):
                for lx in (2e-6, 5e-6, 9.5e-6, -2e-6, -9.5e-6, 1.5e-5):
                    o = m @ np.array([ox, 0.3, -dist, 1.0])
                    d = m @ np.array([lx, 0.0, 1.0, 0.0])
                    d = d / np.linalg.norm(d[:3])
                    rays.append(np.concatenate([o[:3], d[:3]]))
                    x_at_cube = ox + lx * (dist - 1.0)  # object-space x where the ray reaches the front face
                    drifted.append(abs(lx) < 1e-5 and abs(x_at_cube) > 1.0005)
    return world, np.array(rays), np.array(drifted)


def test_cluster_boxes_cover_the_epsilon_drift(rtc, oracle, hostsim):
    """The cluster BVH (flatten.hpp emit_cluster) pads a cube's box by EPSILON * (longest admitted ray) * scale because the
    reference reports 'parallel-axis' hits up to EPSILON * t outside the true cube; rays that start beyond the cluster's
    reach take the exact linear scan.  Both must give the reference's colours, drifted hits included."""
    w, rays, drifted = _drift_world(rtc.api())
    world = rtc.World(_handle=w.h)
    w.h = None
    ow, _, _ = _drift_world(oracle)
    ref = oracle.color_at(ow, rays)
    assert drifted.sum() > 40 and (ref[drifted] != 0).any(axis=1).sum() > 40  # the reference does report such hits
    scene = hostsim.scene(world)
    assert scene.tables()[2] == 1  # the six cubes did become one cluster
    assert _bits_equal(ref, scene.color_at(rays))
    assert _bits_equal(ref, hostsim.scene(world, clusters=False).color_at(rays))


@pytest.mark.parametrize("limit", [2, 3, 6, 8, 9, 12])
@pytest.mark.parametrize("name,w,h", [("table", 96, 54), ("pumpkin", 48, 27)])
def test_general_depth_bit_exact(rtc, oracle, hostsim, name, w, h, limit):
    """RECURSION_LIMIT (world.rs:11) as a parameter.  The iterative integrator with its explicit frame stack
    (rt_core.cuh color_at_general) against the reference's recursion with the constant edited: 2-3 = surfaces only, 6 =
    the reference's picture with the last generation's secondary rays counted, 8-9 = two bounces (binary branching at
    every glass hit), 12 = three.  Pixels and ray counts, bit for bit."""
    world, cam = rtc.build_scene(name, w, h)
    world.set_recursion_limit(limit)
    ow, oc = helpers.scenes.build(oracle, name, w, h)
    ow.set_recursion_limit(limit)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    rgb, _, scnt = hostsim.scene(world).render(cam)
    assert _bits_equal(ref, rgb), f"{np.count_nonzero((ref != rgb).any(axis=1))} pixels differ"
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]
    if limit >= 8:
        base, _ = oracle.render(helpers.scenes.build(oracle, name, w, h)[0], oc, mode=oracle.CACHED)
        assert not _bits_equal(base, ref)  # a second bounce is visible in these scenes


@pytest.mark.parametrize("seed", range(6))
def test_general_depth_random_worlds_bit_exact(rtc, oracle, hostsim, seed):
    world, cam = _wrap(rtc, *worldgen.random_world(rtc.api(), seed))
    ow, oc = worldgen.random_world(oracle, seed)
    limit = (8, 9, 11, 6, 3, 12)[seed]
    world.set_recursion_limit(limit)
    ow.set_recursion_limit(limit)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    rgb, _, scnt = hostsim.scene(world).render(cam)
    assert _bits_equal(ref, rgb), f"seed {seed}: {np.count_nonzero((ref != rgb).any(axis=1))} pixels differ"
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]


def test_recursion_limits_the_reference_cannot_run(rtc, hostsim):
    """limit % 3 == 1 reaches shade_hit with remaining = 0: `remaining - 1` underflows usize (world.rs:68)."""
    for limit in (1, 4, 7, 10):
        world, cam = rtc.build_scene("hexagon", 16, 8)
        world.set_recursion_limit(limit)
        with pytest.raises(rtc.RtcError) as e:
            world.flatten_info()
        assert e.value.code == rtc.RTC_ERR_PANIC and "world.rs:68" in e.value.message
    world.set_recursion_limit(20)
    with pytest.raises(rtc.RtcError) as e:
        world.flatten_info()
    assert e.value.code == rtc.RTC_ERR_UNSUPPORTED


def _many_leaves_world(api, seed, n=48):
    """n bounded leaves (cubes, spheres, capped cylinders; some glass, some mirrors) side by side at World level — more than
    kClusterListMax, so the cluster is a BVH (device_scene.h) — plus a floor plane that stays a PRIM entry."""
    rng = np.random.default_rng(500 + seed)
    T, S = worldgen.sa.Transformations(api), worldgen.sa.Shapes(api)
    cam = worldgen.sa.CameraHandle(api, 56, 40, 0.9)
    cam.set_transform(T.view_transform((0.0, 6.0, -14.0), (0.0, 0.5, 0.0), (0.0, 1.0, 0.0)))
    world = worldgen.sa.WorldHandle(api, worldgen.sa.Light((-6.0, 12.0, -10.0), (1.0, 1.0, 0.95)))
    floor = S.plane()
    floor.set_transform(T.translation(0, -1.0, 0))
    floor.material = worldgen._rand_material(T, rng, allow_glass=False)
    world.push(floor)
    for _ in range(n):
        k = rng.integers(0, 3)
        leaf = S.cube() if k == 0 else (S.sphere() if k == 1 else S.cylinder(-1.0, 1.0, True))
        leaf.set_transform(T.translation(rng.uniform(-8, 8), rng.uniform(-0.5, 3.0), rng.uniform(-6, 8)) *
                           T.rotation_y(rng.uniform(0, 3)) * T.scaling(*(rng.uniform(0.2, 0.9, 3))))
        leaf.material = worldgen._rand_material(T, rng)
        world.push(leaf)
    return world, cam


@pytest.mark.parametrize("seed", range(3))
def test_large_cluster_is_a_tree_bit_exact(rtc, oracle, hostsim, seed):
    world, cam = _wrap(rtc, *_many_leaves_world(rtc.api(), seed))
    ow, oc = _many_leaves_world(oracle, seed)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    scene = hostsim.scene(world)
    assert scene.tables()[0] >= 40 and scene.tables()[2] == 1  # one cluster, with BVH nodes
    rgb, _, scnt = scene.render(cam)
    assert _bits_equal(ref, rgb), f"{np.count_nonzero((ref != rgb).any(axis=1))} pixels differ"
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]


# ---- smooth triangles.  Parity UNPINNED against the reference: it does not implement them (its
# scenarios are commented out at intersection.rs:381-386 and obj_file.rs:295-335); the oracle follows the book's definition
# those comments quote, and what is checked here is that the product computes that definition bit for bit.
@pytest.mark.parametrize("seed", range(4))
def test_smooth_triangle_worlds_bit_exact(rtc, oracle, hostsim, seed):
    world, cam = _wrap(rtc, *worldgen.smooth_world(rtc.api(), seed))
    ow, oc = worldgen.smooth_world(oracle, seed)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    assert world.kernel_features()[0] & 2048
    rgb, _, scnt = hostsim.scene(world).render(cam)
    assert _bits_equal(ref, rgb), f"seed {seed}: {np.count_nonzero((ref != rgb).any(axis=1))} pixels differ"
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]


def test_smooth_cow_and_teddy_bit_exact_and_different_from_flat(rtc, oracle, hostsim):
    world, cam = rtc.build_scene("cow_teddy_smooth", 96, 54)
    assert world.kernel_features() == (98 + 2048, 98 + 2048)  # its own instantiation
    ow, oc = helpers.scenes.build(oracle, "cow_teddy_smooth", 96, 54)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    rgb, _, scnt = hostsim.scene(world).render(cam)
    assert _bits_equal(ref, rgb)
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]
    fw, fc = helpers.scenes.build(oracle, "cow_teddy", 96, 54)
    flat, _ = oracle.render(fw, fc, mode=oracle.CACHED)
    assert (flat != ref).any(axis=1).mean() > 0.05  # interpolated normals do change the shading


OBJ_WITH_NORMALS = """
v 0 1 0
v -1 0 0
v 1 0 0
v 0 -1 0.5

vn -1 0 -0.2
vn 1 0 -0.2
vn 0 1 -0.2
vn 0 -1 -1

f 1//3 2//1 3//2
f 2/7/1 4/8/4 3/9/2
f 1 2 4
g side
f 1//3 3//2 4//4 2//1
"""


def test_obj_with_vertex_normals_parses_like_the_oracle(rtc, oracle, hostsim):
    """obj_file.rs:295-335 (commented scenarios): `vn` records, `f v//n` and `f v/t/n` corners, a flat face among them,
    fan triangulation carrying the normals, a named group — parsed by the product's parser and by the oracle's, rendered
    by both."""
    def build(api):
        T, S = worldgen.sa.Transformations(api), worldgen.sa.Shapes(api)
        cam = worldgen.sa.CameraHandle(api, 40, 30, 0.8)
        cam.set_transform(T.view_transform((0.3, 0.4, -4.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)))
        world = worldgen.sa.WorldHandle(api, worldgen.sa.Light((-3.0, 4.0, -5.0), (1.0, 1.0, 1.0)))
        g = S.obj_str(OBJ_WITH_NORMALS)
        assert g.ignored_lines == 0 and g.leaf_count() == 5
        g.set_transform(T.rotation_y(0.3) * T.scaling(1.2, 1.0, 1.0))
        world.push(g)
        return world, cam
    world, cam = _wrap(rtc, *build(rtc.api()))
    ow, oc = build(oracle)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    assert (ref != 0).any()
    rgb, _, scnt = hostsim.scene(world).render(cam)
    assert _bits_equal(ref, rgb)
    assert scnt == [cnt.primary, cnt.shadow, cnt.reflect, cnt.refract]
    with pytest.raises(ValueError):  # a normal index past the `vn` records panics like a vertex index does
        worldgen.sa.Shapes(rtc.api()).obj_str("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//2 3//1\n")
    with pytest.raises(ValueError):
        worldgen.sa.Shapes(oracle).obj_str("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//2 3//1\n")


def test_cluster_skip_list_is_well_formed(rtc, hostsim):
    """The table scene's 18 cubes are one LIST cluster: every leaf is in its skip list exactly once, every header covers
    entries inside the list, headers nest, and there ARE headers (the table and what stands on it, the pictures on the
    walls) — a ray that misses a header's box skips its leaves."""
    world, cam = rtc.build_scene("table", 64, 36)
    e = hostsim.scene(world).cluster_entries()
    leaves = sorted(p for s, p in e if s < 0)
    assert leaves == list(range(18))
    headers = [(i, s) for i, (s, p) in enumerate(e) if s >= 0]
    assert len(headers) >= 2 and len(e) == 18 + len(headers)
    for i, s in headers:
        assert e[i][1] == -1 and s >= 3 and i + s < len(e) + 0  # covers at least three entries, all inside the list
        for j, t in headers:
            if i < j <= i + s:
                assert j + t <= i + s  # nested, never straddling
