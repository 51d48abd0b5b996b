"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same worlds (-m gpu)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

# BASELINE.json's correctness bar for PPMs: every channel within +-1/255, >= 99.9 % of pixels exact.  The f64 frame is
# additionally required to be bit-identical except where the specular pow() (CUDA <= 2 ulp vs glibc < 1 ulp,
# material.rs:68) contributes, where a relative 1e-12 is allowed.
PPM_MAX_DIFF = 1
PPM_MIN_EXACT = 0.999
F64_RTOL = 1e-12

SMALL = [("hexagon", 400, 200), ("table", 320, 180), ("teapot", 160, 90), ("cow", 160, 80), ("cow_teddy", 160, 90),
         ("pumpkin", 160, 90)]
# (config, hsize, vsize, oracle pixel step): 1 = every pixel of the frame, 8 = the 1/64 subset
FULL = [("hexagon", 1920, 960, 1), ("table", 1920, 1080, 1), ("teapot", 1920, 1080, 8), ("cow_teddy", 3840, 2160, 8),
        ("pumpkin", 7680, 4320, 8)]


def _check(ref_rgb, got_rgb, ref_rgba, got_rgba):
    exact, md = helpers.compare_rgba(ref_rgba, got_rgba)
    assert md <= PPM_MAX_DIFF, f"channel differs by {md}/255"
    assert exact >= PPM_MIN_EXACT, f"only {exact:.5f} of pixels exact"
    if got_rgb is not None:
        np.testing.assert_allclose(got_rgb, ref_rgb, rtol=F64_RTOL, atol=1e-15)
        bit_equal = (ref_rgb.view(np.uint64) == got_rgb.view(np.uint64)).all(axis=1).mean()
        assert bit_equal >= 0.95, f"only {bit_equal:.4f} of f64 pixels bit-identical"
    return exact, md


@pytest.mark.parametrize("name,w,h", SMALL)
def test_full_frame_matches_oracle(rtc, oracle, name, w, h):
    world, cam = rtc.build_scene(name, w, h)
    st = rtc.Stats()
    canvas = cam.render(world, want_f64=True, stats=st)
    ow, oc = helpers.scenes.build(oracle, name, w, h)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    got = canvas.pixels_f64().reshape(-1, 3)
    _check(ref, got, oracle.quantise_rgba8(ref), canvas.pixels_rgba8())
    # ray accounting is exact
    assert (st.primary_rays, st.shadow_rays, st.reflect_rays, st.refract_rays) == \
        (cnt.primary, cnt.shadow, cnt.reflect, cnt.refract)
    # and the PPM text is the reference encoder's, byte for byte, when the quantised frame is identical
    if np.array_equal(oracle.quantise_rgba8(ref).reshape(-1), canvas.pixels_rgba8().reshape(-1)):
        assert canvas.to_ppm() == oracle.ppm(ref, w, h)


@pytest.mark.parametrize("name,w,h,step", FULL)
def test_full_resolution_matches_oracle(rtc, oracle, name, w, h, step):
    """BASELINE resolutions.  The mesh-free configs are compared over the WHOLE frame, f64 colours included (the cached
    oracle renders a 1080p table frame in about a second on the box's cores); the mesh configs, where the oracle scans
    every triangle per ray as the reference does, over the deterministic 1/64 pixel subset of the same camera."""
    world, cam = rtc.build_scene(name, w, h)
    rgba = np.empty((h, w, 4), dtype=np.uint8)
    rgb = np.empty((h, w, 3)) if step == 1 else None
    st = rtc.Stats()
    cam.render_into(world, rgba8=rgba, rgb_f64=rgb, stats=st)
    assert st.primary_rays == w * h
    ow, oc = helpers.scenes.build(oracle, name, w, h)
    if step == 1:
        ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
        _check(ref, rgb.reshape(-1, 3), oracle.quantise_rgba8(ref), rgba.reshape(-1, 4))
        assert (st.primary_rays, st.shadow_rays, st.reflect_rays, st.refract_rays) == \
            (cnt.primary, cnt.shadow, cnt.reflect, cnt.refract)
    else:
        px = helpers.subset_pixels(w, h, step, step // 2)
        ref, _ = oracle.render(ow, oc, mode=oracle.CACHED, pixels=px)
        _check(ref, None, oracle.quantise_rgba8(ref), rgba[px[:, 1], px[:, 0]])


@pytest.mark.parametrize("seed", range(8))
def test_random_worlds_match_oracle(rtc, oracle, seed):
    """Seeded random worlds (nested groups, glass in glass, cones, shears, patterns, small meshes) on the GPU."""
    import worldgen
    w, c = worldgen.random_world(rtc.api(), seed, hsize=96, vsize=64)
    world = rtc.World(_handle=w.h)
    w.h = None
    cam = rtc.Camera.__new__(rtc.Camera)
    cam.api, cam.hsize, cam.vsize, cam.field_of_view, cam.h = c.api, c.hsize, c.vsize, c.field_of_view, c.h
    c.h = None
    st = rtc.Stats()
    canvas = cam.render(world, stats=st)
    ow, oc = worldgen.random_world(oracle, seed, hsize=96, vsize=64)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    _check(ref, canvas.pixels_f64().reshape(-1, 3), oracle.quantise_rgba8(ref), canvas.pixels_rgba8())
    assert (st.primary_rays, st.shadow_rays, st.reflect_rays, st.refract_rays) == \
        (cnt.primary, cnt.shadow, cnt.reflect, cnt.refract)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("build", ["host", "device"])
def test_value_equal_shapes_are_one_container(rtc, oracle, variant, build):
    """shape.rs:638-646 / intersection.rs:33,42 on the GPU: twin glass shapes around the camera are ONE container of the
    n1/n2 walk (classes of value-equal leaves, flatten.hpp:build_classes); an identity comparison renders other pixels."""
    import worldgen
    w, c = worldgen.duplicate_glass_world(rtc.api(), variant)
    world = rtc.World(_handle=w.h)
    w.h = None
    cam = rtc.Camera.__new__(rtc.Camera)
    cam.api, cam.hsize, cam.vsize, cam.field_of_view, cam.h = c.api, c.hsize, c.vsize, c.field_of_view, c.h
    c.h = None
    world.set_build(build)
    st = rtc.Stats()
    canvas = cam.render(world, stats=st)
    ow, oc = worldgen.duplicate_glass_world(oracle, variant)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    _check(ref, canvas.pixels_f64().reshape(-1, 3), oracle.quantise_rgba8(ref), canvas.pixels_rgba8())
    assert (st.primary_rays, st.shadow_rays, st.reflect_rays, st.refract_rays) == \
        (cnt.primary, cnt.shadow, cnt.reflect, cnt.refract)


def test_default_world_known_answers(rtc):
    """camera.rs:145-155 and world.rs:212-260 through the CUDA path."""
    import math
    T = rtc.Transformations(rtc.api())
    w = rtc.World.default_world()
    c = rtc.Camera(11, 11, math.pi / 2.)
    c.set_transform(T.view_transform((0., 0., -5.), (0., 0., 0.), (0., 1., 0.)))
    image = c.render(w)
    np.testing.assert_allclose(image.get_pixel(5, 5), [0.38066, 0.47583, 0.2855], atol=1e-5)
    cols = w.color_at([[0, 0, -5, 0, 1, 0], [0, 0, -5, 0, 0, 1]])
    np.testing.assert_allclose(cols[0], [0, 0, 0], atol=1e-5)
    np.testing.assert_allclose(cols[1], [0.38066, 0.47583, 0.2855], atol=1e-5)


def test_reference_world_known_answers(rtc):
    """world.rs:335-352, 488-546 (reflective / transparent / Schlick shade_hit) through World::color_at on the GPU.
    The reference tests call shade_hit(comps, 5) directly; color_at reaches shade_hit with 4 — both budgets give exactly
    two shaded generations, so the 5-decimal answers are the same."""
    import math
    T, S = rtc.Transformations(rtc.api()), rtc.Shapes(rtc.api())
    r2 = math.sqrt(2.0) / 2.0
    ray = [[0, 0, -3, 0, -r2, r2]]

    w = rtc.World.default_world()
    plane = S.plane()
    plane.material.reflective = 0.5
    plane.set_transform(T.translation(0, -1, 0))
    w.push(plane)
    np.testing.assert_allclose(w.color_at(ray)[0], [0.87675, 0.92434, 0.82918], atol=1e-5)

    for reflective, want in ((0.0, [0.93642, 0.68642, 0.68642]), (0.5, [0.93391, 0.69643, 0.69243])):
        w = rtc.World.default_world()
        floor = S.plane()
        floor.set_transform(T.translation(0, -1, 0))
        floor.material.reflective, floor.material.transparency, floor.material.refractive_index = reflective, 0.5, 1.5
        w.push(floor)
        ball = S.sphere()
        ball.material.color, ball.material.ambient = (1, 0, 0), 0.5
        ball.set_transform(T.translation(0, -3.5, -0.5))
        w.push(ball)
        np.testing.assert_allclose(w.color_at(ray)[0], want, atol=1e-5)


def test_row_bands_tile_the_frame(rtc):
    """rtc_rows: cyclic row bands rendered separately reassemble into the full frame (the multi-GPU sharding)."""
    world, cam = rtc.build_scene("table", 200, 117)  # 117 rows: the last band is ragged
    full = np.empty((117, 200, 4), dtype=np.uint8)
    cam.render_into(world, rgba8=full)
    for band_rows, stride in ((8, 3), (16, 2), (5, 4)):
        out = np.zeros_like(full)
        for first in range(stride):
            rows = rtc.Rows(band_rows, first, stride)
            n = cam.rows_count(rows)
            part = np.empty((n, 200, 4), dtype=np.uint8)
            cam.render_into(world, rgba8=part, rows=rows)
            k = 0
            for b in range(first, (117 + band_rows - 1) // band_rows, stride):
                r0, r1 = b * band_rows, min(117, (b + 1) * band_rows)
                out[r0:r1] = part[k:k + r1 - r0]
                k += r1 - r0
            assert k == n
        assert np.array_equal(out, full)


def test_frame_layout_rows_write_in_place(rtc):
    """RTC_ROWS_FRAME: each band call stores its rows at their frame position in ONE shared buffer (what the multi-GPU
    peer path does with rank 0's frame mapped over NVLink)."""
    import torch
    world, cam = rtc.build_scene("hexagon", 200, 120)
    full = np.empty((120, 200, 4), dtype=np.uint8)
    cam.render_into(world, rgba8=full)
    frame = torch.zeros((120, 200, 4), dtype=torch.uint8, device="cuda:0")
    for first in range(3):
        cam.render_device(world, d_rgba8=frame.data_ptr(), rows=rtc.Rows(8, first, 3, rtc.Rows.FRAME),
                          stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(frame.cpu().numpy(), full)


@pytest.mark.parametrize("w,h,band_rows,stride", [(200, 117, 8, 3), (96, 1100, 8, 2), (64, 1043, 5, 3), (128, 600, 600, 1)])
def test_frame_layout_rows_to_a_host_frame(rtc, w, h, band_rows, stride):
    """rtc_render with RTC_ROWS_FRAME and HOST buffers: each call's bands are copied to their frame positions in one host
    frame (the copies of one rank of a sharded render into the shared canvas) and every other row is left alone; both the
    single-launch path and the chunked, overlapped one (>= 256 local rows), ragged last bands, RGBA8 and the f64 Canvas."""
    world, cam = rtc.build_scene("table", w, h)
    full8, full64 = np.empty((h, w, 4), dtype=np.uint8), np.empty((h, w, 3))
    cam.render_into(world, rgba8=full8, rgb_f64=full64)
    out8, out64 = np.full((h, w, 4), 0x5A, dtype=np.uint8), np.full((h, w, 3), -7.0)
    for first in range(stride):
        before8, before64 = out8.copy(), out64.copy()
        rows = rtc.Rows(band_rows, first, stride, rtc.Rows.FRAME)
        cam.render_into(world, rgba8=out8, rgb_f64=out64, rows=rows)
        mine = ((np.arange(h) // band_rows) % stride) == first
        assert np.array_equal(out8[mine], full8[mine]) and np.array_equal(out64[mine].view(np.uint64), full64[mine].view(np.uint64))
        assert np.array_equal(out8[~mine], before8[~mine]) and np.array_equal(out64[~mine], before64[~mine])
    assert np.array_equal(out8, full8) and np.array_equal(out64.view(np.uint64), full64.view(np.uint64))
    # one buffer only, with stats (a single timed launch)
    st = rtc.Stats()
    out8[:] = 0
    cam.render_into(world, rgba8=out8, rows=rtc.Rows(band_rows, 0, stride, rtc.Rows.FRAME), stats=st)
    mine = ((np.arange(h) // band_rows) % stride) == 0
    assert np.array_equal(out8[mine], full8[mine]) and not out8[~mine].any() and st.primary_rays == int(mine.sum()) * w


def test_shared_canvas_single_rank(rtc):
    """multi.SharedCanvasRenderer with one rank: the canvas lives in a page-locked POSIX shared-memory segment and holds the
    frame rtc_render produces, f64 colours and RGBA8 pixels, frame after frame."""
    import importlib
    multi = importlib.import_module("ray-tracer-challenge-rust_b200.multi")
    world, cam = rtc.build_scene("hexagon", 320, 160)
    full8, full64 = np.empty((160, 320, 4), dtype=np.uint8), np.empty((160, 320, 3))
    cam.render_into(world, rgba8=full8, rgb_f64=full64)
    r = multi.SharedCanvasRenderer(world, cam, 0, 1, 0, want_f64=True, want_rgba8=True)
    try:
        for _ in range(3):
            f64, rgba = r.render()
            assert np.array_equal(rgba, full8) and np.array_equal(f64.view(np.uint64), full64.view(np.uint64))
            f64[:] = 0
            rgba[:] = 0
        assert os.path.exists("/dev/shm" + r.name)
    finally:
        r.close()
    assert not os.path.exists("/dev/shm" + r.name)


def test_full_size_properties_8k(rtc):
    """BASELINE config 5 at its full 7680x4320: properties that need no oracle — rendering is idempotent, eight cyclic
    band shards (the 8-GPU decomposition) written in place reproduce the single-launch frame bit for bit, and the
    shards' exact ray counts add up to the frame's."""
    import torch
    w, h = 7680, 4320
    world, cam = rtc.build_scene("pumpkin", w, h)
    stream = torch.cuda.current_stream().cuda_stream
    a = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
    b = torch.zeros_like(a)
    st = rtc.Stats()
    cam.render_device(world, d_rgba8=a.data_ptr(), stream=stream, stats=st)
    assert st.primary_rays == w * h and st.refract_rays > 0 and st.reflect_rays > 0
    st2 = rtc.Stats()
    cam.render_device(world, d_rgba8=b.data_ptr(), stream=stream, stats=st2)
    assert torch.equal(a, b)
    assert (st.shadow_rays, st.reflect_rays, st.refract_rays) == (st2.shadow_rays, st2.reflect_rays, st2.refract_rays)
    b.zero_()
    tot = [0, 0, 0, 0]
    for r in range(8):
        s8 = rtc.Stats()
        cam.render_device(world, d_rgba8=b.data_ptr(), rows=rtc.Rows(8, r, 8, rtc.Rows.FRAME), stream=stream, stats=s8)
        for i, v in enumerate((s8.primary_rays, s8.shadow_rays, s8.reflect_rays, s8.refract_rays)):
            tot[i] += v
    assert torch.equal(a, b)
    assert tot == [st.primary_rays, st.shadow_rays, st.reflect_rays, st.refract_rays]
    assert int(a[..., 3].min()) == 255  # every pixel was written


@pytest.mark.parametrize("w,h", [(1280, 720), (2560, 1442), (4096, 2160)])
def test_host_output_chunks_match_single_launch(rtc, w, h):
    """rtc_render's chunked path (csrc/render.cu render_host: 2 launches — 3/4 and 1/4 of the rows — for an RGBA8 frame;
    when the f64 Canvas colours are wanted too, 4 launches — 1/8, 1/8, 1/4, 1/2 — below 2 Mpx (the first size here) and 8
    launches — weights 1,2,2,2,2,2,2,1 — from there on; each chunk's copy overlaps the next chunk's kernel; ragged last
    tile row included) returns the one-launch frame."""
    world, cam = rtc.build_scene("cow_teddy", w, h)
    one = np.empty((h, w, 4), dtype=np.uint8)
    two = np.zeros_like(one)
    f_one, f_two = np.empty((h, w, 3)), np.zeros((h, w, 3))
    cam.render_into(world, rgba8=one, rgb_f64=f_one, stats=rtc.Stats())  # stats requested -> single launch
    cam.render_into(world, rgba8=two)                                     # no stats -> chunked + overlapped copies
    assert np.array_equal(one, two)
    two[:] = 0
    cam.render_into(world, rgba8=two, rgb_f64=f_two)
    assert np.array_equal(one, two) and np.array_equal(f_one.view(np.uint64), f_two.view(np.uint64))
    f_two[:] = 0
    cam.render_into(world, rgb_f64=f_two)
    assert np.array_equal(f_one.view(np.uint64), f_two.view(np.uint64))


def test_device_ppm_encoder_is_byte_identical(rtc, oracle):
    """rtc_ppm_encode_device (to_ppm on the GPU) == the host encoder == the reference encoder, on random
    frames (every digit count, every wrap position), ragged sizes, and a rendered frame that never left the device."""
    import torch
    rng = np.random.default_rng(11)
    for w, h in ((1, 1), (5, 3), (23, 7), (24, 1), (640, 33), (1921, 5)):
        frame = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        frame[..., 3] = 255
        if w > 20:
            frame[0, :, :3] = 255   # a row of three-digit values
            frame[-1, :, :3] = 7    # a row of one-digit values
        d = torch.from_numpy(frame).cuda()
        got = rtc.ppm_from_device(d.data_ptr(), w, h)
        assert got == rtc.ppm_from_rgba8(frame, w, h)
        rgb = frame[..., :3].reshape(-1, 3).astype(np.float64) / 255.0
        assert got == oracle.ppm(rgb, w, h)
    world, cam = rtc.build_scene("table", 640, 360)
    buf = torch.zeros((360, 640, 4), dtype=torch.uint8, device="cuda:0")
    cam.render_device(world, d_rgba8=buf.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rtc.ppm_from_device(buf.data_ptr(), 640, 360) == cam.render(world).to_ppm()


def test_device_output_and_torch_stream(rtc):
    """rtc_render_device writes into torch-owned device memory on torch's current stream."""
    import torch
    world, cam = rtc.build_scene("hexagon", 256, 128)
    host = np.empty((128, 256, 4), dtype=np.uint8)
    cam.render_into(world, rgba8=host)
    buf = torch.zeros((128, 256, 4), dtype=torch.uint8, device="cuda:0")
    st = rtc.Stats()
    cam.render_device(world, d_rgba8=buf.data_ptr(), stream=torch.cuda.current_stream().cuda_stream, stats=st)
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy(), host)
    assert st.kernel_launches == 1 and st.device_ms > 0


def test_edge_sizes(rtc, oracle):
    """1x1, a single row, a single column, widths that are not a multiple of the 8x4 tile."""
    for w, h in ((1, 1), (37, 1), (1, 29), (13, 7)):
        world, cam = rtc.build_scene("table", w, h)
        canvas = cam.render(world)
        ow, oc = helpers.scenes.build(oracle, "table", w, h)
        ref, _ = oracle.render(ow, oc, mode=oracle.CACHED)
        _check(ref, canvas.pixels_f64().reshape(-1, 3), oracle.quantise_rgba8(ref), canvas.pixels_rgba8())


def test_empty_world_is_black(rtc):
    w = rtc.World(rtc.Light((0, 10, 0), (1, 1, 1)))
    c = rtc.Camera(16, 8, 1.0)
    assert not c.render(w).pixels_f64().any()


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h", [("teapot", 640, 360), ("cow_teddy", 960, 540), ("pumpkin", 1280, 720)])
def test_device_mesh_build_renders_the_same_frame(rtc, oracle, name, w, h):
    """RTC_BUILD_DEVICE_LBVH: triangle tables, normals and BVH built by csrc/lbvh.cu on the GPU.  Another tree, the same
    frame and the same exact ray counts as the host-built scene, and the oracle's pixels on a subset."""
    world, cam = rtc.build_scene(name, w, h)
    host = np.empty((h, w, 4), dtype=np.uint8)
    st_h = rtc.Stats()
    cam.render_into(world, rgba8=host, stats=st_h)
    info_h = world.scene_info()
    world.set_build("device")
    dev = np.zeros_like(host)
    st_d = rtc.Stats()
    cam.render_into(world, rgba8=dev, stats=st_d)
    info_d = world.scene_info()
    assert info_d["mesh_triangles"] == info_h["mesh_triangles"] and info_d["meshes"] == info_h["meshes"]
    assert np.array_equal(host, dev)
    assert (st_h.shadow_rays, st_h.reflect_rays, st_h.refract_rays) == (st_d.shadow_rays, st_d.reflect_rays, st_d.refract_rays)
    # f64 colours: the two builds agree bit for bit (same device arithmetic, only the set of boxes tested differs) ...
    rgb_d = np.empty((h, w, 3))
    cam.render_into(world, rgb_f64=rgb_d)
    world.set_build("host")
    rgb_h = np.empty((h, w, 3))
    cam.render_into(world, rgb_f64=rgb_h)
    assert np.array_equal(rgb_d.view(np.uint64), rgb_h.view(np.uint64))
    # ... and the oracle's pixels on a subset, to the usual bar
    px = helpers.subset_pixels(w, h, 32, 16)
    ow, oc = helpers.scenes.build(oracle, name, w, h)
    ref, _ = oracle.render(ow, oc, mode=oracle.CACHED, pixels=px)
    _check(ref, np.ascontiguousarray(rgb_d[px[:, 1], px[:, 0]]), oracle.quantise_rgba8(ref), dev[px[:, 1], px[:, 0]])
    # rebuilt from scratch, the device build is deterministic
    world.set_build("device")
    again = np.zeros_like(host)
    cam.render_into(world, rgba8=again)
    assert np.array_equal(dev, again)


@pytest.mark.gpu
def test_device_mesh_build_gathers_scattered_triangles(rtc):
    """Triangle payloads stored out of shape order: the device build gathers its input instead of reading the caller's
    array in place; both builds of both descriptions render one frame."""
    world, cam = rtc.build_scene("teapot", 320, 180)
    api = rtc.api()
    cdesc = cam.desc()
    frames = []
    with helpers.scattered_description(world) as (desc, scattered):
        for d in (desc, scattered):
            for flags in (rtc.RTC_BUILD_HOST_SAH, rtc.RTC_BUILD_DEVICE_LBVH):
                scene = C.c_void_p()
                api.check(api.scene_create_ex(C.cast(C.byref(d), C.c_void_p), 0, flags, C.byref(scene)))
                out = np.zeros((180, 320, 4), dtype=np.uint8)
                api.check(api.render(scene, C.byref(cdesc), None, out.ctypes.data_as(C.c_void_p), None, None))
                api.scene_destroy(scene)
                frames.append(out)
    for f in frames[1:]:
        assert np.array_equal(frames[0], f)


@pytest.mark.gpu
def test_device_build_falls_back_to_the_host_for_the_reference_panic(rtc):
    """bounds.rs:143 for a mesh with a non-finite vertex: the device gate fold flags it, the library rebuilds on the host
    and reports the reference's panic — the same error as the host build."""
    v, f = helpers.scenes.load_mesh("teapot")
    v = np.array(v, dtype=np.float64)
    v[17, 1] = np.inf
    for build in ("host", "device"):
        S = rtc.Shapes(rtc.api())
        w = rtc.World(rtc.Light((0, 5, -5), (1, 1, 1)))
        w.push(S.mesh(v, f))
        w.set_build(build)
        cam = rtc.Camera(32, 16, 0.8)
        with pytest.raises(rtc.RtcError) as e:
            cam.render(w)
        assert e.value.code == rtc.RTC_ERR_PANIC and "bounds.rs:143" in e.value.message


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_random_triangle_soups_under_both_builds(rtc, oracle, seed):
    """Glass and mirror triangle soups with duplicate and degenerate triangles (worldgen.random_soup_world): the host SAH
    scene and the device-built scene render the same frame with the same ray counts, which is the oracle's."""
    import worldgen
    w, c = worldgen.random_soup_world(rtc.api(), seed, 160, 112)
    world = rtc.World(_handle=w.h)
    w.h = None
    cam = rtc.Camera.__new__(rtc.Camera)
    cam.api, cam.hsize, cam.vsize, cam.field_of_view, cam.h = c.api, c.hsize, c.vsize, c.field_of_view, c.h
    c.h = None
    ow, oc = worldgen.random_soup_world(oracle, seed, 160, 112)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    frames = []
    for build in ("host", "device"):
        world.set_build(build)
        st = rtc.Stats()
        rgba = np.zeros((112, 160, 4), dtype=np.uint8)
        rgb = np.zeros((112, 160, 3))
        cam.render_into(world, rgba8=rgba, rgb_f64=rgb, stats=st)
        _check(ref, rgb.reshape(-1, 3), oracle.quantise_rgba8(ref), rgba.reshape(-1, 4))
        assert st.total_rays == cnt.total_rays
        frames.append(rgb)
    assert np.array_equal(frames[0].view(np.uint64), frames[1].view(np.uint64))


@pytest.mark.parametrize("name,w,h", [("table", 640, 363), ("cow_teddy", 320, 180)])
def test_multi_device_render_in_one_process(rtc, name, w, h):
    """rtc_multi_* / rtc_render_multi (csrc/multi.cu): the frame sharded over every GPU of the box from ONE process — both
    the pinned host frame (per-device strided copies) and the device-0 frame (peer stores) equal the single-GPU frame, ray
    counts included; a frame height that is not a multiple of the band height included."""
    import torch
    world, cam = rtc.build_scene(name, w, h)
    one = np.empty((h, w, 4), dtype=np.uint8)
    st1 = rtc.Stats()
    cam.render_into(world, rgba8=one, stats=st1)
    for ngpus in sorted({1, min(2, rtc.device_count()), rtc.device_count()}):
        mr = rtc.MultiRenderer(world, ngpus)
        st = rtc.Stats()
        assert np.array_equal(mr.render(cam, "host", stats=st), one), ngpus
        assert st.total_rays == st1.total_rays and st.kernel_launches == ngpus
        ptr = mr.render(cam, "device")

        class _Iface:
            __cuda_array_interface__ = {"shape": (h, w, 4), "typestr": "|u1", "data": (ptr, False), "version": 2}
        assert np.array_equal(torch.as_tensor(_Iface(), device="cuda:0").cpu().numpy(), one), ngpus
        mr.close()
        st = rtc.Stats()
        assert np.array_equal(rtc.render_multi(world, cam, ngpus, stats=st), one)
        assert st.total_rays == st1.total_rays


# ---- smooth triangles.  Parity UNPINNED against the reference, which does not implement them
# (intersection.rs:381-386, obj_file.rs:295-335 are commented-out scenarios): the oracle follows the book's definition.
@pytest.mark.parametrize("name,w,h,build", [("table", 640, 363, "host"), ("cow_teddy", 320, 300, "device")])
def test_drop_in_call_over_all_devices(rtc, name, w, h, build):
    """rtc_camera_render(device = RTC_DEVICE_ALL) and rtc_multi_render_host: Camera::render sharded over every device of
    this process, each device copying its bands into the one canvas — the f64 colours and the RGBA8 pixels are those of a
    single-device render, bit for bit, and the ray counts add up (on a one-GPU box this is device 0)."""
    world, cam = rtc.build_scene(name, w, h)
    world.set_build(build)
    st1, sta = rtc.Stats(), rtc.Stats()
    one = cam.render(world, want_f64=True, stats=st1, device=0)
    f_one, p_one = one.pixels_f64().copy(), one.pixels_rgba8().copy()
    for _ in range(2):  # the second call finds every device's scene cached
        allc = cam.render(world, want_f64=True, stats=sta, device=rtc.ALL_DEVICES)
        assert np.array_equal(allc.pixels_f64().view(np.uint64), f_one.view(np.uint64))
        assert np.array_equal(allc.pixels_rgba8(), p_one)
        assert (sta.primary_rays, sta.shadow_rays, sta.reflect_rays, sta.refract_rays) == \
               (st1.primary_rays, st1.shadow_rays, st1.reflect_rays, st1.refract_rays)
        assert sta.kernel_launches == rtc.device_count()
    world.drop_scenes()
    only8 = cam.render(world, want_f64=False, device=rtc.ALL_DEVICES)
    assert np.array_equal(only8.pixels_rgba8(), p_one)
    for ngpus in sorted({1, rtc.device_count()}):
        m = rtc.MultiRenderer(world, ngpus, build=build)
        out8, out64 = np.zeros((h, w, 4), dtype=np.uint8), np.zeros((h, w, 3))
        m.render_into(cam, rgba8=out8, rgb_f64=out64)
        assert np.array_equal(out8, p_one) and np.array_equal(out64.view(np.uint64), f_one.view(np.uint64))
        out64[:] = 0
        stm = rtc.Stats()
        m.render_into(cam, rgb_f64=out64, stats=stm)
        assert np.array_equal(out64.view(np.uint64), f_one.view(np.uint64)) and stm.total_rays == st1.total_rays
        m.close()


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("build", ["host", "device"])
def test_smooth_triangle_worlds_match_oracle(rtc, oracle, seed, build):
    """Vertex-normal meshes (one of glass), a run that mixes flat and smooth triangles, a mirror floor.  "device": the
    request for the GPU mesh build is honoured by falling back to the host build (vertex normals are placed by the host)."""
    import worldgen
    w, c = worldgen.smooth_world(rtc.api(), seed, hsize=96, vsize=64)
    world = rtc.World(_handle=w.h)
    w.h = None
    cam = rtc.Camera.__new__(rtc.Camera)
    cam.api, cam.hsize, cam.vsize, cam.field_of_view, cam.h = c.api, c.hsize, c.vsize, c.field_of_view, c.h
    c.h = None
    world.set_build(build)
    st = rtc.Stats()
    canvas = cam.render(world, stats=st)
    ow, oc = worldgen.smooth_world(oracle, seed, hsize=96, vsize=64)
    ref, cnt = oracle.render(ow, oc, mode=oracle.CACHED)
    _check(ref, canvas.pixels_f64().reshape(-1, 3), oracle.quantise_rgba8(ref), canvas.pixels_rgba8())
    assert (st.primary_rays, st.shadow_rays, st.reflect_rays, st.refract_rays) == \
        (cnt.primary, cnt.shadow, cnt.reflect, cnt.refract)


def test_config_4_with_smooth_triangles(rtc, oracle):
    """BASELINE config 4 (cow + teddy + reflective floor at 3840x2160) with every triangle a smooth triangle over generated
    vertex normals: the 1/64 pixel subset of the full-resolution camera against the oracle, and the frame differs from the
    flat-triangle render of the same scene."""
    w, h = 3840, 2160
    world, cam = rtc.build_scene("cow_teddy_smooth", w, h)
    assert world.kernel_features() == (2146, 2146)
    rgba = np.empty((h, w, 4), dtype=np.uint8)
    st = rtc.Stats()
    cam.render_into(world, rgba8=rgba, stats=st)
    assert st.primary_rays == w * h
    ow, oc = helpers.scenes.build(oracle, "cow_teddy_smooth", w, h)
    px = helpers.subset_pixels(w, h, 64, 32)
    ref, _ = oracle.render(ow, oc, mode=oracle.CACHED, pixels=px)
    _check(ref, None, oracle.quantise_rgba8(ref), rgba[px[:, 1], px[:, 0]])
    fworld, fcam = rtc.build_scene("cow_teddy", w, h)
    flat = np.empty((h, w, 4), dtype=np.uint8)
    fcam.render_into(fworld, rgba8=flat)
    assert (flat != rgba).any(axis=2).mean() > 0.02


def test_two_process_peer_exchange_renders_the_single_gpu_frame(rtc):
    """One process per GPU (multi.ShardedRenderer, peer exchange with completion counters): bench.py under torchrun at
    N = 2 on a small frame; its own sharded_frame_check compares the sharded frame with rank 0's single-GPU render byte for
    byte, after two dozen pipelined frames through both buffers.  Needs two GPUs (skipped on the one-GPU test box)."""
    if rtc.device_count() < 2:
        pytest.skip("needs two GPUs")
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "bench.py"), "--gpus", "2", "--steps", "8", "--warmup", "3",
           "--no-extras", "--workload", "cow_teddy", "--width", "640", "--height", "360"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().split("\n")[-1])
    assert line["n_gpus"] == 2 and line["config_detail"]["exchange"] == "peer"
    assert line["sharded_frame_check"]["identical_to_single_gpu_render"] is True
    # the e2e loops ran through the canvas in shared host memory, and it held the single-GPU render's f64 colours
    assert line["e2e"]["exchange"] == "shared host canvas" and line["sharded_frame_check"]["f64_canvas_identical"] is True
