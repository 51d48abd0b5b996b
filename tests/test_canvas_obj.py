"""Canvas / to_ppm (canvas.rs tests :75-175) and the OBJ parser (obj_file.rs tests :139-293) of the host mirror."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

import helpers

capi = importlib.import_module("ray-tracer-challenge-rust_b200._capi")


def test_creating_a_canvas_and_writing_pixels(rtc):
    c = rtc.Canvas(10, 20)
    assert (c.width, c.height) == (10, 20)
    assert not c.pixels_f64().any()
    c.set_pixel(2, 3, (1, 0, 0))
    assert list(c.get_pixel(2, 3)) == [1, 0, 0]
    with pytest.raises(rtc.RtcError):
        c.get_pixel(10, 0)


def test_ppm_header_pixel_data_and_terminator(rtc):
    c = rtc.Canvas(5, 3)
    lines = c.to_ppm().decode().split("\n")
    assert lines[:3] == ["P3", "5 3", "255"] and lines[-1] == ""
    c.set_pixel(0, 0, (1.5, 0, 0))
    c.set_pixel(2, 1, (0, 0.5, 0))
    c.set_pixel(4, 2, (-0.5, 0, 1))
    lines = c.to_ppm().decode().split("\n")
    assert len(lines) == 7
    assert lines[3] == "255 0 0 0 0 0 0 0 0 0 0 0 0 0 0"
    assert lines[4] == "0 0 0 0 0 0 0 128 0 0 0 0 0 0 0"
    assert lines[5] == "0 0 0 0 0 0 0 0 0 0 0 0 0 0 255"


def test_splitting_long_lines(rtc):
    c = rtc.Canvas(10, 2)
    for y in range(2):
        for x in range(10):
            c.set_pixel(x, y, (1, 0.8, 0.6))
    lines = c.to_ppm().decode().split("\n")
    assert len(lines) == 8
    assert lines[3] == lines[5] == "255 204 153 255 204 153 255 204 153 255 204 153 255 204 153 255 204"
    assert lines[4] == lines[6] == "153 255 204 153 255 204 153 255 204 153 255 204 153"


def test_ppm_matches_oracle_on_random_frames(rtc, oracle):
    rng = np.random.default_rng(3)
    for w, h in ((1, 1), (7, 3), (23, 5), (64, 9), (301, 4)):
        rgb = rng.uniform(-0.2, 1.2, (h * w, 3))
        rgb[rng.integers(0, h * w, 3)] = np.nan  # NaN quantises to 0 (`as i32`)
        rgba = oracle.quantise_rgba8(rgb)
        assert rtc.ppm_from_rgba8(rgba, w, h) == oracle.ppm(rgb, w, h)
        c = rtc.Canvas(w, h)
        for i in range(h * w):
            c.set_pixel(i % w, i // w, rgb[i])
        assert c.to_ppm() == oracle.ppm(rgb, w, h)
        assert np.array_equal(c.pixels_rgba8().reshape(-1, 4), rgba)


TRIANGLES_OBJ = "v -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\n\ng FirstGroup\nf 1 2 3\ng SecondGroup\nf 1 3 4\n"


def _leaves(rtc, shape):
    api = rtc.api()
    w = rtc.World(rtc.Light((0, 0, 0), (1, 1, 1)))
    w.push(shape)
    m = C.c_void_p()
    api.check(api.world_marshal(w.h, C.byref(m)))
    d = C.cast(api.marshalled_desc(m), C.POINTER(capi.SceneDesc)).contents
    shapes = [(d.shapes[i].kind, d.shapes[i].child_count, d.shapes[i].triangle) for i in range(d.shape_count)]
    tris = [np.array([d.triangles[i].p1[:], d.triangles[i].p2[:], d.triangles[i].p3[:], d.triangles[i].e1[:],
                      d.triangles[i].e2[:], d.triangles[i].normal[:]]) for i in range(d.triangle_count)]
    api.marshalled_free(m)
    return shapes, tris


def test_obj_ignoring_unrecognized_lines(rtc):
    S = rtc.Shapes(rtc.api())
    g = S.obj_str("\nThere was a young lady named Bright\nwho traveled much faster than light.\nShe set out one day\n"
                  "in a relative way,\nand came back the previous night.\n")
    assert g.ignored_lines == 5 and g.leaf_count() == 0


def test_obj_faces_fan_and_groups(rtc):
    S = rtc.Shapes(rtc.api())
    g = S.obj_str("v -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\nv 0 2 0\n\nf 1 2 3 4 5\n")
    shapes, tris = _leaves(rtc, g)
    assert [s[0] for s in shapes] == [5, 5, 6, 6, 6]  # group{default_group{3 triangles}}
    v = np.array([[-1, 1, 0], [-1, 0, 0], [1, 0, 0], [1, 1, 0], [0, 2, 0]], dtype=float)
    for t, (a, b, c) in zip(tris, ((0, 1, 2), (0, 2, 3), (0, 3, 4))):
        assert np.array_equal(t[0], v[a]) and np.array_equal(t[1], v[b]) and np.array_equal(t[2], v[c])
        assert np.array_equal(t[3], v[b] - v[a]) and np.array_equal(t[4], v[c] - v[a])
    # shape.rs:1545-1562: normal = normalize(e2 x e1)
    shapes, tris = _leaves(rtc, S.triangle((0, 1, 0), (-1, 0, 0), (1, 0, 0)))
    assert np.array_equal(tris[0][3], [-1, -1, 0]) and np.array_equal(tris[0][4], [1, -1, 0])
    assert np.array_equal(np.abs(tris[0][5]), [0, 0, 1]) and tris[0][5][2] == -1
    g = S.obj_str(TRIANGLES_OBJ)
    shapes, tris = _leaves(rtc, g)
    assert [s[:2] for s in shapes] == [(5, 3), (5, 0), (5, 1), (6, 0), (5, 1), (6, 0)]
    assert g.ignored_lines == 0


def test_obj_parser_panics(rtc):
    S = rtc.Shapes(rtc.api())
    for bad in ("v 1 2\n", "v 1 2 x\n", "f 1/2/3 2 3\n", "v 0 0 0\nf 1 2 9\n", "g\n", "f 1\n"):
        with pytest.raises(ValueError):
            S.obj_str(bad)


def test_obj_matches_mesh_from_arrays_and_oracle(rtc, oracle, hostsim, tmp_path):
    """Parsing OBJ text == building from arrays; product parser and oracle parser render identically."""
    v, f = helpers.scenes.load_mesh("teddy")
    text = "".join(f"v {repr(float(a))} {repr(float(b))} {repr(float(c))}\n" for a, b, c in v[:400])
    faces = f[(f <= 400).all(axis=1)][:300]
    text += "# comment\n" + "".join(f"f {a} {b} {c}\n" for a, b, c in faces)
    p = tmp_path / "part.obj"
    p.write_text(text)
    T = rtc.Transformations(rtc.api())
    outs = []
    for make in (lambda S: S.obj_file(p), lambda S: S.obj_str(text), lambda S: S.mesh(v[:400], faces)):
        for api in (rtc.api(), oracle):
            S, Tx = rtc.Shapes(api), rtc.Transformations(api)
            sa = importlib.import_module("ray-tracer-challenge-rust_b200.scene_api")
            g = make(S)
            g.set_transform(Tx.scaling(0.1, 0.1, 0.1))
            w = sa.WorldHandle(api, rtc.Light((0, 6.9, -5), (1, 1, 0.9)))
            w.push(g)
            c = sa.CameraHandle(api, 40, 30, 0.8)
            c.set_transform(Tx.view_transform((0, 1, -8), (0, 0, 0), (0, 1, 0)))
            if api is oracle:
                outs.append(oracle.render(w, c, mode=oracle.CACHED)[0])
            else:
                world = rtc.World(_handle=w.h)
                w.h = None
                cam = rtc.Camera.__new__(rtc.Camera)
                cam.api, cam.hsize, cam.vsize, cam.field_of_view, cam.h = c.api, 40, 30, 0.8, c.h
                c.h = None
                outs.append(hostsim.scene(world).render(cam)[0])
    assert outs[0].any()
    for o in outs[1:]:
        assert np.array_equal(o.view(np.uint64), outs[0].view(np.uint64))


def test_obj_numbers_are_correctly_rounded(rtc):
    """Vertex coordinates parse like str::parse::<f64> (correctly rounded): the short-decimal fast path (<= 15 digits, no
    exponent) and the general path (long mantissas, exponents, inf) both give Python's float() bit for bit."""
    import ctypes as C
    import importlib
    capi = importlib.import_module("ray-tracer-challenge-rust_b200._capi")
    rng = np.random.default_rng(11)
    toks = ["0", "-0", "-0.000000", "+1", "1.", ".5", "-.25", "007.500", "123456789012345", "0.000000000000001",
            "1234567.12345678", "9007199254740993", "0.1234567890123456789", "1e-3", "-2.5E+2", "1e22", "1e23",
            "179769313486231570000000000000000000000", "4.9e-324", "inf", "-Infinity", "3.141592653589793"]
    for _ in range(400):
        nd = int(rng.integers(1, 19))
        digits = "".join(str(int(d)) for d in rng.integers(0, 10, nd))
        cut = int(rng.integers(0, nd + 1))
        t = digits[:cut] + ("." + digits[cut:] if rng.random() < 0.8 else digits[cut:])
        if t in (".", ""):
            t = "1"
        if rng.random() < 0.5:
            t = "-" + t
        if rng.random() < 0.1:
            t += "e" + str(int(rng.integers(-30, 30)))
        toks.append(t)
    while len(toks) % 3:
        toks.append("1")
    lines = [f"v {toks[i]} {toks[i + 1]} {toks[i + 2]}" for i in range(0, len(toks), 3)]
    nv = len(lines)
    lines += [f"f {i + 1} {(i + 1) % nv + 1} {(i + 2) % nv + 1}" for i in range(nv)]
    S = rtc.Shapes(rtc.api())
    g = S.obj_str("\n".join(lines) + "\n")
    w = rtc.World(rtc.Light((0, 0, 0), (1, 1, 1)))
    w.push(g)
    api = rtc.api()
    m = C.c_void_p()
    api.check(api.world_marshal(w.h, C.byref(m)))
    try:
        desc = C.cast(api.marshalled_desc(m), C.POINTER(capi.SceneDesc)).contents
        assert desc.triangle_count == nv
        for i in range(nv):  # triangle i starts at vertex i
            got = np.array(list(desc.triangles[i].p1))
            want = np.array([float(toks[3 * i]), float(toks[3 * i + 1]), float(toks[3 * i + 2])])
            assert got.tobytes() == want.tobytes(), (lines[i], got, want)
    finally:
        api.marshalled_free(m)


def test_wrap_rule_in_prefix_form_equals_the_sequential_rule(rtc):
    """csrc/ppm_encode.cu runs Canvas::to_ppm's 70-column wrap (canvas.rs:44-55) in PREFIX form, 32 tokens per step: a token
    of n digits costs a = n + 1 characters; with P the running sum of a over a row's tokens, a line that starts at token s
    holds the tokens t with P_t - P_{s-1} <= 71, starts at byte P_{s-1} of the row, and the row (with its final newline) is
    P_last bytes.  Restated here step for step (chunks of 32, the first token over the limit opens a line) and compared
    with the sequential host encoder on every digit count and wrap position; the kernels themselves are compared byte for
    byte on the GPU (tests/test_gpu_parity.py::test_device_ppm_encoder_is_byte_identical)."""
    def prefix_form(rgba):
        h, w, _ = rgba.shape
        out = bytearray(f"P3\n{w} {h}\n255\n".encode())
        for y in range(h):
            vals = rgba[y, :, :3].reshape(-1).astype(np.int64)
            ntok = len(vals)
            a_all = 2 + (vals >= 10) + (vals >= 100)
            run, base, lines = 0, 0, [(0, 0)]
            for t0 in range(0, ntok, 32):  # pass 1: a warp per row
                a = np.zeros(32, dtype=np.int64)
                m = min(32, ntok - t0)
                a[:m] = a_all[t0:t0 + m]
                P = run + np.cumsum(a)
                run, start = int(P[31]), 0
                while True:
                    over = [l for l in range(start, m) if P[l] - base > 71]
                    if not over:
                        break
                    j = over[0]
                    base = int(P[j] - a[j])
                    lines.append((t0 + j, base))
                    start = j + 1
            row = bytearray(run)
            for k, (s, off) in enumerate(lines):  # pass 3: a warp per line, a lane per token
                end = lines[k + 1][0] if k + 1 < len(lines) else ntok
                assert end - s <= 35
                before = 0
                for t in range(s, end):
                    q = off + before
                    if t != s:
                        row[q - 1] = 32
                    digits = str(int(vals[t])).encode()
                    row[q:q + len(digits)] = digits
                    if t + 1 == end:
                        row[q + len(digits)] = 10
                    before += int(a_all[t])
            out += row
        return bytes(out)

    rng = np.random.default_rng(5)
    for h, w in ((3, 5), (2, 17), (4, 23), (2, 24), (3, 36), (2, 100), (1, 1), (2, 12), (1, 641)):
        for mode in range(4):
            if mode == 0:
                img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            elif mode == 1:
                img = rng.integers(0, 10, (h, w, 4), dtype=np.uint8)
            elif mode == 2:
                img = np.full((h, w, 4), 255, dtype=np.uint8)
            else:
                img = rng.choice(np.array([0, 9, 10, 99, 100, 255], dtype=np.uint8), (h, w, 4))
            assert prefix_form(img) == rtc.ppm_from_rgba8(img, w, h), (h, w, mode)
