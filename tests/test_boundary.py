"""The drop-in boundary as files: rtc-sys (the Rust FFI crate north_star asks for), integration/reference.patch and the one
translation-unit manifest both builds read.  No Rust toolchain exists in this image, so what can be — and is — checked
here is that the crate's declarations agree with include/rtc.h and with the built library, symbol for symbol and field for
field, that build.rs cannot drift from build.py, and that the patch applies to the reference."""
import ctypes as C
import importlib
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rtc.h")
LIB_RS = os.path.join(ROOT, "rtc-sys", "src", "lib.rs")
BUILD_RS = os.path.join(ROOT, "rtc-sys", "build.rs")
CSRC = os.path.join(ROOT, "ray-tracer-challenge-rust_b200", "csrc")

C_TO_RUST = {"double": "f64", "int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "uint8_t": "u8", "int": "c_int",
             "char": "c_char", "void": "c_void"}
SIZES = {"f64": 8, "i32": 4, "u32": 4, "u64": 8, "u8": 1, "c_int": 4, "c_char": 1}


def _strip_c(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return "\n".join(ln for ln in text.splitlines() if not ln.strip().startswith("#"))


def _rust_type(ctype, name_suffix=""):
    """'const rtc_shape_desc*' -> '*const rtc_shape_desc'; arrays ('[3]' suffix) as [T; 3] (fields) — see callers."""
    t = ctype.strip()
    const = t.startswith("const ")
    if const:
        t = t[len("const "):].strip()
    stars = t.count("*")
    base = t.replace("*", "").strip()
    base = C_TO_RUST.get(base, base)
    out = base
    for k in range(stars):
        out = ("*const " if const and k == 0 else "*mut ") + out
    return out


def parse_header():
    text = _strip_c(open(HEADER).read())
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            head, *more = [d.strip() for d in decl.split(",")]
            mm = re.match(r"(.*?)(\w+)\s*(\[\d+\])?$", head)
            ctype = mm.group(1).strip()
            names = [(mm.group(2), mm.group(3))] + [re.match(r"(\w+)\s*(\[\d+\])?$", x).groups() for x in more]
            for name, arr in names:
                rt = _rust_type(ctype)
                if arr:
                    rt = f"[{rt}; {arr[1:-1]}]"
                fields.append((name, rt))
        structs[m.group(3)] = fields
    funcs = {}
    flat = " ".join(text.split())
    flat = re.sub(r"typedef struct \w+ \{.*?\} \w+ ;", " ", flat)
    for m in re.finditer(r"([\w][\w\s\*]*?)\b(rtc_\w+)\s*\(([^()]*)\)\s*;", flat):
        ret = m.group(1).strip()
        if ret.startswith("typedef") or "struct" in ret:
            continue
        args = []
        body = m.group(3).strip()
        if body and body != "void":
            for a in body.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)\s*(\[\d+\])?$", a)
                ctype, arr = mm.group(1).strip(), mm.group(3)
                if arr:  # an array parameter is a pointer
                    ctype += "*"
                args.append((mm.group(2), _rust_type(ctype)))
        funcs[m.group(2)] = (None if ret == "void" else _rust_type(ret), args)
    return structs, funcs


def parse_rust():
    text = re.sub(r"//[^\n]*", " ", open(LIB_RS).read())
    structs = {}
    for m in re.finditer(r"#\[repr\(C\)\][^{;]*?pub struct (\w+)\s*\{(.*?)\}", text, flags=re.S):
        fields = []
        for f in m.group(2).split(","):
            f = " ".join(f.split())
            if not f:
                continue
            # `[f64; 3]` contains no comma, so a plain split is enough
            mm = re.match(r"(?:pub )?(\w+)\s*:\s*(.+)$", f)
            fields.append((mm.group(1), mm.group(2).strip()))
        structs[m.group(1)] = fields
    ext = re.search(r'extern "C" \{(.*?)\n\}', text, flags=re.S).group(1)
    funcs = {}
    for m in re.finditer(r"pub fn (\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+?))?\s*;", ext, flags=re.S):
        args = []
        for a in m.group(2).split(","):
            a = " ".join(a.split())
            if a:
                n, t = a.split(":", 1)
                args.append((n.strip(), t.strip()))
        funcs[m.group(1)] = (m.group(3).strip() if m.group(3) else None, args)
    return structs, funcs


def _sizeof(rt, structs):
    m = re.match(r"\[(.+); (\d+)\]$", rt)
    if m:
        return _sizeof(m.group(1), structs) * int(m.group(2))
    if rt.startswith("*"):
        return 8
    if rt in SIZES:
        return SIZES[rt]
    return _layout(structs[rt], structs)[0]


def _layout(fields, structs):
    """(size, alignment) of a #[repr(C)] struct by the C rules."""
    off, align = 0, 1
    for _, rt in fields:
        m = re.match(r"\[(.+); (\d+)\]$", rt)
        elem = m.group(1) if m else rt
        a = 8 if elem.startswith("*") else (SIZES[elem] if elem in SIZES else _layout(structs[elem], structs)[1])
        off = (off + a - 1) // a * a + _sizeof(rt, structs)
        align = max(align, a)
    return (off + align - 1) // align * align, align


def test_rtc_sys_structs_match_the_header():
    hs, _ = parse_header()
    rs, _ = parse_rust()
    opaque = {"rtc_scene", "rtc_multi"}  # opaque handles are zero-sized markers on the Rust side
    for name, fields in rs.items():
        if name in opaque:
            continue
        assert name in hs, f"rtc-sys declares {name}, include/rtc.h does not"
        assert fields == hs[name], f"{name}: fields differ\n rust   {fields}\n header {hs[name]}"
    # every struct that crosses the CORE boundary is bound
    for name in ("rtc_material", "rtc_transform_desc", "rtc_triangle_desc", "rtc_vertex_normals", "rtc_shape_desc",
                 "rtc_scene_desc", "rtc_camera_desc", "rtc_rows", "rtc_stats", "rtc_computations"):
        assert name in rs, name
    # and its size is the one the ctypes host (which IS exercised on the GPU) uses
    capi = importlib.import_module("ray-tracer-challenge-rust_b200._capi")
    for name, ct in (("rtc_material", capi.Material), ("rtc_camera_desc", capi.CameraDesc), ("rtc_rows", capi.Rows),
                     ("rtc_stats", capi.Stats), ("rtc_scene_desc", capi.SceneDesc), ("rtc_shape_desc", capi.ShapeDesc),
                     ("rtc_triangle_desc", capi.TriangleDesc), ("rtc_transform_desc", capi.TransformDesc),
                     ("rtc_computations", capi.Computations), ("rtc_vertex_normals", capi.VertexNormals)):
        assert _layout(rs[name], rs)[0] == C.sizeof(ct), name
        assert [n for n, _ in rs[name]] == [n for n, _ in ct._fields_], name


def test_rtc_sys_functions_match_the_header_and_the_library(rtc):
    _, hf = parse_header()
    _, rf = parse_rust()
    lib = rtc.api().lib
    for name, (ret, args) in rf.items():
        assert name in hf, f"rtc-sys binds {name}, which include/rtc.h does not declare"
        hret, hargs = hf[name]
        assert ret == hret, f"{name}: returns {ret} in rtc-sys, {hret} in the header"
        assert [t for _, t in args] == [t for _, t in hargs], f"{name}: parameters differ\n rust   {args}\n header {hargs}"
        assert hasattr(lib, name), f"librtc_b200.so does not export {name}"
    # the whole CORE boundary (layer 1 of the header) is bound: everything that takes or makes an rtc_scene / rtc_multi
    core = [n for n, (ret, args) in hf.items()
            if any("rtc_scene" in t or "rtc_multi" in t for _, t in args) or n in ("rtc_last_error", "rtc_device_count")]
    missing = [n for n in core if n not in rf and not n.startswith(("rtc_world_", "rtc_marshalled_"))]
    assert not missing, missing


def test_every_declared_function_is_exported(rtc):
    _, hf = parse_header()
    lib = rtc.api().lib
    assert len(hf) > 70
    for name in hf:
        assert hasattr(lib, name), f"include/rtc.h declares {name}, librtc_b200.so does not export it"


def test_one_manifest_drives_both_builds():
    build = importlib.import_module("ray-tracer-challenge-rust_b200.build")
    man = build.read_manifest()
    listed = {f for _, f, _ in man["units"]}
    on_disk = {f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp"))}
    assert listed == on_disk, f"csrc/manifest.txt and csrc/ disagree: {sorted(listed ^ on_disk)}"
    assert any(d for _, f, d in man["units"] if f == "render_inst.cu" and d and d.startswith("RTC_INST_MASK="))
    assert "-fmad=false" in man["nvccflags"] and "arch=compute_100a,code=sm_100a" in man["nvccflags"]
    assert "-ffp-contract=off" in man["hostflags"]
    rs = open(BUILD_RS).read()
    assert "manifest.txt" in rs
    # build.rs names no translation unit itself: nothing to forget when one is added
    assert not re.findall(r'"\w+\.(?:cu|cpp)"', rs)
    py = open(os.path.join(ROOT, "ray-tracer-challenge-rust_b200", "build.py")).read()
    assert not re.findall(r'"\w+\.(?:cu|cpp)"', py)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree is not on this box")
def test_reference_patch_applies(tmp_path):
    if not shutil.which("patch"):
        pytest.skip("patch(1) is not installed")
    work = tmp_path / "ref"
    (work / "src").mkdir(parents=True)
    for f in ("Cargo.toml", "src/camera.rs", "src/pattern.rs", "src/main.rs"):
        shutil.copy(os.path.join("/root/reference", f), work / f)
    p = subprocess.run(["patch", "-p1", "--dry-run", "-i", os.path.join(ROOT, "integration", "reference.patch")],
                       cwd=work, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    subprocess.run(["patch", "-p1", "-i", os.path.join(ROOT, "integration", "reference.patch")], cwd=work, check=True,
                   capture_output=True)
    cam = (work / "src" / "camera.rs").read_text()
    assert "crate::gpu::render(world, &camera)" in cam and "world.color_at(&ray)" not in cam.split("#[cfg(test)]")[0]
    assert (work / "src" / "gpu.rs").exists() and "rtc-sys" in (work / "Cargo.toml").read_text()
    # the patch binds only what rtc-sys declares
    used = set(re.findall(r"\b(rtc_\w+)\b", (work / "src" / "gpu.rs").read_text() + cam))
    rs, rf = parse_rust()
    assert used <= set(rs) | set(rf) | {"rtc_sys"}, used - set(rs) - set(rf)
