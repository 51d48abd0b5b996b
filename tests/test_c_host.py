"""examples/rtc_main.c — main.rs restated in C over the C ABI — builds against include/rtc.h with a plain C compiler,
and on the GPU writes the same PPM bytes as the Python host (which the other tests pin against the oracle)."""
import os
import subprocess

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ray-tracer-challenge-rust_b200")


@pytest.fixture(scope="module")
def rtc_main(rtc, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cbin") / "rtc_main")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "rtc_main.c"), "-L" + LIBDIR, "-lrtc_b200", "-lm", "-o", exe],
                   check=True, capture_output=True)
    return exe


def _run(exe, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([exe, *args], env=env, capture_output=True, text=True)


def test_c_host_links_and_parses_its_command_line(rtc_main):
    p = _run(rtc_main)
    assert p.returncode == 0 and "Expected a filename argument!" in p.stdout  # main.rs:46-49
    p = _run(rtc_main, "a.ppm", "12", "extra")
    assert "too many arguments" in p.stdout                                   # main.rs:52-55
    p = _run(rtc_main, "a.ppm", "wide")
    assert "not number" in p.stderr                                           # main.rs:70-73


def _write_obj(path, name):
    v, f = helpers.scenes.load_mesh(name)
    with open(path, "w") as out:
        for x, y, z in v:
            out.write(f"v {float(x)!r} {float(y)!r} {float(z)!r}\n")
        for a, b, c in f:
            out.write(f"f {a} {b} {c}\n")


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["hexagon", "table", "cow", "teapot"])
def test_c_host_writes_the_same_ppm(rtc, rtc_main, tmp_path, scene):
    objs = tmp_path / "objs"
    objs.mkdir()
    if scene == "cow":
        _write_obj(objs / "cow-nonormals.obj", "cow")
    if scene == "teapot":
        _write_obj(objs / "teapot.obj", "teapot")
    out = tmp_path / "frame.ppm"
    p = _run(rtc_main, str(out), "160", "--scene", scene, "--objs", str(objs))
    assert p.returncode == 0, p.stderr
    world, cam = rtc.build_scene(scene, 160, 80)
    assert out.read_bytes() == cam.render(world).to_ppm()
