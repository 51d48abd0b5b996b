#!/usr/bin/env python3
"""Build launch-shape variants of librtc_b200.so here (no GPU needed), then time them on the GPU box:

    python tools/tune_variants.py build            # on the CPU box: writes ray-tracer-challenge-rust_b200/variants/*.so
    gpurun -- python tools/tune_variants.py run    # on the B200: times every variant on every config

Each variant is the same source with other -D macros for render.cu (block size, CTAs per SM -> register cap)."""
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "ray-tracer-challenge-rust_b200", "variants")
VARIANTS = {
    "t128_b5": ["RTC_BLOCKS_PER_SM=5", "RTC_BLOCKS_PER_SM_PRIMS=5"],
    "t128_b6": ["RTC_BLOCKS_PER_SM=6", "RTC_BLOCKS_PER_SM_PRIMS=6"],
    "t128_b8": ["RTC_BLOCKS_PER_SM=8", "RTC_BLOCKS_PER_SM_PRIMS=8"],
}
SCENES = [("table", 1920, 1080), ("teapot", 1920, 1080), ("hexagon", 1920, 960), ("cow_teddy", 3840, 2160),
          ("pumpkin", 3840, 2160)]


def build():
    b = importlib.import_module("ray-tracer-challenge-rust_b200.build")
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        out = os.path.join(VDIR, f"librtc_{name}.so")
        b.build(force=True, defines=defs, out=out)
        log = open(os.path.join(b.BUILD, os.path.basename(out) + ".d", "build.log")).read()
        regs = [l for l in log.split("\n") if "render_kernel" in l or "Used" in l]
        for i, l in enumerate(regs):
            if "render_kernel" in l:
                print(name, [x.strip() for x in regs[i + 1:i + 3]])
                break


def run_one():
    import torch
    rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
    out = {}
    for name, w, h in SCENES:
        world, cam = rtc.build_scene(name, w, h)
        buf = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda:0")
        st = rtc.Stats()
        ms = []
        for _ in range(6):
            cam.render_device(world, d_rgba8=buf.data_ptr(), stats=st)
            ms.append(st.device_ms)
        out[name] = min(ms[1:])
    print(json.dumps(out))


def run():
    res = {}
    for name in VARIANTS:
        lib = os.path.join(VDIR, f"librtc_{name}.so")
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, RTC_B200_LIB=lib)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True)
        try:
            res[name] = json.loads(p.stdout.strip().split("\n")[-1])
        except Exception:
            res[name] = {"error": (p.stdout + p.stderr)[-400:]}
        print(name, res[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)


if __name__ == "__main__":
    {"build": build, "run": run, "one": run_one}[sys.argv[1]]()
