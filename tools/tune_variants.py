#!/usr/bin/env python3
"""A/B harness for kernel variants: build variants of librtc_b200.so here (no GPU needed), then time them on the GPU box.

    python tools/tune_variants.py build name=DEF1,DEF2 other=DEF3 ...   # CPU box: ray-tracer-challenge-rust_b200/variants/*.so
    gpurun -- python tools/tune_variants.py run [rounds]                # B200: every variant x every config, alternating

`base` (no extra defines) is always built.  Each variant is the same source with extra -D macros; `run` renders every
config with every variant in alternating rounds (so clock / thermal drift hits all variants alike), checks that every
variant's frame is byte-identical to base's, and writes gpurun_out/variants.json: per variant and config the best and
median kernel time (CUDA events around the one render_kernel launch, L2 flushed by a 256 MiB memset between launches)."""
import importlib
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "ray-tracer-challenge-rust_b200", "variants")
SCENES = [("table", 1920, 1080), ("hexagon", 1920, 960), ("teapot", 1920, 1080), ("cow_teddy", 3840, 2160),
          ("pumpkin", 7680, 4320)]


def build(specs):
    b = importlib.import_module("ray-tracer-challenge-rust_b200.build")
    os.makedirs(VDIR, exist_ok=True)
    for f in os.listdir(VDIR):
        if f.endswith(".so"):
            os.remove(os.path.join(VDIR, f))
    variants = {"base": []}
    for spec in specs:
        name, _, defs = spec.partition("=")
        variants[name] = [d for d in defs.split(",") if d]
    for name, defs in variants.items():
        out = os.path.join(VDIR, f"librtc_{name}.so")
        b.build(force=True, defines=defs, out=out)
        log = open(os.path.join(b.BUILD, os.path.basename(out) + ".d", "build.log")).read().split("\n")
        regs = []
        for i, l in enumerate(log):
            if "render_kernel" in l and "Compiling" in l:
                regs.append(" ".join(x.strip() for x in log[i + 1:i + 4] if "Used" in x or "spill" in x))
        print(name, defs, "|", regs[0] if regs else "")
    json.dump(variants, open(os.path.join(VDIR, "variants.json"), "w"))


def run_one(rounds):
    import hashlib
    import torch
    rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    out = {}
    for name, w, h in SCENES:
        world, cam = rtc.build_scene(name, w, h)
        buf = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
        st = rtc.Stats()
        ms = []
        for _ in range(rounds + 1):
            flush.zero_()
            cam.render_device(world, d_rgba8=buf.data_ptr(), stats=st)
            ms.append(st.device_ms)
        digest = hashlib.sha1(buf.cpu().numpy().tobytes()).hexdigest()[:16]
        out[name] = {"ms": ms[1:], "frame": digest, "rays": st.total_rays}
    print(json.dumps(out))


def run(rounds, passes=3):
    variants = json.load(open(os.path.join(VDIR, "variants.json")))
    res = {name: {} for name in variants}
    for p in range(passes):
        for name in variants:
            lib = os.path.join(VDIR, f"librtc_{name}.so")
            env = dict(os.environ, RTC_B200_LIB=lib)
            q = subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(rounds)], env=env, capture_output=True,
                               text=True)
            try:
                one = json.loads(q.stdout.strip().split("\n")[-1])
            except Exception:
                res[name]["error"] = (q.stdout + q.stderr)[-600:]
                print(name, "ERROR", res[name]["error"], flush=True)
                continue
            for scene, v in one.items():
                e = res[name].setdefault(scene, {"ms": [], "frame": v["frame"], "rays": v["rays"]})
                e["ms"] += v["ms"]
                if e["frame"] != v["frame"]:
                    e["frame"] = "UNSTABLE"
    table = {}
    for name, scenes in res.items():
        for scene, v in scenes.items():
            if scene == "error":
                continue
            same = v["frame"] == res["base"].get(scene, {}).get("frame") and v["rays"] == res["base"][scene]["rays"]
            table.setdefault(scene, {})[name] = {"best_ms": min(v["ms"]), "median_ms": statistics.median(v["ms"]),
                                                 "frame_equals_base": same}
    for scene, row in table.items():
        print(scene, {k: (round(v["best_ms"], 4), round(v["median_ms"], 4), "ok" if v["frame_equals_base"] else "DIFFERS")
                      for k, v in row.items()}, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"defines": variants, "results": table}, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "build":
        build(sys.argv[2:])
    elif cmd == "run":
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 6)
    elif cmd == "one":
        run_one(int(sys.argv[2]))
