#!/usr/bin/env python3
"""The reference binary's command line on the B200 path:  render_ppm.py <filename.ppm> [width-in-px]

`ray-tracer-challenge-rust <filename.ppm> [width]` (main.rs:39-81) renders the cow scene at width x width/2 (default
400) and writes a P3 PPM; this does the same through librtc_b200.so.  --scene picks any of the scene builders main.rs
holds (hexagon, table, cow, teapot) or the two synthetic BASELINE configs; --height overrides the 2:1 aspect.
"""
import argparse
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("filename")
    p.add_argument("width", nargs="?", type=int, default=400)
    p.add_argument("--height", type=int, default=None)
    p.add_argument("--scene", default="cow", choices=["hexagon", "table", "cow", "teapot", "cow_teddy", "pumpkin"])
    a = p.parse_args()
    rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
    h = a.height or a.width // 2  # Camera::new(width, width / 2, 0.785), main.rs:85,153,329,369
    world, cam = rtc.build_scene(a.scene, a.width, h)
    st = rtc.Stats()
    t0 = time.perf_counter()
    canvas = cam.render(world, want_f64=False, stats=st)
    t1 = time.perf_counter()
    ppm = canvas.to_ppm()
    t2 = time.perf_counter()
    try:
        with open(a.filename, "wb") as f:
            f.write(ppm)
    except OSError as e:  # main.rs:142-145
        print(f"Can't open {a.filename}: {e}")
        return 1
    print(f"{a.scene} {a.width}x{h}: {st.total_rays} rays, kernel {st.device_ms:.3f} ms, render call {1e3 * (t1 - t0):.1f} ms, "
          f"to_ppm {1e3 * (t2 - t1):.1f} ms, {len(ppm)} bytes -> {a.filename}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
