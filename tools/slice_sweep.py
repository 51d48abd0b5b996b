#!/usr/bin/env python3
"""One rank's share of a sharded frame on ONE GPU: kernel time of the 8-row bands a rank renders at N = 2, 4, 8 (band_first 0,
stride N) — what limits the 1 -> 8 curve of the small configs, without needing 8 GPUs.
    python tools/slice_sweep.py [scene w h]"""
import importlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
    capi = importlib.import_module("ray-tracer-challenge-rust_b200._capi")
    name, w, h = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("table", 1920, 1080)
    world, cam = rtc.build_scene(name, w, h)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    buf = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
    out = {"scene": f"{name} {w}x{h}"}
    for n in (1, 2, 4, 8):
        rows = capi.Rows(8, 0, n, capi.Rows.COMPACT)
        st = rtc.Stats()
        ms = []
        for _ in range(9):
            flush.zero_()
            cam.render_device(world, d_rgba8=buf.data_ptr(), rows=rows, stats=st)
            ms.append(st.device_ms)
        out[f"n{n}_ms"] = round(statistics.median(ms[2:]), 4)
    out["ideal_n8_ms"] = round(out["n1_ms"] / 8, 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
