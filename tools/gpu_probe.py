#!/usr/bin/env python3
"""Quick on-GPU timing table: every BASELINE config at its resolution, kernel ms (CUDA events) and Mrays/s."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

rtc = importlib.import_module("ray-tracer-challenge-rust_b200")


def main():
    out = {"fp64_peak_gflops": dict(zip(("nofma", "fma"), rtc.measure_fp64_peak(0)))}
    print(out, flush=True)
    for name, (w, h) in rtc.scenes.CONFIGS.items():
        if name == "hexagon":
            w, h = 1920, 960
        world, cam = rtc.build_scene(name, w, h)
        info = world.scene_info()
        import torch
        buf = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda:0")
        ms = []
        st = rtc.Stats()
        for i in range(6):
            cam.render_device(world, d_rgba8=buf.data_ptr(), stats=st)
            ms.append(st.device_ms)
        best = min(ms[1:])
        rec = {"scene": name, "w": w, "h": h, "kernel_ms": ms, "mrays_s": st.total_rays / best / 1e3,
               "rays": [st.primary_rays, st.shadow_rays, st.reflect_rays, st.refract_rays], "info": info}
        out[name] = rec
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
