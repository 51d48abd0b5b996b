#!/usr/bin/env python3
"""How a host-output render with the f64 Canvas is cut into launches (render_host: each chunk's pixels cross PCIe while the
next chunk renders): the drop-in call's wall time per chunk schedule, RTC_B200_F64_CHUNKS = weights of the chunks.

    gpurun -- python tools/chunk_sweep.py            # writes gpurun_out/chunk_sweep.json

Per schedule a fresh process (the weights are read once); per scene `steps` x (rtc_world_drop_scenes -> rtc_camera_render
(want_f64 = 1) -> rtc_canvas_free), mean and best wall time; every schedule's canvas is compared with the first one's."""
import hashlib
import importlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SCENES = [("table", 1920, 1080, 40), ("teapot", 1920, 1080, 40), ("cow_teddy", 3840, 2160, 20), ("pumpkin", 7680, 4320, 6)]
SCHEDULES = ["1,1,2,4", "1,1,1,1", "1,2,3,2", "1,2,2,2,1", "1,1,1,1,1,1", "1,2,2,2,2,2,2,1", "1,1,1,1,1,1,1,1", "1,3,4", "1,1"]


def one():
    rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
    out = {}
    for name, w, h, steps in SCENES:
        world, cam = rtc.build_scene(name, w, h)
        world.set_build("device")
        canvas = None
        times = []
        for i in range(steps + 3):
            t0 = time.perf_counter()
            world.drop_scenes()
            canvas = None
            canvas = cam.render(world, want_f64=True)
            times.append(time.perf_counter() - t0)
        times = times[3:]
        digest = hashlib.sha1(canvas.pixels_f64().tobytes()).hexdigest()[:16]
        out[name] = {"mean_ms": sum(times) / len(times) * 1e3, "best_ms": min(times) * 1e3, "canvas": digest}
        canvas = None
    print(json.dumps(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        return one()
    res = {}
    for sched in SCHEDULES:
        env = dict(os.environ, RTC_B200_F64_CHUNKS=sched)
        q = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True)
        try:
            res[sched] = json.loads(q.stdout.strip().split("\n")[-1])
        except Exception:
            res[sched] = {"error": (q.stdout + q.stderr)[-400:]}
        print(sched, json.dumps(res[sched]), flush=True)
    first = res[SCHEDULES[0]]
    same = all(res[s].get(n, {}).get("canvas") == first[n]["canvas"] for s in SCHEDULES for n in first if "error" not in res[s])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"schedules": res, "all_canvases_identical": same}, open(os.path.join(ROOT, "gpurun_out", "chunk_sweep.json"), "w"),
              indent=1)
    print("all canvases identical:", same)


if __name__ == "__main__":
    main()
