#!/usr/bin/env python3
"""DRAM traffic of the render kernel per library variant (built by tools/tune_variants.py build ...), one scene:

    gpurun -- python tools/dram_variants.py pumpkin 7680 4320        # writes gpurun_out/dram_variants.json

Every variant renders the scene a few times under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum` (the launches after the first; kernel times under ncu are NOT bench
numbers — the times in the output come from a separate plain run with CUDA events)."""
import csv
import importlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "ray-tracer-challenge-rust_b200", "variants")
METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum"


def one(name, w, h):
    import torch
    rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
    world, cam = rtc.build_scene(name, w, h)
    buf = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    st = rtc.Stats()
    ms = []
    for _ in range(4):
        flush.zero_()
        cam.render_device(world, d_rgba8=buf.data_ptr(), stats=st)
        ms.append(st.device_ms)
    print(json.dumps({"ms": ms[1:]}))


def main():
    if sys.argv[1] == "one":
        return one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    scene, w, h = sys.argv[1], sys.argv[2], sys.argv[3]
    variants = json.load(open(os.path.join(VDIR, "variants.json")))
    out = {"scene": f"{scene} {w}x{h}", "defines": variants, "results": {}}
    for name in variants:
        env = dict(os.environ, RTC_B200_LIB=os.path.join(VDIR, f"librtc_{name}.so"))
        cmd = [sys.executable, os.path.abspath(__file__), "one", scene, w, h]
        plain = subprocess.run(cmd, env=env, capture_output=True, text=True)
        try:
            ms = json.loads(plain.stdout.strip().split("\n")[-1])["ms"]
        except Exception:
            out["results"][name] = {"error": (plain.stdout + plain.stderr)[-400:]}
            continue
        q = subprocess.run(["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", "regex:render_kernel", "--csv"] + cmd,
                           env=env, capture_output=True, text=True)
        rows = [r for r in csv.reader(io.StringIO(q.stdout)) if len(r) > 10 and r[0].isdigit()]
        per = {}
        for r in rows:
            per.setdefault(r[0], {})[r[-3]] = (r[-2], float(r[-1].replace(",", "")))
        launches = [per[k] for k in sorted(per, key=int)][1:]  # skip the first (cold) launch
        def avg(metric):
            vals = []
            for l in launches:
                unit, v = l.get(metric, ("", float("nan")))
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "sector": 1, "": 1}.get(unit, 1)
                vals.append(v * scale)
            return sum(vals) / len(vals) if vals else None
        out["results"][name] = {"kernel_ms_plain": ms, "dram_write_bytes": avg("dram__bytes_write.sum"),
                                "dram_read_bytes": avg("dram__bytes_read.sum"),
                                "l2_write_miss_sectors": avg("lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum"),
                                "launches_profiled": len(launches)}
        print(name, out["results"][name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dram_variants.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
