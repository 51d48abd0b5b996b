#!/usr/bin/env python3
"""Record dram__bytes_read.sum + dram__bytes_write.sum per launch of render_kernel from an ncu --set full report into
profiles/ncu_traffic.json (bench.py reports it as roofline.traffic).
    python tools/ncu_traffic.py "table 1920x1080" gpurun_out/prof.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    key, rep = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    vals = []
    for r in rows[2:]:
        d = {h: (units[i], r[i]) for i, h in enumerate(hdr)}
        if "render_kernel" not in d["Kernel Name"][1]:
            continue
        b = sum(float(d[k][1]) * UNIT[d[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        vals.append((b, float(d["gpu__time_duration.sum"][1]), d["gpu__time_duration.sum"][0]))
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tj = json.load(open(path)) if os.path.exists(path) else {}
    tj[key] = {"dram_bytes_per_launch": sum(v[0] for v in vals) / len(vals), "launches_profiled": len(vals),
               "report": os.path.basename(rep), "kernel_time": [f"{v[1]} {v[2]}" for v in vals]}
    json.dump(tj, open(path, "w"), indent=1)
    print(key, tj[key])


if __name__ == "__main__":
    main()
