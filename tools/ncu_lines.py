#!/usr/bin/env python3
"""Executed warp instructions and pc samples per SOURCE LINE of an .ncu-rep captured with --import-source on (-lineinfo):
    python tools/ncu_lines.py prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = "?", None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            i_line, i_src, i_addr = 0, 1, 2
            i_samp, i_exec = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) <= i_exec or not r[i_line]:
            continue  # SASS rows have an empty line number
        try:
            lines.append((int(r[i_exec] or 0), int(r[i_samp] or 0), cur_file, int(r[i_line]), r[i_src].strip()))
        except ValueError:
            pass
    tot_e, tot_s = sum(l[0] for l in lines) or 1, sum(l[1] for l in lines) or 1
    print(f"warp instructions {tot_e}   samples {tot_s}")
    for e, s, f, ln, src in sorted(lines, key=lambda x: -x[0])[:topn]:
        print(f"{100 * e / tot_e:6.2f} % exec {100 * s / tot_s:6.2f} % samp  {f}:{ln:<5d} {src[:100]}")


if __name__ == "__main__":
    main()
