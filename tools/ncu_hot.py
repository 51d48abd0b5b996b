#!/usr/bin/env python3
"""Per-opcode instruction mix and hottest SASS of an .ncu-rep source page (--import-source on).
    python tools/ncu_hot.py prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, isrc, isamp, iexec, ithr = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed",
                                                            "Thread Instructions Executed"))
    ops = collections.Counter()
    samp = collections.Counter()
    total = tsamp = 0
    body = []
    for r in rows[2:]:
        if len(r) <= ithr or not r[ia].startswith("0x"):
            continue
        n = int(r[iexec] or 0)
        s = int(r[isamp] or 0)
        src = r[isrc].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.split(".")[0]
        ops[op] += n
        samp[op] += s
        total += n
        tsamp += s
        body.append((n, s, r[ia], src, r[ithr]))
    print(f"warp-instructions executed: {total}   pc samples: {tsamp}")
    print("opcode mix (share of executed warp instructions | share of samples):")
    for op, n in ops.most_common(28):
        print(f"  {op:10s} {100 * n / total:6.2f} %   {100 * samp[op] / max(tsamp, 1):6.2f} %")
    print(f"\nhottest {topn} instructions by samples:")
    for n, s, a, src, thr in sorted(body, key=lambda x: -x[1])[:topn]:
        print(f"  {s:6d} samples  {n:10d} exec  avg thr {int(thr) / max(n, 1):5.1f}  {src[:90]}")


if __name__ == "__main__":
    main()
