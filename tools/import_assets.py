#!/usr/bin/env python3
"""Convert the reference's OBJ inputs into compact binary mesh fixtures under assets/.

/root/reference does not exist on the GPU box, so bench.py / tests / smoke() cannot read the OBJ files there at run
time.  This script (run once, in the build container) parses `/root/reference/objs/*.obj` with the reference parser's
rules (obj_file.rs:29-114: `v x y z` as f64, `f a b c ...` fan-triangulated with 1-based integer indices, everything
else ignored; none of the four files has a `g` line) and stores the *data* — vertices as f64, triangles as int32 —
as `assets/<name>.npz`.  Python's float() and Rust's str::parse::<f64>() are both correctly rounded, so the stored
doubles are the ones the reference would hold.
"""
import os
import sys

import numpy as np

SRC = "/root/reference/objs"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
NAMES = {"teapot.obj": "teapot", "cow-nonormals.obj": "cow", "teddy.obj": "teddy", "pumpkin_tall_10k.obj": "pumpkin"}


def parse(path):
    verts, faces, ignored = [], [], 0
    with open(path, "r") as f:
        for line in f.read().split("\n"):
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "f":
                idx = [int(t) for t in tok[1:]]
                for k in range(2, len(idx)):
                    faces.append([idx[0], idx[k - 1], idx[k]])
            elif tok[0] == "g":
                raise SystemExit(f"{path}: unexpected `g` line (named groups are not representable in this fixture)")
            else:
                ignored += 1
    return np.asarray(verts, dtype=np.float64), np.asarray(faces, dtype=np.int32), ignored


def main():
    os.makedirs(DST, exist_ok=True)
    for fn, name in NAMES.items():
        v, f, ign = parse(os.path.join(SRC, fn))
        assert f.min() >= 1 and f.max() <= len(v)
        out = os.path.join(DST, name + ".npz")
        np.savez_compressed(out, vertices=v, faces=f, ignored_lines=np.int64(ign), source=np.bytes_(fn.encode()))
        print(f"{fn}: {len(v)} vertices, {len(f)} triangles, {ign} ignored lines -> {out} ({os.path.getsize(out)} bytes)")


if __name__ == "__main__":
    sys.exit(main())
