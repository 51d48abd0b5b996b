#!/usr/bin/env python3
"""Generates tests/golden/frames.npz: small frames of every BASELINE config rendered by the oracle (cached mode, which
tests prove bit-identical to the faithful reference algorithm).  Stored per case: the quantised RGBA8 frame, the f64
colours as raw bits, and the exact ray counts.  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

CASES = [("hexagon", 400, 200), ("table", 240, 135), ("teapot", 96, 54), ("cow", 80, 40), ("cow_teddy", 96, 54),
         ("pumpkin", 96, 54)]


def main():
    orc = helpers.load_oracle()
    out = {}
    for name, w, h in CASES:
        ow, oc = helpers.scenes.build(orc, name, w, h)
        ref, cnt = orc.render(ow, oc, mode=orc.CACHED)
        key = f"{name}@{w}x{h}"
        out[key + "/rgba8"] = orc.quantise_rgba8(ref).reshape(h, w, 4)
        out[key + "/rgb_bits"] = ref.reshape(h, w, 3).view(np.uint64)
        out[key + "/rays"] = np.array([cnt.primary, cnt.shadow, cnt.reflect, cnt.refract], dtype=np.uint64)
        print(key, "rays", out[key + "/rays"].tolist())
    path = os.path.join(ROOT, "tests", "golden", "frames.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
