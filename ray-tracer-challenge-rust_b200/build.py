"""Builds librtc_b200.so (the C-ABI library of include/rtc.h) in-tree for sm_100a.

    python ray-tracer-challenge-rust_b200/build.py [--force]

Host translation units are compiled by g++ with -ffp-contract=off (their matrices, gate boxes and triangle normals
reach pixels and must round exactly as the reference's Rust does); the CUDA translation unit by nvcc for
compute_100a/sm_100a with -fmad=false (same reason, see csrc/rt_core.cuh) and -lineinfo for ncu source pages.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "librtc_b200.so")
ROOT = os.path.dirname(HERE)

HOST_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fno-unsafe-math-optimizations", "-fPIC",
              "-Wall", "-Wextra", "-Wno-unused-parameter"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-Xptxas", "-v"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _instance_masks():
    """The feature masks render_kernel is instantiated for (RTC_RENDER_INSTANCES in csrc/render_launch.cuh)."""
    import re
    text = open(os.path.join(CSRC, "render_launch.cuh")).read()
    line = re.search(r"#define RTC_RENDER_INSTANCES\(X\)(.*)", text).group(1)
    return [int(m) for m in re.findall(r"X\((\d+)\)", line)]


def _sources():
    out = []
    for d, _, files in os.walk(CSRC):
        out += [os.path.join(d, f) for f in files]
    out += [os.path.join(ROOT, "include", "rtc.h"), os.path.abspath(__file__)]
    return out


STAMP = LIB + ".sources.sha256"


def _digest():
    """Content hash of everything the library is built from (file times do not survive the copy to a GPU box)."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(_sources()):
        h.update(os.path.relpath(path, ROOT).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != _digest()


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.write("$ " + " ".join(cmd) + "\n" + p.stdout + "\n")
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))


def build(force=False, verbose=False, defines=(), out=None):
    """Compile if sources are newer than the library; returns the library path.  `defines` / `out` build a tuning
    variant (extra -D macros for render.cu, written to another path) without touching the default library."""
    lib = out or LIB
    if not force and not out and not _stale():
        return LIB
    bdir = BUILD if not out else os.path.join(BUILD, os.path.basename(out) + ".d")
    os.makedirs(bdir, exist_ok=True)
    nvcc = _nvcc()
    dflags = ["-D" + d for d in defines]
    with open(os.path.join(bdir, "build.log"), "w") as log:
        _run(["g++", *HOST_FLAGS, *dflags, "-c", os.path.join(CSRC, "capi.cpp"), "-o", os.path.join(bdir, "capi.o")], log)
        # the CUDA translation units in parallel: management + small kernels, the tally build, and render_kernel once
        # per feature mask (csrc/render_launch.cuh RTC_RENDER_INSTANCES)
        jobs = [("render.o", ["render.cu"], dflags), ("render_tally.o", ["render_tally.cu"], []),
                ("ppm_encode.o", ["ppm_encode.cu"], []), ("lbvh.o", ["lbvh.cu"], []), ("probe.o", ["probe.cu"], [])]
        for mask in _instance_masks():
            jobs.append((f"render_inst_{mask}.o", ["render_inst.cu"], dflags + [f"-DRTC_INST_MASK={mask}"]))
        from concurrent.futures import ThreadPoolExecutor
        import io

        def compile_one(job):
            obj, srcs, flags = job
            buf = io.StringIO()
            _run([nvcc, *NVCC_FLAGS, *flags, "-c", *[os.path.join(CSRC, s) for s in srcs], "-o",
                  os.path.join(bdir, obj)], buf)
            return buf.getvalue()

        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            for text in pool.map(compile_one, jobs):
                log.write(text)
        _run([nvcc, "-shared", "-o", lib, os.path.join(bdir, "capi.o"), *[os.path.join(bdir, j[0]) for j in jobs]], log)
    if not out:
        with open(STAMP, "w") as f:
            f.write(_digest() + "\n")
    if verbose:
        print(open(os.path.join(bdir, "build.log")).read())
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
