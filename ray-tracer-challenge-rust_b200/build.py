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

MANIFEST = os.path.join(CSRC, "manifest.txt")


def read_manifest(path=MANIFEST):
    """csrc/manifest.txt -> {"hostflags": [...], "nvccflags": [...], "link": [...], "units": [(kind, file, define)]}: the ONE
    list of translation units, shared with rtc-sys/build.rs.  An `instances` line expands to one unit per feature mask."""
    import re
    m = {"hostflags": [], "nvccflags": [], "link": [], "units": []}
    for line in open(path):
        words = line.split()
        if not words or words[0].startswith("#"):
            continue
        key = words[0]
        if key in ("hostflags", "nvccflags"):
            m[key] = words[1:]
        elif key == "link":
            m["link"] += words[1:]
        elif key in ("host", "cuda"):
            m["units"].append((key, words[1], None))
        elif key == "instances":
            file, macro, where = words[1:4]
            header, listmacro = where.split(":")
            text = open(os.path.join(os.path.dirname(path), header)).read()
            entries = re.search(r"#define " + re.escape(listmacro) + r"\(X\)(.*)", text).group(1)
            for mask in re.findall(r"X\((\d+)\)", entries):
                m["units"].append(("cuda", file, f"{macro}={mask}"))
        else:
            raise RuntimeError(f"{path}: unknown directive {key}")
    return m


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _sources():
    out = []
    for d, dirs, files in os.walk(CSRC):
        dirs[:] = [x for x in dirs if not x.startswith(".") and x != "__pycache__"]  # tool caches are not sources
        out += [os.path.join(d, f) for f in files if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", ".txt"))]
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
    man = read_manifest()
    with open(os.path.join(bdir, "build.log"), "w") as log:
        # every translation unit of csrc/manifest.txt, in parallel: the host TU with g++, the CUDA TUs with nvcc
        # (render_inst.cu once per feature mask)
        jobs = []
        for kind, file, define in man["units"]:
            obj = file.replace(".", "_") + ("_" + define.split("=")[1] if define else "") + ".o"
            flags = list(dflags) + (["-D" + define] if define else [])
            jobs.append((obj, kind, file, flags))
        from concurrent.futures import ThreadPoolExecutor
        import io

        def compile_one(job):
            obj, kind, file, flags = job
            buf = io.StringIO()
            compiler = [nvcc, *man["nvccflags"]] if kind == "cuda" else ["g++", *man["hostflags"]]
            _run([*compiler, *flags, "-c", os.path.join(CSRC, file), "-o", os.path.join(bdir, obj)], buf)
            return buf.getvalue()

        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            for text in pool.map(compile_one, jobs):
                log.write(text)
        _run([nvcc, "-shared", *man["link"], "-o", lib, *[os.path.join(bdir, j[0]) for j in jobs]], log)
    if not out:
        with open(STAMP, "w") as f:
            f.write(_digest() + "\n")
    if verbose:
        print(open(os.path.join(bdir, "build.log")).read())
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
