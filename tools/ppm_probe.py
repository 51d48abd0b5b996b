#!/usr/bin/env python3
"""Times the three to_ppm encoders on a rendered frame: device (rtc_ppm_encode_device), host multi-threaded
(rtc_ppm_from_rgba8 after a device->host copy of the pixels) — and checks they agree."""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

rtc = importlib.import_module("ray-tracer-challenge-rust_b200")
for name, w, h in (("table", 1920, 1080), ("cow_teddy", 3840, 2160), ("pumpkin", 7680, 4320)):
    world, cam = rtc.build_scene(name, w, h)
    buf = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
    cam.render_device(world, d_rgba8=buf.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    import ctypes as C
    api = rtc.api()
    cap = api.ppm_max_bytes(w, h)
    pinned = api.pinned_alloc(cap)  # a caller that encodes many frames allocates its pinned buffer once
    n = C.c_uint64(0)
    api.check(api.ppm_encode_device(0, C.c_void_p(buf.data_ptr()), w, h, None, C.c_void_p(pinned), cap, C.byref(n)))  # warm-up
    t0 = time.perf_counter()
    api.check(api.ppm_encode_device(0, C.c_void_p(buf.data_ptr()), w, h, None, C.c_void_p(pinned), cap, C.byref(n)))
    t1 = time.perf_counter()
    a = C.string_at(pinned, n.value)
    api.pinned_free(pinned)
    t1b = time.perf_counter()
    t1 = t1  # noqa
    tA = time.perf_counter()
    host = buf.cpu().numpy()
    t2 = time.perf_counter()
    b = rtc.ppm_from_rgba8(host, w, h)
    t3 = time.perf_counter()
    print(f"{name} {w}x{h}: {len(a) / 1e6:.1f} MB text; device encode + text D2H into a reused pinned buffer {1e3 * (t1 - t0):.1f} ms; "
          f"pixels D2H {1e3 * (t2 - tA):.1f} ms + host encode ({os.cpu_count()} cores) {1e3 * (t3 - t2):.1f} ms; equal={a == b}", flush=True)
