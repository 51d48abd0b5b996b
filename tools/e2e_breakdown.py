#!/usr/bin/env python3
"""Where the end-to-end frame time goes: marshal / host flatten (+BVH) / rtc_scene_create (flatten + upload) / rtc_render
into a pinned host frame / rtc_scene_destroy, median of N calls each, per workload.
    python tools/e2e_breakdown.py [workload ...] > gpurun_out/e2e_breakdown.json"""
import ctypes as C
import importlib
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rtc = importlib.import_module("ray-tracer-challenge-rust_b200")


def med(fn, n=15):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


def main():
    import torch
    api = rtc.api()
    out = {}
    for name in sys.argv[1:] or ["table", "teapot", "cow_teddy", "pumpkin"]:
        world, cam = rtc.build_scene(name)
        m = C.c_void_p()
        api.check(api.world_marshal(world.h, C.byref(m)))
        desc = api.marshalled_desc(m)
        n8 = (C.c_uint64 * 8)()
        frame = torch.empty((cam.vsize, cam.hsize, 4), dtype=torch.uint8).pin_memory()
        cdesc = cam.desc()
        scene = C.c_void_p()
        api.check(api.scene_create(desc, 0, C.byref(scene)))

        def marshal():
            mm = C.c_void_p()
            api.check(api.world_marshal(world.h, C.byref(mm)))
            api.marshalled_free(mm)

        def create_destroy():
            sc = C.c_void_p()
            api.check(api.scene_create(desc, 0, C.byref(sc)))
            api.scene_destroy(sc)

        def create_destroy_device():
            sc = C.c_void_p()
            api.check(api.scene_create_ex(desc, 0, rtc.RTC_BUILD_DEVICE_LBVH, C.byref(sc)))
            api.scene_destroy(sc)

        dscene = C.c_void_p()
        api.check(api.scene_create_ex(desc, 0, rtc.RTC_BUILD_DEVICE_LBVH, C.byref(dscene)))

        def render_device_built():
            api.check(api.render(dscene, C.byref(cdesc), None, C.c_void_p(frame.data_ptr()), None, None))

        def render():
            api.check(api.render(scene, C.byref(cdesc), None, C.c_void_p(frame.data_ptr()), None, None))

        for _ in range(3):
            create_destroy(); render(); create_destroy_device(); render_device_built()
        out[name] = {"marshal_ms": med(marshal), "flatten_only_ms": med(lambda: api.world_flatten_info(world.h, n8, None, 0)),
                     "scene_create_destroy_ms": med(create_destroy), "render_host_ms": med(render),
                     "scene_create_destroy_device_build_ms": med(create_destroy_device),
                     "render_host_device_built_ms": med(render_device_built),
                     "chunk_mib": os.environ.get("RTC_HOST_CHUNK_MIB", "default")}
        api.scene_destroy(scene)
        api.scene_destroy(dscene)
        api.marshalled_free(m)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
