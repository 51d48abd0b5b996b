#!/usr/bin/env python3
"""Phase times of rtc_scene_create_ex for one workload and both mesh builds (run with RTC_B200_TRACE=1; the library
prints one line per call on stderr).
    RTC_B200_TRACE=1 python tools/create_trace.py pumpkin 2> gpurun_out/create_trace.txt"""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rtc = importlib.import_module("ray-tracer-challenge-rust_b200")


def main():
    api = rtc.api()
    for name in sys.argv[1:] or ["teapot", "pumpkin"]:
        world, _ = rtc.build_scene(name)
        m = C.c_void_p()
        api.check(api.world_marshal(world.h, C.byref(m)))
        desc = api.marshalled_desc(m)
        for flags, label in ((rtc.RTC_BUILD_HOST_SAH, "host SAH"), (rtc.RTC_BUILD_DEVICE_LBVH, "device LBVH")):
            sys.stderr.write(f"--- {name}, {label}\n")
            sys.stderr.flush()
            for _ in range(6):
                sc = C.c_void_p()
                api.check(api.scene_create_ex(desc, 0, flags, C.byref(sc)))
                api.scene_destroy(sc)
        api.marshalled_free(m)


if __name__ == "__main__":
    main()
