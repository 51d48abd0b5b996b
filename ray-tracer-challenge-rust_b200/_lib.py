"""Loads librtc_b200.so (built in-tree by build.py) and declares the rest of include/rtc.h."""
import ctypes as C
import os

from ._capi import BuilderApi, CameraDesc, Computations, Material, Rows, Stats, c_double_p, c_u64_p

HERE = os.path.dirname(os.path.abspath(__file__))
# RTC_B200_LIB selects a tuning variant built by tools/tune_variants.py (same sources, other launch shape)
LIB_PATH = os.environ.get("RTC_B200_LIB") or os.path.join(HERE, "librtc_b200.so")

RTC_BUILD_HOST_SAH, RTC_BUILD_DEVICE_LBVH = 0, 1
RTC_OK, RTC_ERR_INVALID, RTC_ERR_PANIC, RTC_ERR_CUDA, RTC_ERR_UNSUPPORTED, RTC_ERR_TIMEOUT = 0, -1, -2, -3, -4, -6


class RtcError(RuntimeError):
    """A non-zero status from librtc_b200.so.  code == RTC_ERR_PANIC marks what the reference would panic on."""

    def __init__(self, code, message):
        super().__init__(f"rtc error {code}: {message}")
        self.code, self.message = code, message


def _load(path):
    """dlopen with a short retry: on a freshly provisioned box the snapshot copy of a 70 MB library has been seen still
    in flight when the first ranks start ("file too short")."""
    import time
    for attempt in range(10):
        try:
            return C.CDLL(path)
        except OSError as e:
            if "too short" not in str(e) and "truncated" not in str(e) or attempt == 9:
                raise
            time.sleep(1.0)


class RtcApi(BuilderApi):
    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with `python ray-tracer-challenge-rust_b200/build.py` "
                              "(there is no CPU fallback)")
        super().__init__(_load(path), "rtc_")
        self.path = path
        f, vp = self._fn, C.c_void_p
        u8p = C.POINTER(C.c_uint8)
        f("device_count", C.c_int)
        f("enable_peer_access", C.c_int, C.c_int, C.c_int)
        f("frame_share_create", C.c_int, C.c_int, C.c_uint64, C.POINTER(vp), C.c_char_p)
        f("frame_share_open", C.c_int, C.c_int, C.c_char_p, C.POINTER(vp))
        f("frame_share_close", C.c_int, C.c_int, vp, C.c_int)
        f("host_share_create", C.c_int, C.c_int, C.c_char_p, C.c_uint64, C.POINTER(vp))
        f("host_share_open", C.c_int, C.c_int, C.c_char_p, C.c_uint64, C.POINTER(vp))
        f("host_share_close", C.c_int, vp, C.c_uint64, C.c_char_p)
        f("host_counter_store", None, vp, C.c_uint64)
        f("host_counter_load", C.c_uint64, vp)
        f("host_counter_wait", C.c_int, vp, C.c_uint64, C.c_double)
        f("scene_create", C.c_int, vp, C.c_int, C.POINTER(vp))
        f("scene_create_ex", C.c_int, vp, C.c_int, C.c_uint32, C.POINTER(vp))
        f("scene_destroy", None, vp)
        f("scene_info", C.c_int, vp, c_u64_p)
        f("scene_upload_bytes", C.c_uint64, vp)
        f("render", C.c_int, vp, C.POINTER(CameraDesc), C.POINTER(Rows), vp, vp, C.POINTER(Stats))
        f("render_device", C.c_int, vp, C.POINTER(CameraDesc), C.POINTER(Rows), vp, vp, vp, C.c_int, C.POINTER(Stats))
        f("multi_create", C.c_int, vp, C.c_int, C.c_uint32, C.POINTER(vp))
        f("multi_render", C.c_int, vp, C.POINTER(CameraDesc), C.c_uint32, vp, C.POINTER(Stats))
        f("multi_render_host", C.c_int, vp, C.POINTER(CameraDesc), vp, vp, C.POINTER(Stats))
        f("multi_host_frame", vp, vp)
        f("multi_device_frame", vp, vp)
        f("multi_destroy", None, vp)
        f("render_multi", C.c_int, vp, C.POINTER(CameraDesc), C.c_int, C.c_uint32, vp, C.POINTER(Stats))
        f("rows_count", C.c_uint32, C.POINTER(CameraDesc), C.POINTER(Rows))
        f("render_device_notify", C.c_int, vp, C.POINTER(CameraDesc), C.POINTER(Rows), vp, vp, vp, vp)
        f("stream_wait_counter", C.c_int, C.c_int, vp, vp, C.c_uint32)
        f("stream_set_counters", C.c_int, C.c_int, vp, C.POINTER(vp), C.c_uint32, C.c_uint32)
        f("color_at", C.c_int, vp, c_double_p, C.c_uint64, c_double_p)
        f("intersect", C.c_int, vp, c_double_p, C.c_uint64, C.c_uint32, c_double_p, C.POINTER(C.c_int32),
          C.POINTER(C.c_uint32))
        f("prepare_computations", C.c_int, vp, c_double_p, C.c_uint64, C.POINTER(Computations))
        f("normal_at", C.c_int, vp, C.c_int32, c_double_p, C.c_uint64, c_double_p)
        f("measure_fp64_peak", C.c_int, C.c_int, c_double_p, c_double_p)
        f("selftest_shared_divisor", C.c_int, C.c_int, C.c_uint64, C.c_uint64, c_u64_p)
        f("world_color_at", C.c_int, vp, c_double_p, C.c_uint64, c_double_p)
        f("world_scene", C.c_int, vp, C.c_int, C.POINTER(vp))
        f("world_set_build", C.c_int, vp, C.c_uint32)
        f("world_drop_scenes", None, vp)
        f("world_describe", C.c_int, vp, c_u64_p)
        f("world_flatten_info", C.c_int, vp, c_u64_p, c_double_p, C.c_uint64)
        f("world_kernel_features", C.c_int, vp, C.POINTER(C.c_uint32))
        f("world_marshal", C.c_int, vp, C.POINTER(vp))
        f("marshalled_desc", vp, vp)
        f("marshalled_free", None, vp)
        f("camera_desc_get", None, vp, C.POINTER(CameraDesc))
        f("camera_render", C.c_int, vp, vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(Stats))
        f("canvas_new", vp, C.c_uint64, C.c_uint64)
        f("canvas_free", None, vp)
        f("canvas_width", C.c_uint64, vp)
        f("canvas_height", C.c_uint64, vp)
        f("canvas_get_pixel", C.c_int, vp, C.c_uint64, C.c_uint64, c_double_p)
        f("canvas_set_pixel", C.c_int, vp, C.c_uint64, C.c_uint64, c_double_p)
        f("canvas_pixels_f64", c_double_p, vp)
        f("canvas_pixels_rgba8", u8p, vp)
        f("canvas_to_ppm", vp, vp, c_u64_p)
        f("ppm_from_rgba8", vp, vp, C.c_uint64, C.c_uint64, c_u64_p)
        f("ppm_max_bytes", C.c_uint64, C.c_uint64, C.c_uint64)
        f("ppm_encode_device", C.c_int, C.c_int, vp, C.c_uint64, C.c_uint64, vp, vp, C.c_uint64, c_u64_p)
        f("pinned_alloc", vp, C.c_uint64)
        f("pinned_free", None, vp)
        f("free", None, vp)

    def check(self, rc):
        if rc != RTC_OK:
            raise RtcError(rc, self.error())


_api = None


def api():
    """The process-wide binding (loads the library on first use; raises ImportError if it was never built)."""
    global _api
    if _api is None:
        _api = RtcApi()
    return _api
