// capi.cpp — the extern "C" surface of librtc_b200.so (include/rtc.h): the CORE boundary that replaces
// Camera::render (camera.rs:67-79) and World::color_at (world.rs:80-82), and the HOST MIRROR of the reference's scene
// API.  There is no CPU rendering path in this library: every render / color_at call needs a CUDA device and fails with
// RTC_ERR_CUDA otherwise.
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>

#include "../../include/rtc.h"
#include "flatten.hpp"
#include "host_model.hpp"
#include "render.cuh"

using namespace rtc;

namespace {
thread_local std::string g_err;
int set_err(int code, const std::string& m) {
    g_err = m;
    return code;
}
DCamera to_dcamera(const rtc_camera_desc& c) {
    DCamera d;
    d.hsize = c.hsize;
    d.vsize = c.vsize;
    std::memcpy(d.inv, c.inverse, sizeof(d.inv));
    d.half_width = c.half_width;
    d.half_height = c.half_height;
    d.pixel_size = c.pixel_size;
    return d;
}
// rows == NULL: the whole frame
int to_drows(const rtc_camera_desc& cam, const rtc_rows* rows, DRows* out) {
    if (!rows) {
        out->band_rows = cam.vsize ? cam.vsize : 1;
        out->band_first = 0;
        out->band_stride = 1;
        out->local_rows = cam.vsize;
        out->frame_layout = 0;
        out->row_begin = 0;
        out->row_count = cam.vsize;
        out->pad = 0;
        out->notify = 0;
        return RTC_OK;
    }
    if (rows->layout > RTC_ROWS_FRAME) return set_err(RTC_ERR_INVALID, "unknown rtc_rows.layout");
    out->frame_layout = rows->layout;
    if (rows->band_rows == 0 || rows->band_stride == 0) return set_err(RTC_ERR_INVALID, "band_rows and band_stride must be > 0");
    out->band_rows = rows->band_rows;
    out->band_first = rows->band_first;
    out->band_stride = rows->band_stride;
    // rows owned by bands band_first, band_first + stride, ... below vsize
    uint64_t local = 0;
    for (uint64_t b = rows->band_first;; b += rows->band_stride) {
        uint64_t r0 = b * rows->band_rows;
        if (r0 >= cam.vsize) break;
        uint64_t r1 = r0 + rows->band_rows;
        if (r1 > cam.vsize) r1 = cam.vsize;
        local += r1 - r0;
    }
    out->local_rows = (uint32_t)local;
    out->row_begin = 0;
    out->row_count = (uint32_t)local;
    out->pad = 0;
    out->notify = 0;
    return RTC_OK;
}
void fill_stats(const LaunchStats& ls, rtc_stats* st) {
    if (!st) return;
    st->primary_rays = ls.primary;
    st->shadow_rays = ls.shadow;
    st->reflect_rays = ls.reflect;
    st->refract_rays = ls.refract;
    st->kernel_launches = ls.launches;
    st->device_ms = ls.device_ms;
}
bool affine_camera(const rtc_camera_desc& c) {
    const double* b = c.inverse;
    return b[12] == 0.0 && b[13] == 0.0 && b[14] == 0.0 && b[15] == 1.0;
}
}  // namespace

struct rtc_scene {
    DeviceScene* dev = nullptr;
    uint64_t info[6] = {0, 0, 0, 0, 0, 0};
};
struct rtc_multi {
    MultiRenderer* m = nullptr;
};
struct rtc_shape {
    std::unique_ptr<HShape> s;
};
struct rtc_world {
    HWorld w;
    std::map<int, rtc_scene*> scenes;  // per device: marshalled + uploaded on first use; dropped when the world changes
    uint32_t build_flags = RTC_BUILD_HOST_SAH;
    uint32_t recursion_limit = 0;      // 0: the reference's RECURSION_LIMIT = 5 (world.rs:11)
    std::mutex mu;                     // held for the whole of every world-level call (render, color_at, push, ...)
};
struct rtc_camera {
    HCamera c;
};
struct rtc_marshalled {
    Marshalled m;
};
struct rtc_canvas {
    uint64_t width = 0, height = 0;
    double* rgb = nullptr;     // Canvas.pixels (canvas.rs:8), 3 f64 per pixel; may be absent for want_f64 = 0 renders
    // the same pixels quantised as canvas.rs:61-63.  A canvas rendered WITH its f64 colours holds only those at first — as the
    // reference's Canvas does, which quantises at PPM time — and this array is filled from them the first time it is asked
    // for (canvas_rgba8): a frame's 4 B/px then never cross PCIe next to its 24 B/px
    mutable uint8_t* rgba8 = nullptr;
    bool rgb_pinned = false;
    mutable bool rgba_pinned = false;
    mutable std::mutex lazy;
};

// RTC_B200_TRACE=1: where the host time of a scene build went (stderr, one line per call)
static void trace_phases(const char* who, const FlatScene& flat) {
    static const bool trace = std::getenv("RTC_B200_TRACE") != nullptr;
    if (!trace) return;
    const double* t = flat.phase_ms;
    std::fprintf(stderr, "[rtc] %s: validate %.3f  group bounds %.3f  bvh items %.3f / build %.3f / splice %.3f  "
                         "triangle tables %.3f  upload %.3f ms (%zu triangles, %zu bvh nodes)\n",
                 who, t[0], t[1], t[2], t[3], t[4], t[5], t[6], flat.tris.size() + flat.device_tris,
                 flat.bvh.size() + flat.device_nodes);
}

// Flattens `desc` (validation, gate boxes, tables, host BVHs) and hands the result to `upload`; when the device mesh
// build asks for it (a tree deeper than the traversal stack, a coordinate the reference panics on) the scene is flattened
// again with the host build.  upload returns 0, kDeviceBuildTooDeep or another non-zero code with *e set.
template <class Upload>
static int flatten_and_upload(const rtc_scene_desc* desc, uint32_t flags, const char* who, Upload upload) {
    for (int attempt = 0; attempt < 2; attempt++) {
        FlatScene flat;
        std::string e;
        FlattenOptions opts;
        opts.device_mesh_build = (flags & RTC_BUILD_DEVICE_LBVH) && attempt == 0;
        static const bool no_clusters = std::getenv("RTC_B200_NO_CLUSTERS") != nullptr;  // A/B switch (profiles/r02)
        opts.clusters = opts.clusters && !no_clusters;
        // the device build's input arrays are megabytes: keep their pages across calls on this thread instead of
        // faulting fresh ones in every time (a third of the gather time of a 10 k-triangle mesh)
        thread_local std::vector<rtc_triangle_desc> keep_tri;
        thread_local std::vector<int32_t> keep_mat;
        struct Lend {
            FlatScene& f;
            Lend(FlatScene& fs) : f(fs) {
                f.pending_tri.swap(keep_tri);
                f.pending_material.swap(keep_mat);
                f.pending_tri.clear();
                f.pending_material.clear();
            }
            ~Lend() {
                f.pending_tri.swap(keep_tri);
                f.pending_material.swap(keep_mat);
            }
        } lend(flat);
        int rc = flatten_scene(*desc, flat, &e, opts);
        if (rc != RTC_OK) return set_err(rc, e);
        PhaseClock clock;
        rc = upload(flat, &e);
        if (rc == kDeviceBuildTooDeep && attempt == 0) continue;  // rebuild this scene's meshes on the host
        if (rc != 0) return set_err(RTC_ERR_CUDA, e);
        clock.lap(flat.phase_ms, FlatScene::T_UPLOAD);
        trace_phases(who, flat);
        return RTC_OK;
    }
    return set_err(RTC_ERR_CUDA, "scene build failed");
}

extern "C" {

const char* rtc_last_error(void) { return g_err.c_str(); }
int rtc_enable_peer_access(int device, int peer) {
    std::string e;
    if (enable_peer_access(device, peer, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}
int rtc_frame_share_create(int device, uint64_t bytes, void** d_ptr, uint8_t handle64[64]) {
    if (!d_ptr || !handle64) return set_err(RTC_ERR_INVALID, "null argument");
    std::string e;
    if (frame_share_create(device, bytes, d_ptr, handle64, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}
int rtc_frame_share_open(int device, const uint8_t handle64[64], void** d_ptr) {
    if (!d_ptr || !handle64) return set_err(RTC_ERR_INVALID, "null argument");
    std::string e;
    if (frame_share_open(device, handle64, d_ptr, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}
int rtc_frame_share_close(int device, void* d_ptr, int owner) {
    std::string e;
    if (frame_share_close(device, d_ptr, owner, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}
/* Host memory shared by the processes of a sharded render: a POSIX shared-memory segment, mapped and page-locked in every
 * process that opens it.  The creator reserves the pages up front (posix_fallocate), so a /dev/shm that is too small is an
 * error here and not a SIGBUS in the middle of a frame. */
static std::mutex g_share_mu;
static std::map<void*, bool> g_share_locked;  // live mappings -> page-locked?
static int host_share_map(int device, const char* name, uint64_t bytes, bool create, void** out) {
    if (!name || name[0] != '/' || !out || bytes == 0) return set_err(RTC_ERR_INVALID, "host share: name must start with '/', bytes > 0");
    *out = nullptr;
    const int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return set_err(RTC_ERR_INVALID, std::string("shm_open ") + name + ": " + std::strerror(errno));
    const auto fail = [&](const std::string& what, int code) {
        close(fd);
        if (create) shm_unlink(name);
        return set_err(code, what);
    };
    if (create) {
        const int rc = posix_fallocate(fd, 0, (off_t)bytes);
        if (rc != 0) return fail(std::string("posix_fallocate ") + name + ": " + std::strerror(rc), RTC_ERR_INVALID);
    } else {
        struct stat st;
        if (fstat(fd, &st) != 0 || (uint64_t)st.st_size < bytes) return fail(std::string(name) + " is smaller than the requested size", RTC_ERR_INVALID);
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) return fail(std::string("mmap ") + name + ": " + std::strerror(errno), RTC_ERR_INVALID);
    std::string e;
    if (device >= 0 && host_register(device, p, bytes, &e) != 0) {
        munmap(p, bytes);
        return fail(e, RTC_ERR_CUDA);
    }
    close(fd);
    {
        std::lock_guard<std::mutex> lk(g_share_mu);
        g_share_locked[p] = device >= 0;
    }
    *out = p;
    return RTC_OK;
}
int rtc_host_share_create(int device, const char* name, uint64_t bytes, void** out) {
    return host_share_map(device, name, bytes, true, out);
}
int rtc_host_share_open(int device, const char* name, uint64_t bytes, void** out) {
    return host_share_map(device, name, bytes, false, out);
}
int rtc_host_share_close(void* p, uint64_t bytes, const char* unlink_name) {
    int rc = RTC_OK;
    if (p) {
        bool locked = false;
        {
            std::lock_guard<std::mutex> lk(g_share_mu);
            auto it = g_share_locked.find(p);
            if (it == g_share_locked.end()) return set_err(RTC_ERR_INVALID, "not a mapping of rtc_host_share_create / _open");
            locked = it->second;
            g_share_locked.erase(it);
        }
        std::string e;
        if (locked && host_unregister(p, &e) != 0) rc = set_err(RTC_ERR_CUDA, e);
        munmap(p, bytes);
    }
    if (unlink_name) shm_unlink(unlink_name);
    return rc;
}
void rtc_host_counter_store(void* counter, uint64_t value) { __atomic_store_n((uint64_t*)counter, value, __ATOMIC_RELEASE); }
uint64_t rtc_host_counter_load(const void* counter) { return __atomic_load_n((const uint64_t*)counter, __ATOMIC_ACQUIRE); }
int rtc_host_counter_wait(const void* counter, uint64_t at_least, double timeout_s) {
    if (!counter) return set_err(RTC_ERR_INVALID, "null argument");
    const auto t0 = std::chrono::steady_clock::now();
    for (uint64_t spin = 0;; spin++) {
        if (__atomic_load_n((const uint64_t*)counter, __ATOMIC_ACQUIRE) >= at_least) return RTC_OK;
        if ((spin & 1023) == 1023) {
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
                return set_err(RTC_ERR_TIMEOUT, "host counter did not reach " + std::to_string(at_least) + " (a rank of the sharded render is missing)");
            if (spin > (1u << 20)) sched_yield();
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}
int rtc_device_count(void) {
    std::string e;
    int n = cuda_device_count(&e);
    if (n == 0 && !e.empty()) g_err = e;
    return n;
}

/* ---------------------------------------------------------------------------------------------- 1. CORE BOUNDARY */
static rtc_scene* scene_handle(const FlatScene& flat, DeviceScene* dev) {
    rtc_scene* s = new rtc_scene();
    s->dev = dev;
    s->info[0] = flat.leaf_count;
    s->info[1] = flat.gates.size();
    s->info[2] = flat.meshes.size();
    s->info[3] = flat.tris.size() + flat.device_tris;
    s->info[4] = flat.bvh.size() + flat.device_nodes;
    s->info[5] = device_scene_bytes(dev);
    return s;
}
int rtc_scene_create(const rtc_scene_desc* desc, int device, rtc_scene** out) {
    return rtc_scene_create_ex(desc, device, RTC_BUILD_HOST_SAH, out);
}
int rtc_scene_create_ex(const rtc_scene_desc* desc, int device, uint32_t flags, rtc_scene** out) {
    if (!desc || !out) return set_err(RTC_ERR_INVALID, "null argument");
    if (flags & ~(uint32_t)RTC_BUILD_DEVICE_LBVH) return set_err(RTC_ERR_INVALID, "unknown build flag");
    *out = nullptr;
    return flatten_and_upload(desc, flags, "rtc_scene_create", [&](const FlatScene& flat, std::string* e) {
        DeviceScene* dev = nullptr;
        const int rc = device_scene_create(flat, device, &dev, e, nullptr);
        if (rc != 0) return rc;
        *out = scene_handle(flat, dev);
        return 0;
    });
}
void rtc_scene_destroy(rtc_scene* scene) {
    if (!scene) return;
    device_scene_destroy(scene->dev);
    delete scene;
}
uint64_t rtc_scene_upload_bytes(const rtc_scene* scene) { return scene ? device_scene_upload_bytes(scene->dev) : 0; }
int rtc_scene_info(const rtc_scene* scene, uint64_t n[6]) {
    if (!scene || !n) return set_err(RTC_ERR_INVALID, "null argument");
    std::memcpy(n, scene->info, sizeof(scene->info));
    return RTC_OK;
}

int rtc_render(const rtc_scene* scene, const rtc_camera_desc* camera, const rtc_rows* rows, uint8_t* rgba8_out,
               double* rgb_f64_out, rtc_stats* stats) {
    if (!scene || !camera) return set_err(RTC_ERR_INVALID, "null argument");
    if (!affine_camera(*camera)) return set_err(RTC_ERR_UNSUPPORTED, "non-affine camera transform");
    DRows dr;
    int rc = to_drows(*camera, rows, &dr);
    if (rc != RTC_OK) return rc;
    LaunchStats ls;
    std::string e;
    rc = render_host(scene->dev, to_dcamera(*camera), dr, rgba8_out, rgb_f64_out, stats ? &ls : nullptr, &e);
    if (rc != 0) return set_err(RTC_ERR_CUDA, e);
    fill_stats(ls, stats);
    return RTC_OK;
}

int rtc_render_device(const rtc_scene* scene, const rtc_camera_desc* camera, const rtc_rows* rows, void* d_rgba8_out,
                      void* d_rgb_f64_out, void* cuda_stream, int sync_stats, rtc_stats* stats) {
    if (!scene || !camera) return set_err(RTC_ERR_INVALID, "null argument");
    if (!affine_camera(*camera)) return set_err(RTC_ERR_UNSUPPORTED, "non-affine camera transform");
    DRows dr;
    int rc = to_drows(*camera, rows, &dr);
    if (rc != RTC_OK) return rc;
    LaunchStats ls;
    std::string e;
    const bool want = sync_stats && stats;
    rc = render_device(scene->dev, to_dcamera(*camera), dr, d_rgba8_out, d_rgb_f64_out, cuda_stream, want ? &ls : nullptr, &e);
    if (rc != 0) return set_err(RTC_ERR_CUDA, e);
    if (want) fill_stats(ls, stats);
    return RTC_OK;
}

int rtc_render_device_notify(const rtc_scene* scene, const rtc_camera_desc* camera, const rtc_rows* rows, void* d_rgba8_out,
                             void* d_rgb_f64_out, void* cuda_stream, void* d_counter) {
    if (!scene || !camera) return set_err(RTC_ERR_INVALID, "null argument");
    if (!affine_camera(*camera)) return set_err(RTC_ERR_UNSUPPORTED, "non-affine camera transform");
    DRows dr;
    int rc = to_drows(*camera, rows, &dr);
    if (rc != RTC_OK) return rc;
    dr.notify = (unsigned long long)(uintptr_t)d_counter;
    std::string e;
    if (dr.row_count == 0 || camera->hsize == 0) {  // nothing to launch: the counter still has to move
        void* one[1] = {d_counter};
        if (d_counter && stream_counters_add(device_scene_device(scene->dev), cuda_stream, one, 1, &e) != 0)
            return set_err(RTC_ERR_CUDA, e);
        return RTC_OK;
    }
    rc = render_device(scene->dev, to_dcamera(*camera), dr, d_rgba8_out, d_rgb_f64_out, cuda_stream, nullptr, &e);
    if (rc != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}
int rtc_stream_wait_counter(int device, void* cuda_stream, void* d_counter, uint32_t at_least) {
    if (!d_counter) return set_err(RTC_ERR_INVALID, "null argument");
    std::string e;
    if (stream_counter_wait(device, cuda_stream, d_counter, at_least, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}
int rtc_stream_set_counters(int device, void* cuda_stream, void** d_counters, uint32_t n, uint32_t value) {
    if (n && !d_counters) return set_err(RTC_ERR_INVALID, "null argument");
    if (n > 16) return set_err(RTC_ERR_INVALID, "at most 16 counters per call");
    std::string e;
    if (stream_counters_set(device, cuda_stream, d_counters, n, value, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}

int rtc_multi_create(const rtc_scene_desc* desc, int ngpus, uint32_t flags, rtc_multi** out) {
    if (!desc || !out) return set_err(RTC_ERR_INVALID, "null argument");
    if (flags & ~(uint32_t)RTC_BUILD_DEVICE_LBVH) return set_err(RTC_ERR_INVALID, "unknown build flag");
    *out = nullptr;
    std::string ce;
    const int have = cuda_device_count(&ce);
    if (have < 1) return set_err(RTC_ERR_CUDA, ce.empty() ? "no CUDA device" : ce);
    if (ngpus < 1 || ngpus > have) return set_err(RTC_ERR_INVALID, "ngpus must be between 1 and rtc_device_count()");
    return flatten_and_upload(desc, flags, "rtc_multi_create", [&](const FlatScene& flat, std::string* e) {
        MultiRenderer* m = nullptr;
        const int rc = multi_create(flat, ngpus, &m, e);
        if (rc != 0) return rc;
        *out = new rtc_multi{m};
        return 0;
    });
}
void rtc_multi_destroy(rtc_multi* m) {
    if (!m) return;
    multi_destroy(m->m);
    delete m;
}
int rtc_multi_render(rtc_multi* m, const rtc_camera_desc* camera, uint32_t where, uint8_t* rgba8_out, rtc_stats* stats) {
    if (!m || !camera) return set_err(RTC_ERR_INVALID, "null argument");
    if (where > RTC_MULTI_DEVICE_FRAME) return set_err(RTC_ERR_INVALID, "unknown rtc_multi_render target");
    if (where == RTC_MULTI_DEVICE_FRAME && rgba8_out) return set_err(RTC_ERR_INVALID, "rgba8_out needs RTC_MULTI_HOST_FRAME");
    if (!affine_camera(*camera)) return set_err(RTC_ERR_UNSUPPORTED, "non-affine camera transform");
    LaunchStats ls;
    std::string e;
    double frame_ms = 0.;
    if (multi_render(m->m, to_dcamera(*camera), where == RTC_MULTI_DEVICE_FRAME, stats ? &ls : nullptr, &frame_ms, &e) != 0)
        return set_err(RTC_ERR_CUDA, e);
    if (rgba8_out) std::memcpy(rgba8_out, multi_host_frame(m->m), (size_t)camera->hsize * camera->vsize * 4);
    fill_stats(ls, stats);
    return RTC_OK;
}
int rtc_multi_render_host(rtc_multi* m, const rtc_camera_desc* camera, uint8_t* rgba8_out, double* rgb_f64_out,
                          rtc_stats* stats) {
    if (!m || !camera) return set_err(RTC_ERR_INVALID, "null argument");
    if (!affine_camera(*camera)) return set_err(RTC_ERR_UNSUPPORTED, "non-affine camera transform");
    LaunchStats ls;
    std::string e;
    if (multi_render_host(m->m, to_dcamera(*camera), rgba8_out, rgb_f64_out, stats ? &ls : nullptr, &e) != 0)
        return set_err(RTC_ERR_CUDA, e);
    fill_stats(ls, stats);
    return RTC_OK;
}
const uint8_t* rtc_multi_host_frame(const rtc_multi* m) { return m ? (const uint8_t*)multi_host_frame(m->m) : nullptr; }
void* rtc_multi_device_frame(const rtc_multi* m) { return m ? multi_device_frame(m->m) : nullptr; }
int rtc_render_multi(const rtc_scene_desc* desc, const rtc_camera_desc* camera, int ngpus, uint32_t flags,
                     uint8_t* rgba8_out, rtc_stats* stats) {
    if (!rgba8_out) return set_err(RTC_ERR_INVALID, "null argument");
    rtc_multi* m = nullptr;
    int rc = rtc_multi_create(desc, ngpus, flags, &m);
    if (rc != RTC_OK) return rc;
    rc = rtc_multi_render(m, camera, RTC_MULTI_HOST_FRAME, rgba8_out, stats);
    rtc_multi_destroy(m);
    return rc;
}

uint32_t rtc_rows_count(const rtc_camera_desc* camera, const rtc_rows* rows) {
    if (!camera) return 0;
    DRows dr;
    if (to_drows(*camera, rows, &dr) != RTC_OK) return 0;
    return dr.local_rows;
}

int rtc_color_at(const rtc_scene* scene, const double* rays, uint64_t n, double* rgb_out) {
    if (!scene || (n && (!rays || !rgb_out))) return set_err(RTC_ERR_INVALID, "null argument");
    std::string e;
    if (color_at_host(scene->dev, rays, n, rgb_out, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}

int rtc_intersect(const rtc_scene* scene, const double* rays, uint64_t n, uint32_t cap, double* t_out, int32_t* leaf_out,
                  uint32_t* counts) {
    if (!scene || (n && (!rays || !counts || (cap && (!t_out || !leaf_out))))) return set_err(RTC_ERR_INVALID, "null argument");
    std::string e;
    if (probe_intersect(scene->dev, rays, n, cap, t_out, leaf_out, counts, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}
int rtc_prepare_computations(const rtc_scene* scene, const double* rays, uint64_t n, rtc_computations* out) {
    if (!scene || (n && (!rays || !out))) return set_err(RTC_ERR_INVALID, "null argument");
    std::string e;
    if (probe_prepare(scene->dev, rays, n, out, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}
int rtc_normal_at(const rtc_scene* scene, int32_t leaf, const double* points, uint64_t n, double* normals_out) {
    if (!scene || (n && (!points || !normals_out))) return set_err(RTC_ERR_INVALID, "null argument");
    if (leaf < 0 || (uint64_t)leaf >= scene->info[0]) return set_err(RTC_ERR_INVALID, "no such leaf");
    std::string e;
    const int rc = probe_normal_at(scene->dev, scene->info[3], leaf, points, n, normals_out, &e);
    if (rc == -1) return set_err(RTC_ERR_INVALID, e);
    if (rc != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}

int rtc_selftest_shared_divisor(int device, uint64_t pairs, uint64_t seed, uint64_t* mismatches) {
    if (!mismatches) return set_err(RTC_ERR_INVALID, "null argument");
    std::string e;
    if (divisor_selftest(device, pairs, seed, mismatches, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}

int rtc_tally_count(void) { return tally_count(); }
int rtc_render_tally(const rtc_scene* scene, const rtc_camera_desc* camera, const rtc_rows* rows, uint64_t* counts) {
    if (!scene || !camera || !counts) return set_err(RTC_ERR_INVALID, "null argument");
    DRows dr;
    int rc = to_drows(*camera, rows, &dr);
    if (rc != RTC_OK) return rc;
    std::string e;
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    if (render_tally(scene->dev, to_dcamera(*camera), dr, (unsigned long long*)counts, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}

int rtc_measure_fp64_peak(int device, double* nofma_gflops, double* fma_gflops) {
    std::string e;
    if (measure_fp64_peak(device, nofma_gflops, fma_gflops, &e) != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}

/* ------------------------------------------------------------------------------------------------ 2. HOST MIRROR */
void rtc_translation(double x, double y, double z, double* o) { std::memcpy(o, translation(x, y, z).m, 128); }
void rtc_scaling(double x, double y, double z, double* o) { std::memcpy(o, scaling(x, y, z).m, 128); }
void rtc_rotation_x(double r, double* o) { std::memcpy(o, rotation_x(r).m, 128); }
void rtc_rotation_y(double r, double* o) { std::memcpy(o, rotation_y(r).m, 128); }
void rtc_rotation_z(double r, double* o) { std::memcpy(o, rotation_z(r).m, 128); }
void rtc_shearing(double xy, double xz, double yx, double yz, double zx, double zy, double* o) {
    std::memcpy(o, shearing(xy, xz, yx, yz, zx, zy).m, 128);
}
int rtc_view_transform(const double* from, const double* to, const double* up, double* o) {
    if (!from || !to || !up || !o) return set_err(RTC_ERR_INVALID, "null argument");
    Mat4 m = view_transform(point(from[0], from[1], from[2]), point(to[0], to[1], to[2]), vector(up[0], up[1], up[2]));
    std::memcpy(o, m.m, 128);
    return RTC_OK;
}
void rtc_matrix_mul(const double* a, const double* b, double* o) {
    Mat4 r = mul(Mat4::from(a), Mat4::from(b));
    std::memcpy(o, r.m, 128);
}
void rtc_matrix_transpose(const double* a, double* o) {
    Mat4 r = transpose(Mat4::from(a));
    std::memcpy(o, r.m, 128);
}
int rtc_matrix_inverse(const double* a, double* o) {
    Mat4 r;
    if (!inverse(Mat4::from(a), &r)) return set_err(RTC_ERR_PANIC, "matrix is not invertible: |det| < 1e-5 (src/matrix.rs:140)");
    std::memcpy(o, r.m, 128);
    return RTC_OK;
}
void rtc_matrix_mul_tuple(const double* a, const double* t, double* o) {
    Vec4 r = mul(Mat4::from(a), Vec4{t[0], t[1], t[2], t[3]});
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}

void rtc_material_default(rtc_material* m) { material_default(m); }
int rtc_material_set_pattern_transform(rtc_material* m, const double* t) {
    Mat4 inv;
    if (!inverse(Mat4::from(t), &inv)) return set_err(RTC_ERR_PANIC, "should be invertible (src/pattern.rs:65)");
    std::memcpy(m->pattern_transform, t, 128);
    std::memcpy(m->pattern_inverse, inv.m, 128);
    return RTC_OK;
}

rtc_shape* rtc_shape_new(int kind, double minimum, double maximum, int capped) {
    if (kind < RTC_SPHERE || kind > RTC_GROUP) {
        g_err = "unknown shape kind";
        return nullptr;
    }
    return new rtc_shape{shape_new(kind, minimum, maximum, capped != 0)};
}
rtc_shape* rtc_shape_triangle(const double* p1, const double* p2, const double* p3) {
    if (!p1 || !p2 || !p3) {
        g_err = "null argument";
        return nullptr;
    }
    return new rtc_shape{shape_triangle(p1, p2, p3)};
}
rtc_shape* rtc_shape_smooth_triangle(const double* p1, const double* p2, const double* p3, const double* n1, const double* n2,
                                     const double* n3) {
    if (!p1 || !p2 || !p3 || !n1 || !n2 || !n3) {
        g_err = "null argument";
        return nullptr;
    }
    return new rtc_shape{shape_smooth_triangle(p1, p2, p3, n1, n2, n3)};
}
void rtc_shape_free(rtc_shape* s) { delete s; }
int rtc_shape_set_transform(rtc_shape* s, const double* m) {
    if (!s || !s->s || !m) return set_err(RTC_ERR_INVALID, "null argument");
    try {
        shape_set_transform(s->s.get(), Mat4::from(m));
        return RTC_OK;
    } catch (const HostPanic& e) { return set_err(RTC_ERR_PANIC, e.what()); }
}
int rtc_shape_set_material(rtc_shape* s, const rtc_material* m) {
    if (!s || !s->s || !m) return set_err(RTC_ERR_INVALID, "null argument");
    shape_set_material(s->s.get(), *m);
    return RTC_OK;
}
int rtc_shape_push_shape(rtc_shape* group, rtc_shape* child) {
    if (!group || !group->s || !child || !child->s) return set_err(RTC_ERR_INVALID, "null argument");
    try {
        shape_push(group->s.get(), std::move(child->s));
        delete child;
        return RTC_OK;
    } catch (const HostPanic& e) { return set_err(RTC_ERR_PANIC, e.what()); }
}
uint64_t rtc_shape_leaf_count(const rtc_shape* s) { return (s && s->s) ? shape_leaf_count(s->s.get()) : 0; }

rtc_shape* rtc_obj_parse_str(const char* text, uint64_t len, uint64_t* ignored_lines) {
    if (!text && len) {
        g_err = "null argument";
        return nullptr;
    }
    try {
        ObjResult r = obj_parse(text, (size_t)len);
        if (ignored_lines) *ignored_lines = r.ignored_lines;
        return new rtc_shape{std::move(r.group)};
    } catch (const HostPanic& e) {
        set_err(RTC_ERR_PANIC, e.what());
        return nullptr;
    }
}
rtc_shape* rtc_obj_parse_file(const char* path, uint64_t* ignored_lines) {
    std::ifstream f(path ? path : "", std::ios::binary);
    if (!f) {
        set_err(RTC_ERR_PANIC, std::string("something went wrong reading ") + (path ? path : "(null)") + ". (src/obj_file.rs:25)");
        return nullptr;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    std::string text = ss.str();
    return rtc_obj_parse_str(text.data(), text.size(), ignored_lines);
}
rtc_shape* rtc_mesh_from_arrays(const double* verts, uint64_t nverts, const int32_t* faces, uint64_t nfaces) {
    if ((nverts && !verts) || (nfaces && !faces)) {
        g_err = "null argument";
        return nullptr;
    }
    auto def = shape_new(RTC_GROUP, 0., 0., false);
    for (uint64_t f = 0; f < nfaces; f++) {
        const double* p[3];
        for (int k = 0; k < 3; k++) {
            int64_t i = faces[f * 3 + k];
            if (i < 1 || (uint64_t)i > nverts) {
                set_err(RTC_ERR_PANIC, "index out of bounds (src/obj_file.rs:117)");
                return nullptr;
            }
            p[k] = verts + (i - 1) * 3;
        }
        shape_push(def.get(), shape_triangle(p[0], p[1], p[2]));
    }
    auto g = shape_new(RTC_GROUP, 0., 0., false);
    shape_push(g.get(), std::move(def));
    return new rtc_shape{std::move(g)};
}

rtc_shape* rtc_smooth_mesh_from_arrays(const double* verts, uint64_t nverts, const double* normals, uint64_t nnormals,
                                       const int32_t* faces, const int32_t* face_normals, uint64_t nfaces) {
    if ((nverts && !verts) || (nnormals && !normals) || (nfaces && (!faces || !face_normals))) {
        g_err = "null argument";
        return nullptr;
    }
    auto def = shape_new(RTC_GROUP, 0., 0., false);
    for (uint64_t f = 0; f < nfaces; f++) {
        const double *p[3], *n[3];
        for (int k = 0; k < 3; k++) {
            int64_t i = faces[f * 3 + k], j = face_normals[f * 3 + k];
            if (i < 1 || (uint64_t)i > nverts || j < 1 || (uint64_t)j > nnormals) {
                set_err(RTC_ERR_PANIC, "index out of bounds (src/obj_file.rs:117)");
                return nullptr;
            }
            p[k] = verts + (i - 1) * 3;
            n[k] = normals + (j - 1) * 3;
        }
        shape_push(def.get(), shape_smooth_triangle(p[0], p[1], p[2], n[0], n[1], n[2]));
    }
    auto g = shape_new(RTC_GROUP, 0., 0., false);
    shape_push(g.get(), std::move(def));
    return new rtc_shape{std::move(g)};
}

rtc_world* rtc_world_new(const double* light_position3, const double* light_intensity3) {
    if (!light_position3 || !light_intensity3) {
        g_err = "null argument";
        return nullptr;
    }
    rtc_world* w = new rtc_world();
    for (int k = 0; k < 3; k++) {
        w->w.light_position[k] = light_position3[k];
        w->w.light_intensity[k] = light_intensity3[k];
    }
    return w;
}
rtc_world* rtc_world_default(void) {
    rtc_world* w = new rtc_world();
    auto d = world_default();
    w->w = std::move(*d);
    return w;
}
static void world_drop_scenes(rtc_world* w) {  // w->mu held
    for (auto& kv : w->scenes) rtc_scene_destroy(kv.second);
    w->scenes.clear();
}
// w->mu held: the world's scene on `device`, marshalled and uploaded on first use
static int world_scene_locked(rtc_world* w, int device, rtc_scene** out) {
    auto it = w->scenes.find(device);
    if (it == w->scenes.end()) {
        Marshalled m;
        marshal_world(w->w, m);
        m.desc.recursion_limit = w->recursion_limit;
        rtc_scene* s = nullptr;
        int rc = rtc_scene_create_ex(&m.desc, device, w->build_flags, &s);
        if (rc != RTC_OK) return rc;
        it = w->scenes.emplace(device, s).first;
    }
    *out = it->second;
    return RTC_OK;
}
void rtc_world_free(rtc_world* w) {
    if (!w) return;
    world_drop_scenes(w);
    delete w;
}
int rtc_world_push(rtc_world* w, rtc_shape* s) {
    if (!w || !s || !s->s) return set_err(RTC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(w->mu);
    w->w.objects.push_back(std::move(s->s));
    delete s;
    world_drop_scenes(w);
    return RTC_OK;
}
int rtc_world_scene(rtc_world* w, int device, rtc_scene** out) {
    if (!w || !out) return set_err(RTC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(w->mu);
    return world_scene_locked(w, device, out);
}
int rtc_world_set_build(rtc_world* w, uint32_t flags) {
    if (!w) return set_err(RTC_ERR_INVALID, "null argument");
    if (flags & ~(uint32_t)RTC_BUILD_DEVICE_LBVH) return set_err(RTC_ERR_INVALID, "unknown build flag");
    std::lock_guard<std::mutex> lk(w->mu);
    if (flags != w->build_flags) world_drop_scenes(w);
    w->build_flags = flags;
    return RTC_OK;
}
int rtc_world_set_recursion_limit(rtc_world* w, uint32_t limit) {
    if (!w) return set_err(RTC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(w->mu);
    if (limit != w->recursion_limit) world_drop_scenes(w);
    w->recursion_limit = limit;
    return RTC_OK;
}
void rtc_world_drop_scenes(rtc_world* w) {
    if (!w) return;
    std::lock_guard<std::mutex> lk(w->mu);
    world_drop_scenes(w);
}
// The layer-1 description of a world — exactly what rtc_world_scene hands to rtc_scene_create, and what the Rust-side
// Camera::render patch builds from its &World.  The description borrows from the returned object.
int rtc_world_marshal(rtc_world* w, rtc_marshalled** out) {
    if (!w || !out) return set_err(RTC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(w->mu);
    rtc_marshalled* m = new rtc_marshalled();
    marshal_world(w->w, m->m);
    m->m.desc.recursion_limit = w->recursion_limit;
    *out = m;
    return RTC_OK;
}
const rtc_scene_desc* rtc_marshalled_desc(const rtc_marshalled* m) { return m ? &m->m.desc : nullptr; }
void rtc_marshalled_free(rtc_marshalled* m) { delete m; }
// The marshalled description of a world (what the Rust Camera::render patch would build), for tests of the flattener
// that need no device: returns counts {shapes, roots, transforms, materials, triangles}.
int rtc_world_describe(rtc_world* w, uint64_t n[5]) {
    if (!w || !n) return set_err(RTC_ERR_INVALID, "null argument");
    Marshalled m;
    marshal_world(w->w, m);
    n[0] = m.shapes.size();
    n[1] = w->w.objects.size();
    n[2] = m.transforms.size();
    n[3] = m.materials.size();
    n[4] = m.triangles.size();
    return RTC_OK;
}
// Host-only half of rtc_scene_create (validation, gate boxes, flattening, BVH) without the upload, for reports and CPU
// tests: n = {leaves, gates, meshes, mesh triangles, bvh nodes, bvh max depth, program nodes, distinct transforms}.
// Optionally copies the gate boxes (6 doubles each: lo xyz, hi xyz) into gates_out (capacity gates_cap boxes).
int rtc_world_flatten_info(rtc_world* w, uint64_t n[8], double* gates_out, uint64_t gates_cap) {
    if (!w || !n) return set_err(RTC_ERR_INVALID, "null argument");
    Marshalled m;
    marshal_world(w->w, m);
    m.desc.recursion_limit = w->recursion_limit;
    FlatScene flat;
    std::string e;
    int rc = flatten_scene(m.desc, flat, &e);
    if (rc != RTC_OK) return set_err(rc, e);
    trace_phases("rtc_world_flatten_info", flat);
    n[0] = flat.leaf_count;
    n[1] = flat.gates.size();
    n[2] = flat.meshes.size();
    n[3] = flat.tris.size();
    n[4] = flat.bvh.size();
    n[5] = (uint64_t)flat.bvh_max_depth;
    n[6] = flat.program.size();
    n[7] = flat.xforms.size();
    if (gates_out)
        for (uint64_t i = 0; i < flat.gates.size() && i < gates_cap; i++) {
            std::memcpy(gates_out + 6 * i, flat.gates[i].lo, 24);
            std::memcpy(gates_out + 6 * i + 3, flat.gates[i].hi, 24);
        }
    return RTC_OK;
}
int rtc_world_kernel_features(rtc_world* w, uint32_t features[2]) {
    if (!w || !features) return set_err(RTC_ERR_INVALID, "null argument");
    Marshalled m;
    marshal_world(w->w, m);
    m.desc.recursion_limit = w->recursion_limit;
    FlatScene flat;
    std::string e;
    int rc = flatten_scene(m.desc, flat, &e);
    if (rc != RTC_OK) return set_err(rc, e);
    features[0] = (uint32_t)flat.feature_mask;
    features[1] = (uint32_t)render_instance_mask(flat.feature_mask);
    return RTC_OK;
}
int rtc_world_color_at(rtc_world* w, const double* rays, uint64_t n, double* rgb_out) {
    if (!w) return set_err(RTC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(w->mu);  // the scene cannot be dropped under the call
    rtc_scene* s = nullptr;
    int rc = world_scene_locked(w, 0, &s);
    if (rc != RTC_OK) return rc;
    return rtc_color_at(s, rays, n, rgb_out);
}

rtc_camera* rtc_camera_new(uint64_t hsize, uint64_t vsize, double fov) {
    rtc_camera* c = new rtc_camera();
    c->c = *camera_new(hsize, vsize, fov);
    return c;
}
void rtc_camera_free(rtc_camera* c) { delete c; }
int rtc_camera_set_transform(rtc_camera* c, const double* m) {
    if (!c || !m) return set_err(RTC_ERR_INVALID, "null argument");
    try {
        camera_set_transform(&c->c, Mat4::from(m));
        return RTC_OK;
    } catch (const HostPanic& e) { return set_err(RTC_ERR_PANIC, e.what()); }
}
void rtc_camera_desc_get(const rtc_camera* c, rtc_camera_desc* out) {
    out->hsize = (uint32_t)c->c.hsize;
    out->vsize = (uint32_t)c->c.vsize;
    std::memcpy(out->inverse, c->c.inverse.m, 128);
    out->half_width = c->c.half_width;
    out->half_height = c->c.half_height;
    out->pixel_size = c->c.pixel_size;
}

// Pinned host buffers are expensive to create (cudaHostAlloc of a 1080p f64 canvas: milliseconds — more than the frame):
// a freed canvas leaves its pinned buffers in a small pool and the next canvas of that size takes them back.
namespace {
struct PinnedPool {
    struct Entry {
        void* p;
        size_t bytes;
    };
    std::mutex mu;
    std::vector<Entry> free_list;
    size_t pooled = 0;
    static constexpr size_t kMaxPooled = (size_t)4 << 30;
    void* take(size_t bytes) {
        {
            std::lock_guard<std::mutex> lk(mu);
            for (size_t i = 0; i < free_list.size(); i++)
                if (free_list[i].bytes >= bytes && free_list[i].bytes <= bytes + bytes / 8 + 4096) {
                    void* p = free_list[i].p;
                    pooled -= free_list[i].bytes;
                    free_list.erase(free_list.begin() + i);
                    return p;
                }
        }
        return pinned_alloc(bytes);
    }
    void give(void* p, size_t bytes) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (pooled + bytes <= kMaxPooled && free_list.size() < 16) {
                free_list.push_back(Entry{p, bytes});
                pooled += bytes;
                return;
            }
        }
        pinned_free(p);
    }
};
PinnedPool& pinned_pool() {
    static PinnedPool* pool = new PinnedPool();  // never destroyed: the CUDA runtime may already be gone at exit
    return *pool;
}
}  // namespace

static rtc_canvas* canvas_alloc(uint64_t width, uint64_t height, bool want_f64, bool pinned, bool want_rgba8 = true) {
    rtc_canvas* cv = new rtc_canvas();
    cv->width = width;
    cv->height = height;
    const size_t px = (size_t)width * height;
    if (want_rgba8 && pinned) {
        cv->rgba8 = (uint8_t*)pinned_pool().take(px * 4 ? px * 4 : 1);
        cv->rgba_pinned = cv->rgba8 != nullptr;
    }
    if (want_rgba8 && !cv->rgba8) cv->rgba8 = (uint8_t*)std::malloc(px * 4 ? px * 4 : 1);
    if (want_f64) {
        if (pinned) {
            cv->rgb = (double*)pinned_pool().take(px * 24 ? px * 24 : 1);
            cv->rgb_pinned = cv->rgb != nullptr;
        }
        if (!cv->rgb) cv->rgb = (double*)std::malloc(px * 24 ? px * 24 : 1);
    }
    return cv;
}

// The canvas's RGBA8 pixels, quantised from its f64 colours on first use (canvas.rs:61-63 per channel, rows in parallel).
static const uint8_t* canvas_rgba8(const rtc_canvas* c) {
    std::lock_guard<std::mutex> lk(c->lazy);
    if (c->rgba8 || !c->rgb) return c->rgba8;
    const size_t px = (size_t)c->width * c->height;
    uint8_t* out = (uint8_t*)std::malloc(px * 4 ? px * 4 : 1);
    const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(16, std::thread::hardware_concurrency()), px / 65536));
    auto rows = [&](unsigned t) {
        for (size_t i = px * t / nt; i < px * (t + 1) / nt; i++) {
            for (int k = 0; k < 3; k++) out[4 * i + k] = quantise_channel(c->rgb[3 * i + k]);
            out[4 * i + 3] = 255;
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(rows, t);
    rows(0);
    for (auto& t : th) t.join();
    c->rgba8 = out;
    c->rgba_pinned = false;
    return out;
}

// The world's scene on EVERY device 0 .. ngpus-1: marshalled and flattened once, uploaded to the devices that do not hold it
// yet in parallel (one host thread per device).
static int world_scenes_all_locked(rtc_world* w, int ngpus, std::vector<rtc_scene*>* out) {
    std::vector<int> missing;
    for (int g = 0; g < ngpus; g++)
        if (!w->scenes.count(g)) missing.push_back(g);
    if (!missing.empty()) {
        Marshalled m;
        marshal_world(w->w, m);
        m.desc.recursion_limit = w->recursion_limit;
        std::vector<rtc_scene*> made(missing.size(), nullptr);
        const auto drop_made = [&] {
            for (rtc_scene*& s : made) {
                rtc_scene_destroy(s);
                s = nullptr;
            }
        };
        const int rc = flatten_and_upload(&m.desc, w->build_flags, "rtc_camera_render(all devices)",
                                          [&](const FlatScene& flat, std::string* e) {
            std::vector<int> rcs(missing.size(), 0);
            std::vector<std::string> errs(missing.size());
            const auto upload_one = [&](size_t k) {
                DeviceScene* dev = nullptr;
                rcs[k] = device_scene_create(flat, missing[k], &dev, &errs[k], nullptr);
                if (rcs[k] == 0) made[k] = scene_handle(flat, dev);
            };
            std::vector<std::thread> th;
            for (size_t k = 1; k < missing.size(); k++) th.emplace_back(upload_one, k);
            upload_one(0);
            for (auto& t : th) t.join();
            for (size_t k = 0; k < missing.size(); k++)
                if (rcs[k] != 0) {  // (a tree too deep for the traversal stack on one device is too deep on all of them)
                    drop_made();
                    *e = errs[k];
                    return rcs[k];
                }
            return 0;
        });
        if (rc != RTC_OK) return rc;
        for (size_t k = 0; k < missing.size(); k++) w->scenes.emplace(missing[k], made[k]);
    }
    out->clear();
    for (int g = 0; g < ngpus; g++) out->push_back(w->scenes.at(g));
    return RTC_OK;
}

// Camera::render sharded over every CUDA device of this process (RTC_DEVICE_ALL): the frame's 8-row bands are dealt
// cyclically to the devices, every device renders its bands and its own copy engine writes them to their frame positions in
// the ONE pinned canvas over its own PCIe link (render_host with the frame layout, one host thread per device).
static int camera_render_all_locked(const rtc_camera* c, rtc_world* w, int ngpus, int want_f64, rtc_canvas** out,
                                    rtc_stats* stats) {
    std::vector<rtc_scene*> scenes;
    int rc = world_scenes_all_locked(w, ngpus, &scenes);
    if (rc != RTC_OK) return rc;
    rtc_camera_desc cd;
    rtc_camera_desc_get(c, &cd);
    if (!affine_camera(cd)) return set_err(RTC_ERR_UNSUPPORTED, "non-affine camera transform");
    rtc_canvas* cv = canvas_alloc(c->c.hsize, c->c.vsize, want_f64 != 0, true, want_f64 == 0);
    std::vector<DeviceScene*> devs;
    for (rtc_scene* s : scenes) devs.push_back(s->dev);
    LaunchStats ls;
    std::string e;
    if (render_host_sharded(devs.data(), ngpus, to_dcamera(cd), cv->rgba8, cv->rgb, stats ? &ls : nullptr, &e) != 0) {
        rtc_canvas_free(cv);
        return set_err(RTC_ERR_CUDA, e);
    }
    fill_stats(ls, stats);
    *out = cv;
    return RTC_OK;
}

int rtc_camera_render(const rtc_camera* c, rtc_world* w, int device, int want_f64, rtc_canvas** out, rtc_stats* stats) {
    if (!c || !w || !out) return set_err(RTC_ERR_INVALID, "null argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(w->mu);  // the scene cannot be dropped under the render
    if (device == RTC_DEVICE_ALL) {
        std::string ce;
        const int have = cuda_device_count(&ce);
        if (have < 1) return set_err(RTC_ERR_CUDA, ce.empty() ? "no CUDA device" : ce);
        if (have > 1) return camera_render_all_locked(c, w, have, want_f64, out, stats);
        device = 0;
    }
    rtc_scene* s = nullptr;
    int rc = world_scene_locked(w, device, &s);
    if (rc != RTC_OK) return rc;
    rtc_camera_desc cd;
    rtc_camera_desc_get(c, &cd);
    // with the f64 colours wanted, only they are brought to the host (canvas_rgba8 quantises them on demand)
    rtc_canvas* cv = canvas_alloc(c->c.hsize, c->c.vsize, want_f64 != 0, true, want_f64 == 0);
    rc = rtc_render(s, &cd, nullptr, cv->rgba8, cv->rgb, stats);
    if (rc != RTC_OK) {
        rtc_canvas_free(cv);
        return rc;
    }
    *out = cv;
    return RTC_OK;
}

rtc_canvas* rtc_canvas_new(uint64_t width, uint64_t height) {  // canvas.rs:12-18: all BLACK
    rtc_canvas* cv = canvas_alloc(width, height, true, false);
    const size_t px = (size_t)width * height;
    for (size_t i = 0; i < px * 3; i++) cv->rgb[i] = 0.;
    for (size_t i = 0; i < px; i++) {
        cv->rgba8[4 * i] = cv->rgba8[4 * i + 1] = cv->rgba8[4 * i + 2] = 0;
        cv->rgba8[4 * i + 3] = 255;
    }
    return cv;
}
void rtc_canvas_free(rtc_canvas* c) {
    if (!c) return;
    const size_t px = (size_t)c->width * c->height;
    if (c->rgb) c->rgb_pinned ? pinned_pool().give(c->rgb, px * 24 ? px * 24 : 1) : std::free(c->rgb);
    if (c->rgba8) c->rgba_pinned ? pinned_pool().give(c->rgba8, px * 4 ? px * 4 : 1) : std::free(c->rgba8);
    delete c;
}
uint64_t rtc_canvas_width(const rtc_canvas* c) { return c ? c->width : 0; }
uint64_t rtc_canvas_height(const rtc_canvas* c) { return c ? c->height : 0; }
int rtc_canvas_get_pixel(const rtc_canvas* c, uint64_t x, uint64_t y, double* rgb3) {
    if (!c || !rgb3) return set_err(RTC_ERR_INVALID, "null argument");
    if (x >= c->width || y >= c->height) return set_err(RTC_ERR_PANIC, "index out of bounds (src/canvas.rs:21)");
    if (!c->rgb) return set_err(RTC_ERR_INVALID, "canvas was rendered with want_f64 = 0");
    std::memcpy(rgb3, c->rgb + 3 * (x + y * c->width), 24);
    return RTC_OK;
}
int rtc_canvas_set_pixel(rtc_canvas* c, uint64_t x, uint64_t y, const double* rgb3) {
    if (!c || !rgb3) return set_err(RTC_ERR_INVALID, "null argument");
    if (x >= c->width || y >= c->height) return set_err(RTC_ERR_PANIC, "index out of bounds (src/canvas.rs:25)");
    const size_t i = x + y * c->width;
    if (c->rgb) std::memcpy(c->rgb + 3 * i, rgb3, 24);
    std::lock_guard<std::mutex> lk(c->lazy);
    if (c->rgba8) {  // (not materialised yet: it will be quantised from rgb, which now holds the new colour)
        for (int k = 0; k < 3; k++) c->rgba8[4 * i + k] = quantise_channel(rgb3[k]);
        c->rgba8[4 * i + 3] = 255;
    }
    return RTC_OK;
}
const double* rtc_canvas_pixels_f64(const rtc_canvas* c) { return c ? c->rgb : nullptr; }
const uint8_t* rtc_canvas_pixels_rgba8(const rtc_canvas* c) { return c ? canvas_rgba8(c) : nullptr; }
char* rtc_canvas_to_ppm(const rtc_canvas* c, uint64_t* len) {
    if (!c || !len) {
        g_err = "null argument";
        return nullptr;
    }
    std::string s = ppm_from_rgba8(canvas_rgba8(c), c->width, c->height);
    char* out = (char*)std::malloc(s.size() + 1);
    std::memcpy(out, s.data(), s.size());
    out[s.size()] = 0;
    *len = s.size();
    return out;
}
// canvas.rs:28-58 for a caller-owned RGBA8 frame (e.g. the gathered multi-GPU frame)
char* rtc_ppm_from_rgba8(const uint8_t* rgba8, uint64_t width, uint64_t height, uint64_t* len) {
    if (!rgba8 || !len) {
        g_err = "null argument";
        return nullptr;
    }
    std::string s = ppm_from_rgba8(rgba8, width, height);
    char* out = (char*)std::malloc(s.size() + 1);
    std::memcpy(out, s.data(), s.size());
    out[s.size()] = 0;
    *len = s.size();
    return out;
}
uint64_t rtc_ppm_max_bytes(uint64_t width, uint64_t height) { return ppm_max_bytes(width, height); }
int rtc_ppm_encode_device(int device, const void* d_rgba8, uint64_t width, uint64_t height, void* cuda_stream,
                          char* out_host, uint64_t capacity, uint64_t* len) {
    if (!d_rgba8 || !out_host || !len) return set_err(RTC_ERR_INVALID, "null argument");
    std::string e;
    int rc = ppm_encode_device(device, d_rgba8, width, height, cuda_stream, out_host, capacity, len, &e);
    if (rc == -1) return set_err(RTC_ERR_INVALID, "output buffer too small (see rtc_ppm_max_bytes)");
    if (rc != 0) return set_err(RTC_ERR_CUDA, e);
    return RTC_OK;
}
void* rtc_pinned_alloc(uint64_t bytes) { return pinned_alloc((size_t)bytes); }
void rtc_pinned_free(void* p) { pinned_free(p); }
void rtc_free(void* p) { std::free(p); }

}  // extern "C"
