// render.cuh — host-callable interface of the CUDA translation unit (render.cu).  Internal to librtc_b200.so; the
// public boundary is include/rtc.h.
#pragma once
#include <cstdint>
#include <string>

#include "../../include/rtc.h"
#include "device_scene.h"

namespace rtc {

struct FlatScene;

// Device-resident copy of a FlatScene (one cudaMalloc slab) plus per-scene scratch (work-queue counter, ray counters).
struct DeviceScene;

struct LaunchStats {
    unsigned long long primary = 0, shadow = 0, reflect = 0, refract = 0;
    unsigned long long launches = 0;
    double device_ms = 0.;
};

int cuda_device_count(std::string* err);
int enable_peer_access(int device, int peer, std::string* err);
int frame_share_create(int device, uint64_t bytes, void** d_ptr, unsigned char* handle64, std::string* err);
int frame_share_open(int device, const unsigned char* handle64, void** d_ptr, std::string* err);
int frame_share_close(int device, void* d_ptr, int owner, std::string* err);
// -3: CUDA error; kDeviceBuildTooDeep: a device-built mesh came out deeper than the traversal stack (rebuild on the host).
// *built_depth (optional): depth of the deepest device-built BVH.
constexpr int kDeviceBuildTooDeep = -5;
int device_scene_create(const FlatScene& flat, int device, DeviceScene** out, std::string* err, int* built_depth = nullptr);
void device_scene_destroy(DeviceScene* s);
uint64_t device_scene_bytes(const DeviceScene* s);
uint64_t device_scene_upload_bytes(const DeviceScene* s);
int device_scene_device(const DeviceScene* s);

// Camera::render for the rows selected by `rows` into device buffers (either may be null), asynchronously on `stream`.
// With `stats` non-null the call brackets the kernel with CUDA events, synchronises the stream and fills *stats.
int render_device(DeviceScene* s, const DCamera& cam, const DRows& rows, void* d_rgba8, void* d_rgb_f64, void* stream,
                  LaunchStats* stats, std::string* err);
// The two halves of a timed render_device, for callers that keep several devices busy at once (multi.cu): begin enqueues
// (stream null = the scene's own library stream), end waits and reads counters and kernel time.
int render_device_begin(DeviceScene* s, const DCamera& cam, const DRows& rows, void* d_rgba8, void* d_rgb_f64, void* stream,
                        void** token, std::string* err);
int render_device_end(DeviceScene* s, void* stream, void* token, LaunchStats* stats, std::string* err);
void* device_scene_stream(const DeviceScene* s);
// Same with host outputs (pinned or pageable), including the device->host copies.  rows.frame_layout: the host pointers
// address the whole frame and every rendered band is copied to its frame position.
int render_host(DeviceScene* s, const DCamera& cam, const DRows& rows, uint8_t* rgba8, double* rgb_f64,
                LaunchStats* stats, std::string* err);
// World::color_at for explicit rays (host in, host out).
int color_at_host(DeviceScene* s, const double* rays, uint64_t n, double* rgb, std::string* err);

// One frame sharded over several devices of this process (multi.cu).
struct MultiRenderer;
int multi_create(const FlatScene& flat, int ngpus, MultiRenderer** out, std::string* err);
void multi_destroy(MultiRenderer* m);
int multi_render(MultiRenderer* m, const DCamera& cam, bool to_device_frame, LaunchStats* stats, double* frame_ms,
                 std::string* err);
// The same frame into caller host buffers (RGBA8 and / or the f64 colours), every device copying its own bands to their
// frame positions; and the underlying call over any set of device scenes of this process.
int multi_render_host(MultiRenderer* m, const DCamera& cam, uint8_t* rgba8, double* rgb_f64, LaunchStats* stats,
                      std::string* err);
int render_host_sharded(DeviceScene* const* scenes, int n, const DCamera& cam, uint8_t* rgba8, double* rgb_f64,
                        LaunchStats* stats, std::string* err);
void* multi_device_frame(const MultiRenderer* m);
const void* multi_host_frame(const MultiRenderer* m);
int multi_device_count(const MultiRenderer* m);

// Stream-ordered counters for sharded renders across processes (one per GPU): a render launch can bump a uint32 counter in
// any mapped device memory when its pixels are visible system-wide (DRows.notify); these wait for / set such counters.
//   wait: the stream's later work runs once *d_counter >= at_least (cuStreamWaitValue32 when the driver offers it — no
//         kernel occupies the GPU while waiting — else a one-thread polling kernel);
//   set / add: one tiny kernel stores `value` to (adds 1 to) each of up to 16 counters with system scope.
int stream_counter_wait(int device, void* stream, void* d_counter, uint32_t at_least, std::string* err);
int stream_counters_set(int device, void* stream, void* const* d_counters, uint32_t n, uint32_t value, std::string* err);
int stream_counters_add(int device, void* stream, void* const* d_counters, uint32_t n, std::string* err);

// Feature mask of the render_kernel instantiation a scene with this FEAT_* mask is rendered by (render_launch.cuh); -1: none.
int render_instance_mask(int feature_mask);

// Single-ray probes (probe.cu): World::intersect's sorted list, prepare_computations of the hit, Shape::normal_at.
int probe_intersect(DeviceScene* s, const double* rays, uint64_t n, uint32_t cap, double* t_out, int32_t* leaf_out,
                    uint32_t* counts, std::string* err);
int probe_prepare(DeviceScene* s, const double* rays, uint64_t n, rtc_computations* out, std::string* err);
int probe_normal_at(DeviceScene* s, uint64_t n_tris, int32_t leaf, const double* points, uint64_t n, double* out,
                    std::string* err);

// Self-test of rt_core.cuh's SharedDivisor against the compiler's f64 division over `pairs` generated operand pairs.
int divisor_selftest(int device, uint64_t pairs, uint64_t seed, uint64_t* mismatches, std::string* err);

// Work tallies of one frame (render_tally.cu): counts[tally_count()] in TallyIndex order (rt_core.cuh).
int tally_count();
int render_tally(DeviceScene* s, const DCamera& cam, const DRows& rows, unsigned long long* counts, std::string* err);

int measure_fp64_peak(int device, double* nofma_gflops, double* fma_gflops, std::string* err);

// Canvas::to_ppm on the device (ppm_encode.cu): RGBA8 frame in HBM -> P3 text in out_host.  -1: capacity too small
// (*len is set); -3: CUDA error.
uint64_t ppm_max_bytes(uint64_t width, uint64_t height);
int ppm_encode_device(int device, const void* d_rgba8, uint64_t width, uint64_t height, void* stream, char* out_host,
                      uint64_t capacity, uint64_t* len, std::string* err);

// Pinned host memory for frame buffers (so device->host copies run at PCIe speed); falls back to nothing — returns null
// on failure and the caller reports RTC_ERR_CUDA.
void* pinned_alloc(size_t bytes);
// cudaHostRegister(portable) / cudaHostUnregister of caller-mapped host memory (rtc_host_share_*).
int host_register(int device, void* p, size_t bytes, std::string* err);
int host_unregister(void* p, std::string* err);
void pinned_free(void* p);

}  // namespace rtc
