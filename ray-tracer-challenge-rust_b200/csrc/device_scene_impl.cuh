// device_scene_impl.cuh — the device-resident scene object shared by the CUDA translation units (private).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>

#include "device_scene.h"

namespace rtc {

// Work queue + exact ray counters of one launch.  The queue cleans up after itself: the last CTA of a launch to finish
// moves the counters into `result` and zeroes everything else, so the next launch on the same slot needs no memset (one
// API call and one GPU-side bubble less per frame).
struct DQueue {
    unsigned long long primary, shadow, reflect, refract;  // live counters of the running launch
    unsigned long long result[4];                          // the finished launch's counters
    unsigned int next_tile;
    unsigned int done_ctas;
    unsigned int pad[2];
};

// Slots 0 .. kQueueSlots-2 belong to ONE caller stream each (launches of a stream are ordered, so its slot is never in use
// by two kernels); the last slot is shared by any further streams, with an event between consecutive launches.
constexpr int kQueueSlots = 16;
constexpr int kHostChunks = 8;  // launches a host-output render is cut into at most (render_host)

// Per-device state created on first use and kept for the life of the process: creating streams, events and querying
// device properties costs milliseconds, a frame costs less.
struct DeviceContext {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;  // the library's own stream (uploads, host-output renders)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_stream = nullptr;  // device->host copies that overlap the next chunk's kernel (render_host)
    cudaEvent_t chunk_done[kHostChunks] = {}, copy_done = nullptr;
    DQueue* queues = nullptr;       // kQueueSlots self-cleaning work queues (see DQueue)
    cudaStream_t slot_stream[kQueueSlots] = {};  // which stream owns slot k
    int slots_taken = 0;
    cudaEvent_t shared_slot_free = nullptr;      // recorded after every launch on the shared (last) slot
    // grow-only scratch
    void* out8 = nullptr;
    size_t out8_size = 0;
    void* out64 = nullptr;
    size_t out64_size = 0;
    void* staging = nullptr;        // pinned upload staging
    size_t staging_size = 0;
    std::mutex mu;
};

struct DeviceScene {
    DeviceContext* ctx = nullptr;
    int device = 0;
    int sm_count = 0;
    void* slab = nullptr;           // all tables, one stream-ordered allocation
    size_t slab_size = 0;
    size_t upload_bytes = 0;        // host->device bytes the create call copied (tables, device-build inputs)
    DScene view{};
    double hot_lo[3] = {0, 0, 0}, hot_hi[3] = {0, 0, 0};  // FlatScene::hot_*; hot_valid: finite and not empty
    bool hot_valid = false;
    int feature_mask = 0;           // FEAT_* bits of what the flattened world contains (picks the kernel instantiation)
    cudaStream_t stream = nullptr;  // == ctx->stream
    std::mutex& mu() { return ctx->mu; }
};

// Every entry point that selects a device puts the caller's current device back on the way out: the host program (torch,
// another library, the caller's own CUDA code) keeps the device it had.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            prev = -1;
            cudaGetLastError();
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

inline std::string cuda_err_string(const char* what, cudaError_t e) {
    return std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
}

}  // namespace rtc
