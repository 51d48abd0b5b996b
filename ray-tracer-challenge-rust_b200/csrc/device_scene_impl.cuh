// device_scene_impl.cuh — the device-resident scene object shared by the CUDA translation units (private).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>

#include "device_scene.h"

namespace rtc {

// work queue + exact ray counters of one launch
struct DQueue {
    unsigned long long primary, shadow, reflect, refract;
    unsigned int next_tile;
    unsigned int pad;
};

constexpr int kQueueSlots = 16;

// Per-device state created on first use and kept for the life of the process: creating streams, events and querying
// device properties costs milliseconds, a frame costs less.
struct DeviceContext {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;  // the library's own stream (uploads, host-output renders)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_stream = nullptr;  // device->host copies that overlap the next chunk's kernel (render_host)
    cudaEvent_t chunk_done = nullptr, copy_done = nullptr;
    DQueue* queues = nullptr;       // kQueueSlots work queues handed out round-robin (one per in-flight launch)
    unsigned next_queue = 0;
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
    int feature_mask = 0;           // FEAT_* bits of what the flattened world contains (picks the kernel instantiation)
    cudaStream_t stream = nullptr;  // == ctx->stream
    std::mutex& mu() { return ctx->mu; }
};

inline std::string cuda_err_string(const char* what, cudaError_t e) {
    return std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
}

}  // namespace rtc
