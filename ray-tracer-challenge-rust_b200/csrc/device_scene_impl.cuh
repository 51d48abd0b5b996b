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

struct DeviceScene {
    int device = 0;
    int sm_count = 0;
    void* slab = nullptr;
    size_t slab_size = 0;
    DScene view{};
    DQueue* queue = nullptr;
    // grow-only output scratch for render_host
    void* out8 = nullptr;
    size_t out8_size = 0;
    void* out64 = nullptr;
    size_t out64_size = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::mutex mu;
};

inline std::string cuda_err_string(const char* what, cudaError_t e) {
    return std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
}

}  // namespace rtc
