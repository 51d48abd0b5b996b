// multi.cu — one frame sharded over several GPUs of ONE process (rtc_multi_* / rtc_render_multi in include/rtc.h).
//
// Pixels are independent (camera.rs:70-76 has no loop-carried state): the frame is cut into bands of kBandRows rows dealt
// cyclically to the devices (band b -> device b mod G: meshes sit mid-frame, contiguous slabs would not balance), the scene
// (< 2 MB, flattened ONCE) is uploaded to every device and every device renders its bands with one launch.  Two ways out:
//   HOST frame    every device copies its own bands into one pinned host frame with ONE strided copy over its own PCIe
//                 link (cudaMemcpy2DAsync: compact bands -> every G-th band of the frame), all links in parallel;
//   DEVICE frame  the kernels store each pixel straight into a frame on device 0 through NVLink peer mappings
//                 (RTC_ROWS_FRAME) — no gather, no reassembly; device 0's stream waits for the other devices' events.
// No NCCL, no torch: plain CUDA runtime, so a Rust / C host can shard a frame with one call.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "device_scene_impl.cuh"
#include "flat_scene.hpp"
#include "render.cuh"

namespace rtc {

namespace {
constexpr uint32_t kBandRows = 8;
#define MULTI_CUDA(call)                                  \
    do {                                                  \
        cudaError_t e_ = (call);                          \
        if (e_ != cudaSuccess) {                          \
            if (err) *err = cuda_err_string(#call, e_);   \
            return -3;                                    \
        }                                                 \
    } while (0)
}  // namespace

struct MultiRenderer {
    int n = 0;
    std::vector<DeviceScene*> scene;
    std::vector<void*> local;        // per device: compact buffer of its bands (HOST frame mode)
    std::vector<size_t> local_bytes;
    std::vector<cudaEvent_t> done;   // per device: its part of the current frame is complete
    void* d_frame = nullptr;         // on device 0 (DEVICE frame mode)
    size_t d_frame_bytes = 0;
    void* h_frame = nullptr;         // pinned, portable
    size_t h_frame_bytes = 0;
    bool peers = false;
};

void multi_destroy(MultiRenderer* m) {
    if (!m) return;
    DeviceGuard guard_;
    for (int g = 0; g < (int)m->scene.size(); g++) {
        if (!m->scene[g]) continue;
        cudaSetDevice(device_scene_device(m->scene[g]));
        cudaStreamSynchronize((cudaStream_t)device_scene_stream(m->scene[g]));
        if (g < (int)m->local.size() && m->local[g]) cudaFree(m->local[g]);
        if (g < (int)m->done.size() && m->done[g]) cudaEventDestroy(m->done[g]);
        if (g == 0 && m->d_frame) cudaFree(m->d_frame);
        device_scene_destroy(m->scene[g]);
    }
    if (m->h_frame) cudaFreeHost(m->h_frame);
    delete m;
}

// Uploads the flattened scene to devices 0 .. ngpus-1.  Returns 0, -3 (CUDA, *err set) or kDeviceBuildTooDeep.
int multi_create(const FlatScene& flat, int ngpus, MultiRenderer** out, std::string* err) {
    DeviceGuard guard_;
    MultiRenderer* m = new MultiRenderer();
    m->n = ngpus;
    m->scene.assign(ngpus, nullptr);
    m->local.assign(ngpus, nullptr);
    m->local_bytes.assign(ngpus, 0);
    m->done.assign(ngpus, nullptr);
    {  // the uploads run side by side, one host thread per device (each is a pinned staging copy + one H2D copy + a sync)
        std::vector<int> rcs(ngpus, 0);
        std::vector<std::string> errs(ngpus);
        const auto upload_one = [&](int g) { rcs[g] = device_scene_create(flat, g, &m->scene[g], &errs[g], nullptr); };
        std::vector<std::thread> th;
        for (int g = 1; g < ngpus; g++) th.emplace_back(upload_one, g);
        upload_one(0);
        for (auto& t : th) t.join();
        for (int g = 0; g < ngpus; g++)
            if (rcs[g] != 0) {
                if (err) *err = errs[g];
                multi_destroy(m);
                return rcs[g];
            }
    }
    for (int g = 0; g < ngpus; g++) {
        cudaError_t e = cudaSetDevice(g);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->done[g], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            if (err) *err = cuda_err_string("cudaEventCreate", e);
            multi_destroy(m);
            return -3;
        }
    }
    *out = m;
    return 0;
}

static int ensure_peers(MultiRenderer* m, std::string* err) {
    if (m->peers) return 0;
    for (int g = 1; g < m->n; g++) {
        const int rc = enable_peer_access(g, 0, err);
        if (rc != 0) return rc;
    }
    m->peers = true;
    return 0;
}

// One frame.  to_device_frame: DEVICE frame mode (see the file header), else HOST frame mode.  The finished frame is
// m->d_frame (device 0) or m->h_frame (pinned host); frame_ms = host wall clock from the first launch to the last byte.
int multi_render(MultiRenderer* m, const DCamera& cam, bool to_device_frame, LaunchStats* stats, double* frame_ms,
                 std::string* err) {
    DeviceGuard guard_;
    const size_t row_bytes = (size_t)cam.hsize * 4, frame_bytes = row_bytes * cam.vsize;
    const uint32_t nbands = (cam.vsize + kBandRows - 1) / kBandRows;
    if (to_device_frame) {
        const int rc = ensure_peers(m, err);
        if (rc != 0) return rc;
        if (m->d_frame_bytes < frame_bytes) {
            MULTI_CUDA(cudaSetDevice(0));
            if (m->d_frame) cudaFree(m->d_frame);
            m->d_frame = nullptr;
            m->d_frame_bytes = 0;
            MULTI_CUDA(cudaMalloc(&m->d_frame, frame_bytes ? frame_bytes : 1));
            m->d_frame_bytes = frame_bytes;
        }
    } else if (m->h_frame_bytes < frame_bytes) {
        if (m->h_frame) cudaFreeHost(m->h_frame);
        m->h_frame = nullptr;
        m->h_frame_bytes = 0;
        MULTI_CUDA(cudaHostAlloc(&m->h_frame, frame_bytes ? frame_bytes : 1, cudaHostAllocPortable));
        m->h_frame_bytes = frame_bytes;
    }
    std::vector<void*> token(m->n, nullptr);
    std::vector<DRows> rows(m->n);
    const auto t0 = std::chrono::steady_clock::now();
    for (int g = 0; g < m->n; g++) {
        DRows& r = rows[g];
        r.band_rows = kBandRows;
        r.band_first = (uint32_t)g;
        r.band_stride = (uint32_t)m->n;
        uint32_t local = 0;
        for (uint32_t b = (uint32_t)g; b < nbands; b += (uint32_t)m->n)
            local += std::min(kBandRows, cam.vsize - b * kBandRows);
        r.local_rows = local;
        r.frame_layout = to_device_frame ? 1u : 0u;
        r.row_begin = 0;
        r.row_count = local;
        r.pad = 0;
        void* out8 = m->d_frame;
        if (!to_device_frame) {
            const size_t need = (size_t)local * row_bytes;
            if (m->local_bytes[g] < need) {
                MULTI_CUDA(cudaSetDevice(g));
                if (m->local[g]) cudaFree(m->local[g]);
                m->local[g] = nullptr;
                m->local_bytes[g] = 0;
                MULTI_CUDA(cudaMalloc(&m->local[g], need ? need : 1));
                m->local_bytes[g] = need;
            }
            out8 = m->local[g];
        }
        const int rc = render_device_begin(m->scene[g], cam, r, out8, nullptr, nullptr, &token[g], err);
        if (rc != 0) return rc;
        cudaStream_t st = (cudaStream_t)device_scene_stream(m->scene[g]);
        MULTI_CUDA(cudaSetDevice(g));
        if (!to_device_frame && local) {
            // local band k -> frame band g + k*G: one strided copy; a ragged last band (vsize not a multiple of the band
            // height) goes separately
            const uint32_t my_bands = (nbands - (uint32_t)g + (uint32_t)m->n - 1) / (uint32_t)m->n;
            const bool ragged = (cam.vsize % kBandRows) != 0 && ((nbands - 1) % (uint32_t)m->n) == (uint32_t)g;
            const uint32_t full = ragged ? my_bands - 1 : my_bands;
            const size_t band_bytes = row_bytes * kBandRows;
            unsigned char* dst = (unsigned char*)m->h_frame + (size_t)g * band_bytes;
            if (full)
                MULTI_CUDA(cudaMemcpy2DAsync(dst, band_bytes * (size_t)m->n, m->local[g], band_bytes, band_bytes, full,
                                             cudaMemcpyDeviceToHost, st));
            if (ragged)
                MULTI_CUDA(cudaMemcpyAsync((unsigned char*)m->h_frame + (size_t)(nbands - 1) * band_bytes,
                                           (unsigned char*)m->local[g] + (size_t)full * band_bytes,
                                           (size_t)(cam.vsize % kBandRows) * row_bytes, cudaMemcpyDeviceToHost, st));
        }
        MULTI_CUDA(cudaEventRecord(m->done[g], st));
    }
    // device 0's stream owns the finished frame: it waits for every other device's part
    cudaStream_t s0 = (cudaStream_t)device_scene_stream(m->scene[0]);
    MULTI_CUDA(cudaSetDevice(0));
    for (int g = 1; g < m->n; g++) MULTI_CUDA(cudaStreamWaitEvent(s0, m->done[g], 0));
    MULTI_CUDA(cudaStreamSynchronize(s0));
    if (frame_ms) *frame_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (stats) {
        *stats = LaunchStats{};
        for (int g = 0; g < m->n; g++) {
            LaunchStats ls;
            const int rc = render_device_end(m->scene[g], nullptr, token[g], &ls, err);
            if (rc != 0) return rc;
            stats->primary += ls.primary;
            stats->shadow += ls.shadow;
            stats->reflect += ls.reflect;
            stats->refract += ls.refract;
            stats->launches += ls.launches;
            stats->device_ms = std::max(stats->device_ms, ls.device_ms);  // the slowest device's kernel
        }
    }
    return 0;
}

// Camera::render sharded over `n` device scenes of this process into CALLER host buffers (either may be null): the frame's
// kBandRows-row bands are dealt cyclically, every device renders its bands (in a few launches, so that the copies overlap the
// rendering: render_host) and its own copy engine writes them to their frame positions over its own PCIe link — one host
// thread per device, since render_host returns when its copies have landed.  stats: counters summed, device_ms the slowest
// device's kernel.
int render_host_sharded(DeviceScene* const* scenes, int n, const DCamera& cam, uint8_t* rgba8, double* rgb_f64,
                        LaunchStats* stats, std::string* err) {
    const uint32_t nbands = (cam.vsize + kBandRows - 1) / kBandRows;
    std::vector<int> rcs(n, 0);
    std::vector<std::string> errs(n);
    std::vector<LaunchStats> ls(n);
    const auto render_one = [&](int g) {
        DRows r{};
        r.band_rows = kBandRows;
        r.band_first = (uint32_t)g;
        r.band_stride = (uint32_t)n;
        uint32_t local = 0;
        for (uint32_t b = (uint32_t)g; b < nbands; b += (uint32_t)n) local += std::min(kBandRows, cam.vsize - b * kBandRows);
        r.local_rows = local;
        r.frame_layout = 1;
        r.row_begin = 0;
        r.row_count = local;
        rcs[g] = render_host(scenes[g], cam, r, rgba8, rgb_f64, stats ? &ls[g] : nullptr, &errs[g]);
    };
    std::vector<std::thread> th;
    for (int g = 1; g < n; g++) th.emplace_back(render_one, g);
    render_one(0);
    for (auto& t : th) t.join();
    for (int g = 0; g < n; g++)
        if (rcs[g] != 0) {
            if (err) *err = "device " + std::to_string(g) + ": " + errs[g];
            return rcs[g];
        }
    if (stats) {
        *stats = LaunchStats{};
        for (const LaunchStats& l : ls) {
            stats->primary += l.primary;
            stats->shadow += l.shadow;
            stats->reflect += l.reflect;
            stats->refract += l.refract;
            stats->launches += l.launches;
            stats->device_ms = std::max(stats->device_ms, l.device_ms);
        }
    }
    return 0;
}
int multi_render_host(MultiRenderer* m, const DCamera& cam, uint8_t* rgba8, double* rgb_f64, LaunchStats* stats,
                      std::string* err) {
    return render_host_sharded(m->scene.data(), m->n, cam, rgba8, rgb_f64, stats, err);
}

void* multi_device_frame(const MultiRenderer* m) { return m->d_frame; }
const void* multi_host_frame(const MultiRenderer* m) { return m->h_frame; }
int multi_device_count(const MultiRenderer* m) { return m->n; }

}  // namespace rtc
