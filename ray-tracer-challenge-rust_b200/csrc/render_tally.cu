// render_tally.cu — the same per-ray program compiled with work tallies (RTC_TALLY): one extra kernel that renders a
// frame without storing it and counts how many of each reference operation the frame needed (gate tests, leaf tests by
// kind, triangle tests by outcome, shaded hits, pattern evaluations, ...).  bench.py multiplies these by the per-unit
// flop costs of SURVEY.md 8(d) to get the frame's ALGORITHMIC f64 flops for the FP64 roofline.  Never timed.
#define RTC_TALLY 1
#define RTC_CORE_NS core_tally
#include <cuda_runtime.h>

#include "device_scene_impl.cuh"
#include "render.cuh"
#include "rt_core.cuh"

namespace rtc {

using namespace core_tally;

namespace {
__global__ void __launch_bounds__(128) tally_kernel(const __grid_constant__ DScene s, const __grid_constant__ DCamera cam,
                                                    const __grid_constant__ DRows rows, unsigned long long* out) {
    Tally tl;
    RayCounters rc;
    const uint64_t npx = (uint64_t)cam.hsize * rows.local_rows;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t px = (uint32_t)(i % cam.hsize), lrow = (uint32_t)(i / cam.hsize);
        const uint32_t band = lrow / rows.band_rows;
        const uint32_t py = (rows.band_first + band * rows.band_stride) * rows.band_rows + (lrow % rows.band_rows);
        V3 c = color_at_any(s, ray_for_pixel(cam, px, py), rc, tl);
        if (c.x == -12345.678) out[T_COUNT] = 1;  // keep the colour computation alive
    }
    for (int k = 0; k < T_COUNT; k++) {
        unsigned long long v = tl.c[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + k, v);
    }
}
}  // namespace

int tally_count() { return T_COUNT; }

int render_tally(DeviceScene* s, const DCamera& cam, const DRows& rows, unsigned long long* counts, std::string* err) {
    std::lock_guard<std::mutex> lk(s->mu());
    DeviceGuard guard_;
    cudaError_t e = cudaSetDevice(s->device);
    unsigned long long* d = nullptr;
    if (e == cudaSuccess) e = cudaMalloc((void**)&d, sizeof(unsigned long long) * (T_COUNT + 1));
    if (e == cudaSuccess) e = cudaMemsetAsync(d, 0, sizeof(unsigned long long) * (T_COUNT + 1), s->stream);
    if (e == cudaSuccess && rows.local_rows && cam.hsize) {
        tally_kernel<<<s->sm_count * 8, 128, 0, s->stream>>>(s->view, cam, rows, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts, d, sizeof(unsigned long long) * T_COUNT, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (d) cudaFree(d);
    if (e != cudaSuccess) {
        if (err) *err = cuda_err_string("render_tally", e);
        return -3;
    }
    return 0;
}

}  // namespace rtc
