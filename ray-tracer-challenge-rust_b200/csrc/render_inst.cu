// render_inst.cu — ONE instantiation of the render kernel: compiled once per feature mask with -DRTC_INST_MASK=<mask>
// (render_launch.cuh lists them).  sm_100a, -fmad=false (see rt_core.cuh for why).
//
// Camera::render's `for y { for x { ... } }` (camera.rs:70-76) is ONE kernel launch: a persistent grid (a fixed number
// of CTAs per SM) whose warps pull 8x4-pixel tiles from an atomic work queue, run World::color_at per lane, quantise
// as canvas.rs:61-63 does and store one uchar4 (and optionally the f64 Canvas colour) per pixel.
#ifndef RTC_INST_MASK
#error "compile with -DRTC_INST_MASK=<feature mask>"
#endif
#include <cuda_runtime.h>

// Instantiations without meshes are a short list of leaves: their tables (prims, transforms, gates, program, materials —
// a few KB) are staged in SHARED memory once per persistent CTA and read from there.
#if defined(RTC_TRY_STAGE_SMEM) && ((RTC_INST_MASK & 32) == 0)
#define RTC_STAGE_SMEM 1
#endif
#include "render_launch.cuh"
#include "rt_core.cuh"

namespace rtc {

using namespace core;

namespace {

// kMinBlocks = CTAs per SM the register allocation must allow (6 -> 80 registers, 24 warps/SM: the measured optimum of
// the launch-shape sweeps in profiles/: more warps hide FP64 latency and fetch bubbles, fewer registers spill).
template <int kMinBlocks, int kFeatures>
__global__ void __launch_bounds__(kBlockThreads, kMinBlocks) render_kernel(const __grid_constant__ DScene s_in,
                                                               const __grid_constant__ DCamera cam,
                                                               const __grid_constant__ DRows rows,
                                                               uint32_t* __restrict__ out8, double* __restrict__ out64,
                                                               DQueue* __restrict__ q) {
#if defined(RTC_STAGE_SMEM)
    // stage the small tables (16-byte granules) and retarget the scene's pointers; scenes too large for the buffer keep
    // reading from global memory
    constexpr uint32_t kStageBytes = 16 * 1024;
    __shared__ __align__(16) unsigned char stage[kStageBytes];
    DScene ls = s_in;
    {
        const uint32_t np = (uint32_t)ls.program_count * (uint32_t)sizeof(DProgramNode);
        const uint32_t npr = ls.n_prims * (uint32_t)sizeof(DPrim), nx = ls.n_xforms * (uint32_t)sizeof(DXform);
        const uint32_t ng = ls.n_gates * (uint32_t)sizeof(DGate), nm = ls.n_materials * (uint32_t)sizeof(DMaterial);
        if (np + npr + nx + ng + nm <= kStageBytes) {
            const void* src[5] = {ls.program, ls.prims, ls.xforms, ls.gates, ls.materials};
            const uint32_t len[5] = {np, npr, nx, ng, nm};
            uint32_t off = 0;
            for (int t = 0; t < 5; t++) {
                const int4* g = (const int4*)src[t];
                int4* d = (int4*)(stage + off);
                for (uint32_t i = threadIdx.x; i < len[t] / 16; i += blockDim.x) d[i] = g[i];
                off += len[t];
            }
            __syncthreads();
            off = 0;
            ls.program = (const DProgramNode*)(stage + off); off += np;
            ls.prims = (const DPrim*)(stage + off); off += npr;
            ls.xforms = (const DXform*)(stage + off); off += nx;
            ls.gates = (const DGate*)(stage + off); off += ng;
            ls.materials = (const DMaterial*)(stage + off);
        }
    }
    const DScene& s = ls;
#else
    const DScene& s = s_in;
#endif
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tiles_x = (cam.hsize + kTileW - 1) / kTileW;
    const uint32_t tiles_y = (rows.row_count + kTileH - 1) / kTileH;
    const uint32_t ntiles = tiles_x * tiles_y;
    RayCounters rc;
    Tally tl;
    uint32_t primary = 0;
    // Tile queue with optional guided batches (a warp takes up to kMaxTileBatch consecutive tiles per atomic, shrinking
    // to one as the queue runs out).  MEASURED (profiles/r01g_tile_batch_sweep.json): batches of 4/8/16 are 1.1x-3x
    // SLOWER on every config — neighbouring heavy tiles land on one warp and a warp's time is the sum of its tiles — and
    // the single-address atomic is not a bottleneck (64 800 grabs per 1080p frame, < 13 % of one L2 slice), so the
    // default is one tile per grab.
    const uint32_t nwarps = gridDim.x * (kBlockThreads / 32);
    uint32_t batch = ntiles / (4u * nwarps);
    batch = batch < 1u ? 1u : (batch > kMaxTileBatch ? kMaxTileBatch : batch);
    uint32_t tile = 0, tile_end = 0;
    for (;;) {
        if (tile >= tile_end) {
            unsigned first = 0;
            if (lane == 0) first = atomicAdd(&q->next_tile, batch);
            first = __shfl_sync(0xffffffffu, first, 0);
            if (first >= ntiles) break;
            tile = first;
            tile_end = first + batch < ntiles ? first + batch : ntiles;
            const uint32_t guided = (ntiles - tile_end) / (2u * nwarps);
            batch = guided < 1u ? 1u : (guided > kMaxTileBatch ? kMaxTileBatch : guided);
        }
        const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
        tile++;
        const uint32_t px = tx * kTileW + (lane & (kTileW - 1));
        const uint32_t lrow = rows.row_begin + ty * kTileH + (lane / kTileW);  // row inside this call's compact output
        if (px < cam.hsize && lrow < rows.row_begin + rows.row_count) {
            const uint32_t band = lrow / rows.band_rows;
            const uint32_t py = (rows.band_first + band * rows.band_stride) * rows.band_rows + (lrow % rows.band_rows);
            const Ray ray = ray_for_pixel(cam, px, py);
            primary++;
            const V3 c = color_at<kFeatures>(s, ray, rc, tl);
            const size_t o = (size_t)(rows.frame_layout ? py : lrow) * cam.hsize + px;
            if (out8) out8[o] = quantise(c.x) | (quantise(c.y) << 8) | (quantise(c.z) << 16) | 0xff000000u;
            if (out64) {
                out64[3 * o + 0] = c.x;
                out64[3 * o + 1] = c.y;
                out64[3 * o + 2] = c.z;
            }
        }
    }
    // ray counters: warp reduce, one atomic per warp and counter
    unsigned long long v[4] = {primary, rc.shadow, rc.reflect, rc.refract};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned x = (unsigned)v[k];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        v[k] = x;
    }
    if (lane == 0) {
        if (v[0]) atomicAdd(&q->primary, v[0]);
        if (v[1]) atomicAdd(&q->shadow, v[1]);
        if (v[2]) atomicAdd(&q->reflect, v[2]);
        if (v[3]) atomicAdd(&q->refract, v[3]);
    }
    // the last CTA to finish publishes the launch's counters and leaves the queue zeroed for the next launch (DQueue)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&q->done_ctas, 1u) == gridDim.x - 1) {
            __threadfence();
            q->result[0] = atomicExch(&q->primary, 0ull);
            q->result[1] = atomicExch(&q->shadow, 0ull);
            q->result[2] = atomicExch(&q->reflect, 0ull);
            q->result[3] = atomicExch(&q->refract, 0ull);
            q->next_tile = 0;
            q->done_ctas = 0;
            __threadfence();
        }
    }
}


}  // namespace

#define RTC_CAT2(a, b) a##b
#define RTC_CAT(a, b) RTC_CAT2(a, b)
void RTC_CAT(launch_render_, RTC_INST_MASK)(unsigned grid, cudaStream_t stream, const DScene& s, const DCamera& cam,
                                            const DRows& rows, uint32_t* out8, double* out64, DQueue* q) {
    render_kernel<kBlocksPerSm, RTC_INST_MASK><<<grid, kBlockThreads, 0, stream>>>(s, cam, rows, out8, out64, q);
}

}  // namespace rtc
