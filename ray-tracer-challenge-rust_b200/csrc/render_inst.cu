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

// normalize as its own out-of-line function (three doubles back) instead of a wrapper of normalize_mag (four): 2.4 % faster
// on the cluster kernels, whose register allocation is the tightest, 0.5-3 % slower on the others (profiles/r02zb_variants.json)
#if (RTC_INST_MASK & 256) && !(RTC_INST_MASK & 32) && !defined(RTC_NORMALIZE_SPLIT)
#define RTC_NORMALIZE_SPLIT 1
#endif
#include "render_launch.cuh"
#include "rt_core.cuh"

namespace rtc {

using namespace core;

namespace {

// canvas.rs:24-26 + :61-63: the Canvas colour and its quantised RGBA8.  The frame is written once and never read here:
// streaming stores, so that on its way to HBM it does not push the scene tables and the traversal stacks out of L2.
__device__ __forceinline__ void store_pixel(uint32_t* __restrict__ out8, double* __restrict__ out64, size_t o, const V3& c) {
#if defined(RTC_PLAIN_FRAME_STORES)  // A/B switch (tools/tune_variants.py)
    if (out8) out8[o] = quantise(c.x) | (quantise(c.y) << 8) | (quantise(c.z) << 16) | 0xff000000u;
    if (out64) {
        out64[3 * o + 0] = c.x;
        out64[3 * o + 1] = c.y;
        out64[3 * o + 2] = c.z;
    }
#else
    if (out8) __stcs(out8 + o, quantise(c.x) | (quantise(c.y) << 8) | (quantise(c.z) << 16) | 0xff000000u);
    if (out64) {
        __stcs(out64 + 3 * o + 0, c.x);
        __stcs(out64 + 3 * o + 1, c.y);
        __stcs(out64 + 3 * o + 2, c.z);
    }
#endif
}

// Queue position -> tile.  The tiles of the hot rectangle (DRows.hot_*: where the scene's bounded geometry projects to) come
// first, row by row; then the rows above it, the tiles left and right of it, and the rows below it.  A bijection of
// [0, tiles_x * tiles_y) for any rectangle inside the grid.
__device__ __forceinline__ void tile_of(uint32_t q, uint32_t tiles_x, const DRows& rows, uint32_t& tx, uint32_t& ty) {
    const uint32_t hw = rows.hot_x1 - rows.hot_x0, hh = rows.hot_y1 - rows.hot_y0, hot = hw * hh;
    if (q < hot) {
        tx = rows.hot_x0 + q % hw;
        ty = rows.hot_y0 + q / hw;
        return;
    }
    q -= hot;
    const uint32_t above = rows.hot_y0 * tiles_x;
    if (hot == 0 || q < above) {  // (no rectangle: plain row-major order)
        tx = q % tiles_x;
        ty = q / tiles_x;
        return;
    }
    q -= above;
    const uint32_t side = tiles_x - hw;  // tiles of a rectangle row that are outside the rectangle
    if (q < side * hh) {
        const uint32_t c = q % side;
        tx = c < rows.hot_x0 ? c : c + hw;
        ty = rows.hot_y0 + q / side;
        return;
    }
    q -= side * hh;
    tx = q % tiles_x;
    ty = rows.hot_y1 + q / tiles_x;
}

// kMinBlocks = CTAs per SM the register allocation must allow (6 -> 80 registers, 24 warps/SM: the measured optimum of
// the launch-shape sweeps in profiles/: more warps hide FP64 latency and fetch bubbles, fewer registers spill; 7 for the
// kernels render_launch.cuh names).
template <int kMinBlocks, int kFeatures>
__global__ void __launch_bounds__(kBlockThreads, kMinBlocks) render_kernel(const __grid_constant__ DScene s_in,
                                                               const __grid_constant__ DCamera cam,
                                                               const __grid_constant__ DRows rows,
                                                               uint32_t* __restrict__ out8, double* __restrict__ out64,
                                                               DQueue* __restrict__ q) {
    const DScene& s = s_in;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tiles_x = (cam.hsize + kTileW - 1) / kTileW;
    const uint32_t tiles_y = (rows.row_count + kTileH - 1) / kTileH;
    const uint32_t ntiles = tiles_x * tiles_y;
    RayCounters rc;
    Tally tl;
    uint32_t primary = 0;
#if defined(RTC_STAGE_BVH_TOP)
    if constexpr (kFeatures & FEAT_MESHES) {  // the top of the first mesh's tree into shared memory (rt_core.cuh load_node)
        const int32_t staged = s.n_bvh < RTC_STAGE_BVH_TOP ? s.n_bvh : RTC_STAGE_BVH_TOP;
        for (int32_t k = threadIdx.x; k < staged * 4; k += blockDim.x) g_bvh_top[k] = __ldg((const float4*)s.bvh + k);
        if (threadIdx.x == 0) g_bvh_top_count = staged;
        __syncthreads();
    }
#endif
#if defined(RTC_LANE_REFILL)
    if constexpr (!(kFeatures & FEAT_DEPTH)) {
        // LANE REFILL: a pixel is a PixelTask advanced one scene walk at a time (rt_core.cuh), and the warp's loop is over
        // WALKS, not pixels — a lane whose pixel is finished (a miss after one walk, a matte surface after two) takes the next
        // pixel of the warp's tile at once instead of idling until the slowest pixel of the tile (a mirror or glass hit: six
        // walks) is done.  Idle lanes are found with one ballot and numbered with a population count; the tile counter is
        // warp-uniform, so handing out pixels costs no atomics beyond the one per tile.
        PixelTask t;
        bool active = false;
        uint32_t slot_o = 0;          // where this lane's pixel goes (index into out8 / out64)
        uint32_t tile = 0, taken = 32;  // the warp's current tile and how many of its 32 pixels are handed out
        bool exhausted = false;
        for (;;) {
            unsigned idle = __ballot_sync(0xffffffffu, !active);
            while (idle != 0 && !exhausted) {
                if (taken == 32) {
                    unsigned next = 0;
                    if (lane == 0) next = atomicAdd(&q->next_tile, 1u);
                    tile = __shfl_sync(0xffffffffu, next, 0);
                    taken = 0;
                    if (tile >= ntiles) {
                        exhausted = true;
                        break;
                    }
                }
                const uint32_t avail = 32u - taken;
                const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
                if (!active && rank < avail) {
                    const uint32_t in_tile = taken + rank;
                    uint32_t tx, ty;
            tile_of(tile, tiles_x, rows, tx, ty);
                    const uint32_t px = tx * kTileW + (in_tile & (kTileW - 1));
                    const uint32_t lrow = rows.row_begin + ty * kTileH + (in_tile / kTileW);
                    if (px < cam.hsize && lrow < rows.row_begin + rows.row_count) {
                        const uint32_t band = lrow / rows.band_rows;
                        const uint32_t py = (rows.band_first + band * rows.band_stride) * rows.band_rows + (lrow % rows.band_rows);
                        task_begin(t, ray_for_pixel(cam, px, py));
                        primary++;
                        slot_o = (rows.frame_layout ? py : lrow) * cam.hsize + px;
                        active = true;
                    }
                }
                const uint32_t n_idle = __popc(idle);
                taken += n_idle < avail ? n_idle : avail;
                idle = __ballot_sync(0xffffffffu, !active);
            }
            if (__ballot_sync(0xffffffffu, active) == 0) break;
            if (active) {
                scene_walk<kFeatures>(s, t.ray, t.w, tl);  // the only call site of the walker
                if (task_step<kFeatures>(s, t, rc, tl)) {
                    const V3 c = t.acc;
                    const size_t o = slot_o;
                    store_pixel(out8, out64, o, c);
                    active = false;
                }
            }
        }
    } else
#endif
    {
        // Tile queue: one atomic per tile.  Measured and rejected: guided batches of 4/8/16 tiles per atomic (1.1x-3x slower on
        // every config — neighbouring heavy tiles land on one warp; profiles/r01g_tile_batch_sweep.json) and requesting the
        // NEXT tile before rendering the current one to hide the atomic's round trip (4-8 % slower on every config: the
        // pending result holds a register across the whole tile; profiles/r02c_variants.json).
        for (;;) {
            unsigned tile = 0;
            if (lane == 0) tile = atomicAdd(&q->next_tile, 1u);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= ntiles) break;
            uint32_t tx, ty;
            tile_of(tile, tiles_x, rows, tx, ty);
            const uint32_t px = tx * kTileW + (lane & (kTileW - 1));
            const uint32_t lrow = rows.row_begin + ty * kTileH + (lane / kTileW);  // row inside this call's compact output
            if (px < cam.hsize && lrow < rows.row_begin + rows.row_count) {
                const uint32_t band = lrow / rows.band_rows;
                const uint32_t py = (rows.band_first + band * rows.band_stride) * rows.band_rows + (lrow % rows.band_rows);
                const Ray ray = ray_for_pixel(cam, px, py);
                primary++;
                V3 c;
                if constexpr (kFeatures & FEAT_DEPTH) c = color_at_general<kFeatures & FEAT_ALL>(s, ray, rc, tl);
                else c = color_at<kFeatures>(s, ray, rc, tl);
                const size_t o = (size_t)(rows.frame_layout ? py : lrow) * cam.hsize + px;
                store_pixel(out8, out64, o, c);
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
        if (rows.notify) __threadfence_system();  // this CTA's pixel stores (possibly into a peer GPU's frame) first
        else __threadfence();
        if (atomicAdd(&q->done_ctas, 1u) == gridDim.x - 1) {
            __threadfence();
            q->result[0] = atomicExch(&q->primary, 0ull);
            q->result[1] = atomicExch(&q->shadow, 0ull);
            q->result[2] = atomicExch(&q->reflect, 0ull);
            q->result[3] = atomicExch(&q->refract, 0ull);
            q->next_tile = 0;
            q->done_ctas = 0;
            __threadfence();
            if (rows.notify) {  // every CTA's stores are ordered before its done_ctas increment, which this CTA has seen
                __threadfence_system();
                atomicAdd_system((unsigned int*)rows.notify, 1u);
            }
        }
    }
}


}  // namespace

#define RTC_CAT2(a, b) a##b
#define RTC_CAT(a, b) RTC_CAT2(a, b)
void RTC_CAT(launch_render_, RTC_INST_MASK)(unsigned grid, cudaStream_t stream, const DScene& s, const DCamera& cam,
                                            const DRows& rows, uint32_t* out8, double* out64, DQueue* q) {
    render_kernel<blocks_per_sm_for(RTC_INST_MASK), RTC_INST_MASK><<<grid, kBlockThreads, 0, stream>>>(s, cam, rows, out8, out64, q);
}

}  // namespace rtc
