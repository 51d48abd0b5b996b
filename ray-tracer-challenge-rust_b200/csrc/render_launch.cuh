// render_launch.cuh — launch shape of render_kernel and the list of its instantiations (private).
//
// render_kernel<features> is compiled once per FEATURE MASK (rt_core.cuh FEAT_*), each in its own translation unit
// (render_inst.cu with -DRTC_INST_MASK=<mask>, built in parallel by build.py), so that the kernel a scene runs carries no
// code the scene cannot reach: the hot loop's footprint is what the instruction cache sees (profiles/r01h: the
// everything-kernel spent 40 % of its warp-state samples waiting for instructions on the pumpkin scene; dropping the
// unused quadric code alone was worth 8-10 %).  The dispatcher picks the FIRST listed mask that covers the scene's.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_scene.h"
#include "device_scene_impl.cuh"

namespace rtc {

// launch shape (tunable at build time: tools/tune_variants.py)
#ifndef RTC_BLOCK_THREADS
#define RTC_BLOCK_THREADS 128
#endif
#ifndef RTC_BLOCKS_PER_SM
#define RTC_BLOCKS_PER_SM 6
#endif
constexpr int kBlockThreads = RTC_BLOCK_THREADS;
constexpr int kBlocksPerSm = RTC_BLOCKS_PER_SM;  // 6 -> 80 registers, 24 warps/SM (profiles/r01d, r02c, r02zh)
// ... except the mesh + plane kernels (cow & teddy, flat and smooth): 7 -> 72 registers, 28 warps/SM is 4.9 % faster there.
// Everywhere else 7 is within 1 % and multiplies the local-memory DRAM traffic (table 11.6 -> 52.8 MB, pumpkin 1.0 -> 2.6 GB
// per frame: profiles/r02zh_variants.json, the first r02z ncu captures).
// The mesh + refraction kernel (pumpkin) runs 5 -> 96 registers, 20 warps/SM: its frames hold ~300 bytes of live state per
// thread across every scene walk, and with 6 CTAs per SM that working set no longer stays in L2 — 1.00 GB of DRAM writes per 8K
// frame against 0.24 GB with 5 (the frame itself is 0.13 GB) for 0.6 % of frame time (10.04 -> 10.10 ms;
// profiles/r02zk_dram_variants.json; the size of the traversal stack ALLOCATION, 48 / 32 / 24 entries, changes nothing).
constexpr int blocks_per_sm_for(int mask) {
    return (mask == 98 || mask == 2146) ? kBlocksPerSm + 1 : mask == 226 ? kBlocksPerSm - 1 : kBlocksPerSm;
}
#ifndef RTC_TILE_W
#define RTC_TILE_W 8
#endif
constexpr int kTileW = RTC_TILE_W, kTileH = 32 / RTC_TILE_W;  // one warp = one tile (8 x 4: profiles/r02zi_variants.json)
static_assert(kTileW * kTileH == 32 && (kTileW & (kTileW - 1)) == 0, "a tile is one warp");

// mask, in dispatch order (smallest first):  clustered cubes + refraction (table) | spheres + cylinders + groups (hexagon)
// | mesh + groups (teapot) | + plane (cow & teddy) | + refraction (pumpkin) | every primitive kind, clusters (lists and
// trees), no meshes | smooth-triangle meshes + plane (cow & teddy with vertex normals) |
// everything | everything with the general-depth integrator (a RECURSION_LIMIT other than the reference's 5)
#define RTC_RENDER_INSTANCES(X) X(388) X(73) X(96) X(98) X(226) X(1503) X(2146) X(3583) X(4095)

using RenderLaunchFn = void (*)(unsigned grid, cudaStream_t stream, const DScene& s, const DCamera& cam, const DRows& rows,
                                uint32_t* out8, double* out64, DQueue* q);
#define RTC_DECLARE_LAUNCH(mask) \
    void launch_render_##mask(unsigned, cudaStream_t, const DScene&, const DCamera&, const DRows&, uint32_t*, double*, DQueue*);
RTC_RENDER_INSTANCES(RTC_DECLARE_LAUNCH)
#undef RTC_DECLARE_LAUNCH

}  // namespace rtc
