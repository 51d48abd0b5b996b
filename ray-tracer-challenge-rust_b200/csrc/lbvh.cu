// lbvh.cu — runs the device mesh build (lbvh.cuh) for every pending mesh of a scene, on the scene's upload stream.
// Launch list per mesh: init, bounds, morton, radix sort of (key, triangle) pairs (cub::DeviceRadixSort — CUDA toolkit
// library code, not part of the render path), hierarchy, fit, emit.  No kernel waits on another thread: `fit` climbs with
// one atomic arrival counter per node and the first thread to arrive simply stops.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstring>
#include <string>

#include "device_scene_impl.cuh"
#include "flat_scene.hpp"
#include "lbvh.cuh"
#include "lbvh_launch.cuh"

namespace rtc {
namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(uint32_t n) { return (n + kThreads - 1) / kThreads; }

__global__ void k_init(lbvh::Work w) {
    if (blockIdx.x == 0 && threadIdx.x == 0) lbvh::gbox_init(w);
}
// per-triangle values are reduced across the warp first: seven atomics per warp instead of seven per thread
__global__ void k_bounds(lbvh::Work w) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v[13];  // 7 of tri_bounds, 6 of the gate fold
    bool ok = true;
    for (int a = 0; a < 13; a++) v[a] = (a < 3 || (a >= 7 && a < 10)) ? ~0ull : 0ull;
    if (k < w.n) {
        lbvh::tri_bounds_values(w, k, v);
        if (w.gate_out) lbvh::gate_values(w, k, v + 7, &ok);
    }
    const int count = w.gate_out ? 13 : 7;
    for (int off = 16; off > 0; off >>= 1)
        for (int a = 0; a < count; a++) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[a], off);
            const bool is_min = a < 3 || (a >= 7 && a < 10);
            v[a] = is_min ? (o < v[a] ? o : v[a]) : (o > v[a] ? o : v[a]);
        }
    if (!ok) *w.error = 1;
    if ((threadIdx.x & 31) == 0) {
        for (int a = 0; a < 3; a++) atomicMin(w.gbox + a, v[a]);
        for (int a = 3; a < 7; a++) atomicMax(w.gbox + a, v[a]);
        if (w.gate_out) {
            for (int a = 0; a < 3; a++) atomicMin(w.gate_acc + a, v[7 + a]);
            for (int a = 3; a < 6; a++) atomicMax(w.gate_acc + a, v[7 + a]);
        }
    }
}
__global__ void k_morton(lbvh::Work w) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < w.n) lbvh::tri_morton(w, k);
}
__global__ void k_hierarchy(lbvh::Work w) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < w.n) lbvh::hierarchy(w, i);
}
__global__ void k_fit(lbvh::Work w) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < w.n) lbvh::fit(w, s);
}
__global__ void k_emit(lbvh::Work w) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < w.n) lbvh::emit(w, i);
}

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

#define LBVH_CUDA(call)                                                       \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) {                                              \
            if (err) *err = cuda_err_string(#call, e_);                       \
            if (scratch) cudaFreeAsync(scratch, st);                          \
            return -3;                                                        \
        }                                                                     \
    } while (0)

size_t lbvh_staging_bytes(const FlatScene& f) {
    return align_up(f.pending_material.size() * sizeof(rtc_triangle_desc)) + align_up(f.pending_material.size() * sizeof(int32_t));
}

int lbvh_build_device(const FlatScene& f, unsigned char* pinned, DBvhNode* d_nodes, DTri* d_tris, DTriAttr* d_attr,
                      DMesh* d_meshes, DGate* d_gates, cudaStream_t st, int* max_depth, bool* would_panic,
                      std::string* err) {
    void* scratch = nullptr;
    if (max_depth) *max_depth = 0;
    if (f.pending.empty()) return 0;
    uint32_t nmax = 0;
    for (const PendingMesh& p : f.pending) nmax = std::max(nmax, p.n);
    const size_t total_in = f.pending_material.size();

    size_t cub_bytes = 0;
    LBVH_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                              (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)nmax, 0, 63, st));
    // one stream-ordered allocation: inputs of every pending mesh, then scratch sized for the largest (meshes build one
    // after another on the stream and reuse it)
    size_t at = 0;
    auto take = [&](size_t bytes) {
        const size_t o = at;
        at += align_up(bytes);
        return o;
    };
    const size_t o_tri = take(total_in * sizeof(rtc_triangle_desc)), o_mat = take(total_in * sizeof(int32_t));
    const size_t o_gbox = take(16 * sizeof(unsigned long long)), o_depth = take(2 * sizeof(int32_t));  // + gate_acc; + error
    const size_t o_keys0 = take(nmax * 8), o_keys1 = take(nmax * 8), o_ord0 = take(nmax * 4), o_ord1 = take(nmax * 4);
    const size_t o_left = take(nmax * 4), o_right = take(nmax * 4), o_parent = take(nmax * 4), o_first = take(nmax * 4),
                 o_last = take(nmax * 4), o_leafp = take(nmax * 4), o_arrive = take(nmax * 4);
    const size_t o_box = take((size_t)nmax * 48), o_lbox = take((size_t)nmax * 48), o_cub = take(cub_bytes);
    LBVH_CUDA(cudaMallocAsync(&scratch, at, st));
    unsigned char* base = (unsigned char*)scratch;

    // inputs: pageable vectors -> pinned staging -> device
    const size_t tri_bytes = total_in * sizeof(rtc_triangle_desc), mat_bytes = total_in * sizeof(int32_t);
    for (const PendingMesh& p : f.pending)
        std::memcpy(pinned + p.input_offset * sizeof(rtc_triangle_desc),
                    p.direct ? p.direct : f.pending_tri.data() + p.input_offset, (size_t)p.n * sizeof(rtc_triangle_desc));
    std::memcpy(pinned + align_up(tri_bytes), f.pending_material.data(), mat_bytes);
    LBVH_CUDA(cudaMemcpyAsync(base + o_tri, pinned, tri_bytes, cudaMemcpyHostToDevice, st));
    LBVH_CUDA(cudaMemcpyAsync(base + o_mat, pinned + align_up(tri_bytes), mat_bytes, cudaMemcpyHostToDevice, st));
    LBVH_CUDA(cudaMemsetAsync(base + o_depth, 0, 2 * sizeof(int32_t), st));

    for (const PendingMesh& p : f.pending) {
        lbvh::Work w{};
        w.tri = (const rtc_triangle_desc*)(base + o_tri) + p.input_offset;
        w.material = (const int32_t*)(base + o_mat) + p.input_offset;
        w.n = p.n;
        w.xform = p.xform;
        w.leaf0 = p.leaf0;
        w.tri_base = p.tri_base;
        w.node_base = p.node_base;
        std::memcpy(w.inv_t, p.inv_t, sizeof(w.inv_t));
        w.gbox = (unsigned long long*)(base + o_gbox);
        w.keys = (unsigned long long*)(base + o_keys0);
        w.order = (uint32_t*)(base + o_ord0);
        w.left = (int32_t*)(base + o_left);
        w.right = (int32_t*)(base + o_right);
        w.parent = (int32_t*)(base + o_parent);
        w.first = (int32_t*)(base + o_first);
        w.last = (int32_t*)(base + o_last);
        w.leaf_parent = (int32_t*)(base + o_leafp);
        w.arrive = (uint32_t*)(base + o_arrive);
        w.box = (double*)(base + o_box);
        w.leaf_box = (double*)(base + o_lbox);
        w.depth_max = (int32_t*)(base + o_depth);
        w.error = (int32_t*)(base + o_depth) + 1;
        w.gate_acc = (unsigned long long*)(base + o_gbox) + 8;
        w.gate_out = p.gate_index >= 0 ? d_gates + p.gate_index : nullptr;
        std::memcpy(w.tr, p.transform, sizeof(w.tr));
        w.nodes = d_nodes;
        w.tris = d_tris;
        w.attr = d_attr;
        w.mesh = d_meshes + p.mesh_index;

        LBVH_CUDA(cudaMemsetAsync(w.arrive, 0, (size_t)(p.n - 1) * 4, st));
        LBVH_CUDA(cudaMemsetAsync(d_nodes + p.node_base, 0, (size_t)(p.n - 1) * sizeof(DBvhNode), st));
        k_init<<<1, 32, 0, st>>>(w);
        k_bounds<<<blocks_for(p.n), kThreads, 0, st>>>(w);
        k_morton<<<blocks_for(p.n), kThreads, 0, st>>>(w);
        size_t cb = cub_bytes;
        LBVH_CUDA(cub::DeviceRadixSort::SortPairs(base + o_cub, cb, (const unsigned long long*)(base + o_keys0),
                                                  (unsigned long long*)(base + o_keys1), (const uint32_t*)(base + o_ord0),
                                                  (uint32_t*)(base + o_ord1), (int)p.n, 0, 63, st));
        w.keys = (unsigned long long*)(base + o_keys1);
        w.order = (uint32_t*)(base + o_ord1);
        k_hierarchy<<<blocks_for(p.n - 1), kThreads, 0, st>>>(w);
        k_fit<<<blocks_for(p.n), kThreads, 0, st>>>(w);
        k_emit<<<blocks_for(p.n - 1), kThreads, 0, st>>>(w);
        LBVH_CUDA(cudaGetLastError());
    }
    int32_t result[2] = {0, 0};  // deepest tree, "the reference would panic"
    LBVH_CUDA(cudaMemcpyAsync(result, base + o_depth, sizeof(result), cudaMemcpyDeviceToHost, st));
    LBVH_CUDA(cudaStreamSynchronize(st));
    LBVH_CUDA(cudaFreeAsync(scratch, st));
    if (max_depth) *max_depth = result[0];
    if (would_panic) *would_panic = result[1] != 0;
    return 0;
}

}  // namespace rtc
