// probe.cu — single-ray probes of the per-ray program, so the reference's own unit tests (shape.rs:692-1652,
// intersection.rs:203-379, world.rs:200-209) can be asked of the CUDA path through the C ABI (include/rtc.h:
// rtc_intersect, rtc_prepare_computations, rtc_normal_at).  Same device functions as the render kernel (rt_core.cuh,
// every feature compiled in), sm_100a, -fmad=false.  Never on a frame's path.
#include <cuda_runtime.h>

#include <mutex>

#include "../../include/rtc.h"
#include "device_scene_impl.cuh"
#include "render.cuh"
#include "rt_core.cuh"

namespace rtc {

using namespace core;

namespace {

// World::intersect (world.rs:43-54): all intersections of one ray, then the reference's order — its stable sorts leave
// (t ascending; equal t: DFS leaf order; within a leaf: push order).
struct Collect {
    double* t;
    int32_t* leaf;
    uint32_t cap, n;
    __device__ void offer(const double* ts, int cnt, int32_t lf, int32_t, int32_t, int32_t) {
        for (int k = 0; k < cnt; k++) {
            if (n < cap) {
                t[n] = ts[k];
                leaf[n] = lf;
            }
            n++;
        }
    }
};

__global__ void __launch_bounds__(64) intersect_kernel(const __grid_constant__ DScene s, const double* __restrict__ rays,
                                                       uint64_t n, uint32_t cap, double* __restrict__ t_out,
                                                       int32_t* __restrict__ leaf_out, uint32_t* __restrict__ counts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Ray r{v3(rays[6 * i + 0], rays[6 * i + 1], rays[6 * i + 2]), v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
    Collect c{t_out + i * cap, leaf_out + i * cap, cap, 0};
    Tally tl;
    all_hits_walk<FEAT_ALL>(s, r, RTC_INF, c, tl);
    counts[i] = c.n;
    const uint32_t m = c.n < cap ? c.n : cap;
    for (uint32_t a = 1; a < m; a++) {  // stable insertion sort on (t, leaf)
        const double ta = c.t[a];
        const int32_t la = c.leaf[a];
        uint32_t b = a;
        while (b > 0 && (c.t[b - 1] > ta || (c.t[b - 1] == ta && c.leaf[b - 1] > la))) {
            c.t[b] = c.t[b - 1];
            c.leaf[b] = c.leaf[b - 1];
            b--;
        }
        c.t[b] = ta;
        c.leaf[b] = la;
    }
}

// Intersection::hit (intersection.rs:79-83) + prepare_computations (intersection.rs:17-77) + Computations::schlick
// (intersection.rs:107-128) for the hit of each ray.
__global__ void __launch_bounds__(64) prepare_kernel(const __grid_constant__ DScene s, const double* __restrict__ rays,
                                                     uint64_t n, rtc_computations* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Ray r{v3(rays[6 * i + 0], rays[6 * i + 1], rays[6 * i + 2]), v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
    Tally tl;
    Walk w = walk_closest();
    scene_walk<FEAT_ALL>(s, r, w, tl);
    rtc_computations o{};
    o.leaf = -1;
    if (w.type >= 0) {
        o.hit = 1;
        o.leaf = w.leaf;
        o.t = w.upper;
        // the un-flipped normal first: `inside` is normal_at . eyev < 0 (intersection.rs:22-25)
        const V3 point = position(r, w.upper);
        const V3 eyev = -r.d;
        V3 nv = normal_at<FEAT_ALL>(s, w.type, w.index, point, r, tl);
        o.inside = dot(nv, eyev) < 0.0 ? 1 : 0;
        if (o.inside) nv = -nv;
        const V3 reflectv = reflect(r.d, nv);
        const V3 over = point + nv * kEps, under = point - nv * kEps;
        double n1, n2;
        refraction_indices<FEAT_ALL>(s, r, w.upper, w.leaf, w.type, w.index, n1, n2, tl);
        o.n1 = n1;
        o.n2 = n2;
        o.reflectance = schlick(eyev, nv, n1, n2);
        const V3 v[6] = {point, eyev, nv, reflectv, over, under};
        double* dst[6] = {o.point, o.eyev, o.normalv, o.reflectv, o.over_point, o.under_point};
        for (int k = 0; k < 6; k++) {
            dst[k][0] = v[k].x;
            dst[k][1] = v[k].y;
            dst[k][2] = v[k].z;
        }
    }
    out[i] = o;
}

// Shape::normal_at (shape.rs:466-519) of DFS leaf `leaf` at world points
__global__ void __launch_bounds__(64) normal_kernel(const __grid_constant__ DScene s, int32_t type, int32_t index,
                                                    const double* __restrict__ points, uint64_t n, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Tally tl;
    // a smooth triangle's normal does not depend on the point but on the hit's (u, v): points[] holds (u, v, unused) then —
    // the book's normal_at(tri, point, hit) with an intersection_with_uv
    const bool smooth = type == NODE_MESH && s.tri_smooth != nullptr && s.tri_smooth[index].smooth != 0;
    const Ray none{v3(0., 0., 0.), v3(0., 0., 0.)};
    const V3 nv = smooth ? smooth_normal(s, index, points[3 * i], points[3 * i + 1])
                         : normal_at<FEAT_ALL>(s, type, index, v3(points[3 * i], points[3 * i + 1], points[3 * i + 2]), none, tl);
    out[3 * i + 0] = nv.x;
    out[3 * i + 1] = nv.y;
    out[3 * i + 2] = nv.z;
}

// which table entry holds DFS leaf `leaf` (one thread; probes only)
__global__ void find_leaf_kernel(const __grid_constant__ DScene s, uint32_t n_tris, int32_t leaf, int32_t* out) {
    out[0] = out[1] = -1;
    for (uint32_t k = 0; k < s.n_prims; k++)
        if (s.prims[k].leaf == leaf) {
            out[0] = NODE_PRIM;
            out[1] = (int32_t)k;
            return;
        }
    for (uint32_t k = 0; k < n_tris; k++)
        if (s.tris[k].leaf == leaf) {
            out[0] = NODE_MESH;
            out[1] = (int32_t)k;
            return;
        }
}

// div_by(a, shared_divisor(d)) against the compiler's own a / d (rt_core.cuh): operand pairs from a counter-based
// generator — raw 64-bit patterns (every exponent, NaNs, infinities, subnormals), values of ordinary magnitude, and
// specials paired with everything — compared bit for bit.
__device__ unsigned long long mix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256) divisor_selftest_kernel(unsigned long long pairs_per_thread, unsigned long long seed,
                                                               unsigned long long* mismatches) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double specials[12] = {0.0, -0.0, 1.0, -1.0, 2.0, 0.5, 1e-5, RTC_INF, -RTC_INF, __longlong_as_double(0x7ff8000000000000LL),
                                 __longlong_as_double(0x0000000000000001LL), __longlong_as_double(0x7fefffffffffffffLL)};
    unsigned long long bad = 0;
    for (unsigned long long k = 0; k < pairs_per_thread; k++) {
        const unsigned long long h0 = mix64(seed ^ (tid * pairs_per_thread + k) * 3ull);
        const unsigned long long h1 = mix64(h0), h2 = mix64(h1);
        double a, d;
        switch (h2 & 3) {
            case 0:  // raw bit patterns
                a = __longlong_as_double((long long)h0);
                d = __longlong_as_double((long long)h1);
                break;
            case 1:
            case 2: {  // ordinary magnitudes: random mantissa, exponent within 2^-40 .. 2^40
                const unsigned long long ea = 1023ull - 40ull + ((h2 >> 8) % 81ull), ed = 1023ull - 40ull + ((h2 >> 20) % 81ull);
                a = __longlong_as_double((long long)((h0 & 0x800fffffffffffffull) | (ea << 52)));
                d = __longlong_as_double((long long)((h1 & 0x800fffffffffffffull) | (ed << 52)));
                break;
            }
            default:  // a special on one side or both
                a = ((h2 >> 4) & 1) ? specials[(h2 >> 8) % 12] : __longlong_as_double((long long)h0);
                d = ((h2 >> 5) & 1) ? specials[(h2 >> 16) % 12] : __longlong_as_double((long long)h1);
                break;
        }
        const SharedDivisor sd = shared_divisor(d);
        const double q = div_by(a, sd), ref = a / d;
        if (__double_as_longlong(q) != __double_as_longlong(ref)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

struct Scratch {  // device buffers of one probe call, freed on every path out
    void* p[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Scratch() {
        for (void* q : p)
            if (q) cudaFree(q);
    }
};
#define PROBE_CUDA(call)                                   \
    do {                                                   \
        cudaError_t e_ = (call);                           \
        if (e_ != cudaSuccess) {                           \
            if (err) *err = cuda_err_string(#call, e_);    \
            return -3;                                     \
        }                                                  \
    } while (0)
unsigned blocks_for(uint64_t n) { return (unsigned)((n + 63) / 64); }

}  // namespace

int probe_intersect(DeviceScene* s, const double* rays, uint64_t n, uint32_t cap, double* t_out, int32_t* leaf_out,
                    uint32_t* counts, std::string* err) {
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(s->mu());
    DeviceGuard guard_;
    PROBE_CUDA(cudaSetDevice(s->device));
    Scratch d;
    const size_t slots = (size_t)n * (cap ? cap : 1);
    PROBE_CUDA(cudaMalloc(&d.p[0], n * 48));
    PROBE_CUDA(cudaMalloc(&d.p[1], slots * 8));
    PROBE_CUDA(cudaMalloc(&d.p[2], slots * 4));
    PROBE_CUDA(cudaMalloc(&d.p[3], n * 4));
    PROBE_CUDA(cudaMemcpyAsync(d.p[0], rays, n * 48, cudaMemcpyHostToDevice, s->stream));
    intersect_kernel<<<blocks_for(n), 64, 0, s->stream>>>(s->view, (const double*)d.p[0], n, cap, (double*)d.p[1],
                                                          (int32_t*)d.p[2], (uint32_t*)d.p[3]);
    PROBE_CUDA(cudaGetLastError());
    if (cap) {
        PROBE_CUDA(cudaMemcpyAsync(t_out, d.p[1], slots * 8, cudaMemcpyDeviceToHost, s->stream));
        PROBE_CUDA(cudaMemcpyAsync(leaf_out, d.p[2], slots * 4, cudaMemcpyDeviceToHost, s->stream));
    }
    PROBE_CUDA(cudaMemcpyAsync(counts, d.p[3], n * 4, cudaMemcpyDeviceToHost, s->stream));
    PROBE_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int probe_prepare(DeviceScene* s, const double* rays, uint64_t n, rtc_computations* out, std::string* err) {
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(s->mu());
    DeviceGuard guard_;
    PROBE_CUDA(cudaSetDevice(s->device));
    Scratch d;
    PROBE_CUDA(cudaMalloc(&d.p[0], n * 48));
    PROBE_CUDA(cudaMalloc(&d.p[1], n * sizeof(rtc_computations)));
    PROBE_CUDA(cudaMemcpyAsync(d.p[0], rays, n * 48, cudaMemcpyHostToDevice, s->stream));
    prepare_kernel<<<blocks_for(n), 64, 0, s->stream>>>(s->view, (const double*)d.p[0], n, (rtc_computations*)d.p[1]);
    PROBE_CUDA(cudaGetLastError());
    PROBE_CUDA(cudaMemcpyAsync(out, d.p[1], n * sizeof(rtc_computations), cudaMemcpyDeviceToHost, s->stream));
    PROBE_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int divisor_selftest(int device, uint64_t pairs, uint64_t seed, uint64_t* mismatches, std::string* err) {
    DeviceGuard guard_;
    PROBE_CUDA(cudaSetDevice(device));
    Scratch d;
    PROBE_CUDA(cudaMalloc(&d.p[0], 8));
    PROBE_CUDA(cudaMemset(d.p[0], 0, 8));
    const unsigned blocks = 148 * 8, threads = 256;
    const unsigned long long per_thread = (pairs + (uint64_t)blocks * threads - 1) / ((uint64_t)blocks * threads);
    divisor_selftest_kernel<<<blocks, threads>>>(per_thread, seed, (unsigned long long*)d.p[0]);
    PROBE_CUDA(cudaGetLastError());
    PROBE_CUDA(cudaMemcpy(mismatches, d.p[0], 8, cudaMemcpyDeviceToHost));
    return 0;
}

int probe_normal_at(DeviceScene* s, uint64_t n_tris, int32_t leaf, const double* points, uint64_t n, double* out,
                    std::string* err) {
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(s->mu());
    DeviceGuard guard_;
    PROBE_CUDA(cudaSetDevice(s->device));
    Scratch d;
    PROBE_CUDA(cudaMalloc(&d.p[0], n * 24));
    PROBE_CUDA(cudaMalloc(&d.p[1], n * 24));
    PROBE_CUDA(cudaMalloc(&d.p[2], 8));
    find_leaf_kernel<<<1, 1, 0, s->stream>>>(s->view, (uint32_t)n_tris, leaf, (int32_t*)d.p[2]);
    PROBE_CUDA(cudaGetLastError());
    int32_t where[2] = {-1, -1};
    PROBE_CUDA(cudaMemcpyAsync(where, d.p[2], 8, cudaMemcpyDeviceToHost, s->stream));
    PROBE_CUDA(cudaStreamSynchronize(s->stream));
    if (where[0] < 0) {
        if (err) *err = "no such leaf";
        return -1;
    }
    PROBE_CUDA(cudaMemcpyAsync(d.p[0], points, n * 24, cudaMemcpyHostToDevice, s->stream));
    normal_kernel<<<blocks_for(n), 64, 0, s->stream>>>(s->view, where[0], where[1], (const double*)d.p[0], n, (double*)d.p[1]);
    PROBE_CUDA(cudaGetLastError());
    PROBE_CUDA(cudaMemcpyAsync(out, d.p[1], n * 24, cudaMemcpyDeviceToHost, s->stream));
    PROBE_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

}  // namespace rtc
