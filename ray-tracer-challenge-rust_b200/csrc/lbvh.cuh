// lbvh.cuh — the mesh build on the device: a run of sibling triangles -> Morton order -> Karras
// hierarchy -> fitted f64 boxes -> the same DBvhNode / DTri / DTriAttr tables the host builder (bvh.hpp, flatten.hpp)
// writes.  The tree differs from the host's binned-SAH tree (a linear BVH trades traversal quality for a build that takes
// a few launches instead of milliseconds of host time); pixels do not: a BVH only decides WHICH exact triangle tests run,
// every box is the exact f64 box of its triangles padded and rounded outward as in bvh.hpp, and hits are still chosen by
// (t, DFS leaf index).
//
// Each step is a plain function of a thread index so that the same code runs as CUDA kernels (lbvh.cu) and, for the CPU
// tests, as loops (tests/hostsim — test infrastructure only):
//   1. tri_bounds   per triangle: box and centroid; mesh-wide centroid box and max |coordinate| (atomic min/max);
//      gate_fold    and, when the mesh is all its group holds, the group's gate box exactly as Bounds::new folds it
//   2. tri_morton   per triangle: 63-bit Morton key of the centroid (21 bits per axis)
//      sort         (key, triangle) pairs by key — cub::DeviceRadixSort on the device, std::stable_sort in the simulation
//   3. hierarchy    per inner node (n-1 of them): range and split by longest common key prefix (Karras 2012), equal keys
//                   told apart by their sorted position
//   4. fit          per sorted slot: writes the slot's DTri / DTriAttr (world normal: shape.rs:509-518, same f64
//                   operations as flatten.hpp), then climbs: the second thread to reach a node merges its children
//   5. emit         per inner node covering more than kLeafMax triangles: one DBvhNode; a child covering <= kLeafMax
//                   becomes a leaf run (sorted slots of a subtree are contiguous)
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "../../include/rtc.h"
#include "device_scene.h"

#if defined(__CUDACC__)
#define LBVH_HD __host__ __device__ inline
#else
#define LBVH_HD inline
#endif

#ifndef RTC_BVH_LEAF_MAX
#define RTC_BVH_LEAF_MAX 4
#endif

namespace rtc {
namespace lbvh {

constexpr int kLeafRun = RTC_BVH_LEAF_MAX;
constexpr double kPad = 1e-7;  // bvh.hpp kPadRel

struct Work {
    // input: the mesh's triangles in input order
    const rtc_triangle_desc* tri;
    const int32_t* material;
    uint32_t n;
    int32_t xform, leaf0;
    int32_t tri_base, node_base;  // where this mesh's slots / nodes start in the scene tables
    double inv_t[16];             // transpose of the mesh transform's inverse (shape.rs:216)
    double tr[16];                // the triangles' transform (gate fold)
    // scratch
    unsigned long long* gbox;  // [7] order-preserving encodings: centroid min xyz, centroid max xyz, max |coordinate|
    unsigned long long* keys;  // n Morton keys (sorted with `order`)
    uint32_t* order;           // n: slot -> input triangle
    int32_t *left, *right, *parent, *first, *last;  // n-1 inner nodes; a child >= 0 is an inner node, < 0 is ~slot
    int32_t* leaf_parent;                           // n
    uint32_t* arrive;                               // n-1, zeroed
    double* box;                                    // (n-1) x 6: lo xyz, hi xyz
    double* leaf_box;                               // n x 6, by slot
    int32_t* depth_max;                             // 1, zeroed
    unsigned long long* gate_acc;                   // [6] ordered bits of the group box being folded: lo xyz, hi xyz
    int32_t* error;                                 // 1, zeroed; set when the reference would panic (bounds.rs:143)
    // output (scene tables)
    DBvhNode* nodes;  // n-1 entries at node_base, zeroed beforehand
    DTri* tris;       // n entries at tri_base
    DTriAttr* attr;
    DMesh* mesh;
    DGate* gate_out;  // null: the mesh's group box was folded on the host
};

// ---- atomics: CUDA on the device, plain on the (single-threaded) simulation ------------------------------------------
LBVH_HD unsigned long long order_bits(double v) {  // monotone map double -> u64 (no NaNs reach it)
    unsigned long long u;
    memcpy(&u, &v, 8);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
LBVH_HD double order_value(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double v;
    memcpy(&v, &u, 8);
    return v;
}
LBVH_HD void atomic_min_u64(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
LBVH_HD void atomic_max_u64(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
LBVH_HD uint32_t atomic_inc_u32(uint32_t* p) {
#if defined(__CUDA_ARCH__)
    __threadfence();  // this thread's box is visible before its arrival is
    const uint32_t old = atomicAdd(p, 1u);
    __threadfence();  // and the sibling's box is read after its arrival was seen
    return old;
#else
    return (*p)++;
#endif
}
LBVH_HD void atomic_max_i32(int32_t* p, int32_t v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
LBVH_HD int clz64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

LBVH_HD double min3(double a, double b, double c) { return fmin(a, fmin(b, c)); }
LBVH_HD double max3(double a, double b, double c) { return fmax(a, fmax(b, c)); }

LBVH_HD void tri_box(const rtc_triangle_desc& t, double lo[3], double hi[3]) {
    for (int a = 0; a < 3; a++) {
        lo[a] = min3(t.p1[a], t.p2[a], t.p3[a]);
        hi[a] = max3(t.p1[a], t.p2[a], t.p3[a]);
    }
}

// ---- 1 ---------------------------------------------------------------------------------------------------------------
LBVH_HD void gbox_init(const Work& w) {
    for (int a = 0; a < 3; a++) {
        w.gbox[a] = ~0ull;
        w.gbox[3 + a] = 0ull;
    }
    w.gbox[6] = 0ull;
    for (int a = 0; a < 6; a++) w.gate_acc[a] = order_bits(0.0);  // Bounds::new seeds a group's box with the origin
}
// The gate box of the mesh's group (bounds.rs:50-151), this triangle's share: its origin-seeded box (bounds.rs:126-137),
// the eight corners through the triangle's transform with the sums of matrix.rs:207-227 (left to right, w row included),
// folded with f64::min / f64::max.  v[0..2] minima, v[3..5] maxima as ordered bits; *ok = false where Bounds::add would
// panic (w != 1: a non-finite coordinate).
LBVH_HD void gate_values(const Work& w, uint32_t k, unsigned long long v[6], bool* ok) {
    double lo[3], hi[3];
    tri_box(w.tri[k], lo, hi);
    for (int a = 0; a < 3; a++) {
        lo[a] = fmin(lo[a], 0.0);
        hi[a] = fmax(hi[a], 0.0);
    }
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int c = 0; c < 8; c++) {
        const double x = (c & 4) ? hi[0] : lo[0], y = (c & 2) ? hi[1] : lo[1], z = (c & 1) ? hi[2] : lo[2];
        const double wv = w.tr[12] * x + w.tr[13] * y + w.tr[14] * z + w.tr[15] * 1.0;
        if (!(wv == 1.0)) *ok = false;
        for (int r = 0; r < 3; r++) {
            const double p = w.tr[4 * r] * x + w.tr[4 * r + 1] * y + w.tr[4 * r + 2] * z + w.tr[4 * r + 3] * 1.0;
            mn[r] = fmin(mn[r], p);
            mx[r] = fmax(mx[r], p);
        }
    }
    for (int r = 0; r < 3; r++) {
        v[r] = mn[r] == mn[r] ? order_bits(mn[r]) : ~0ull;
        v[3 + r] = mx[r] == mx[r] ? order_bits(mx[r]) : 0ull;
    }
}
LBVH_HD void gate_fold(const Work& w, uint32_t k) {
    unsigned long long v[6];
    bool ok = true;
    gate_values(w, k, v, &ok);
    for (int a = 0; a < 3; a++) atomic_min_u64(w.gate_acc + a, v[a]);
    for (int a = 3; a < 6; a++) atomic_max_u64(w.gate_acc + a, v[a]);
    if (!ok) *w.error = 1;
}
// v[0..2] centroid (for the min), v[3..5] centroid (for the max), v[6] max |coordinate|, as ordered bits
LBVH_HD void tri_bounds_values(const Work& w, uint32_t k, unsigned long long v[7]) {
    double lo[3], hi[3];
    tri_box(w.tri[k], lo, hi);
    double m = 0.;
    for (int a = 0; a < 3; a++) {
        const double c = 0.5 * (lo[a] + hi[a]);
        v[a] = v[3 + a] = order_bits(c);
        m = fmax(m, fmax(fabs(lo[a]), fabs(hi[a])));
    }
    v[6] = order_bits(m);
}
LBVH_HD void tri_bounds(const Work& w, uint32_t k) {
    unsigned long long v[7];
    tri_bounds_values(w, k, v);
    for (int a = 0; a < 3; a++) atomic_min_u64(w.gbox + a, v[a]);
    for (int a = 3; a < 7; a++) atomic_max_u64(w.gbox + a, v[a]);
}

// ---- 2 ---------------------------------------------------------------------------------------------------------------
LBVH_HD unsigned long long spread21(unsigned long long x) {  // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
LBVH_HD void tri_morton(const Work& w, uint32_t k) {
    double lo[3], hi[3];
    tri_box(w.tri[k], lo, hi);
    unsigned long long key = 0;
    for (int a = 0; a < 3; a++) {
        const double c = 0.5 * (lo[a] + hi[a]);
        const double cmin = order_value(w.gbox[a]), cmax = order_value(w.gbox[3 + a]);
        const double ext = cmax - cmin;
        double q = ext > 0. ? (c - cmin) / ext * 2097152.0 : 0.;
        q = q < 0. ? 0. : (q > 2097151.0 ? 2097151.0 : q);
        key |= spread21((unsigned long long)q) << (2 - a);
    }
    w.keys[k] = key;
    w.order[k] = k;
}

// ---- 3 ---------------------------------------------------------------------------------------------------------------
// length of the common prefix of sorted keys i and j (64 + position bits when the keys are equal); -1 outside the range
LBVH_HD int delta(const Work& w, int i, int j) {
    if (j < 0 || j >= (int)w.n) return -1;
    const unsigned long long a = w.keys[i], b = w.keys[j];
    if (a != b) return clz64(a ^ b);
    return 64 + clz64((unsigned long long)(uint32_t)i ^ (unsigned long long)(uint32_t)j);
}
LBVH_HD void hierarchy(const Work& w, uint32_t node) {
    const int i = (int)node;
    const int d = (delta(w, i, i + 1) - delta(w, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(w, i, i - d);
    int lmax = 2;
    while (delta(w, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(w, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(w, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (delta(w, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    w.first[i] = lo;
    w.last[i] = hi;
    if (lo == gamma) {
        w.left[i] = ~gamma;
        w.leaf_parent[gamma] = i;
    } else {
        w.left[i] = gamma;
        w.parent[gamma] = i;
    }
    if (hi == gamma + 1) {
        w.right[i] = ~(gamma + 1);
        w.leaf_parent[gamma + 1] = i;
    } else {
        w.right[i] = gamma + 1;
        w.parent[gamma + 1] = i;
    }
    if (i == 0) w.parent[0] = -1;
}

// ---- 4 ---------------------------------------------------------------------------------------------------------------
LBVH_HD void child_box(const Work& w, int32_t child, double b[6]) {
    const double* src = child < 0 ? w.leaf_box + 6 * (size_t)(~child) : w.box + 6 * (size_t)child;
#if defined(__CUDA_ARCH__)
    // written by another thread of the same launch (fit): read through L2, a neighbouring box may sit stale in this SM's L1
    for (int a = 0; a < 6; a++) b[a] = __ldcg(src + a);
#else
    for (int a = 0; a < 6; a++) b[a] = src[a];
#endif
}
LBVH_HD void fit(const Work& w, uint32_t slot) {
    const uint32_t k = w.order[slot];
    const rtc_triangle_desc& t = w.tri[k];
    DTri dt;
    for (int a = 0; a < 3; a++) {
        dt.p1[a] = t.p1[a];
        dt.e1[a] = t.e1[a];
        dt.e2[a] = t.e2[a];
    }
    dt.leaf = w.leaf0 + (int32_t)k;
    dt.cls = -1;  // scenes with value-equal leaves inside a device-built mesh are rebuilt on the host (flatten.hpp)
    w.tris[w.tri_base + slot] = dt;
    // normal_at for a triangle (shape.rs:509-518): invT * normal (w = 0), w = 0, normalize, w = 0, normalize — the sums
    // run left to right and include the w terms, as in flatten.hpp / host_math.hpp
    double v[4];
    for (int r = 0; r < 4; r++)
        v[r] = w.inv_t[4 * r] * t.normal[0] + w.inv_t[4 * r + 1] * t.normal[1] + w.inv_t[4 * r + 2] * t.normal[2] +
               w.inv_t[4 * r + 3] * 0.0;
    for (int pass = 0; pass < 2; pass++) {
        v[3] = 0.;
        const double m = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
        if (m == 0.0) {
            v[0] = v[1] = v[2] = v[3] = 0.;
        } else {
            v[0] = v[0] / m;
            v[1] = v[1] / m;
            v[2] = v[2] / m;
            v[3] = v[3] / m;
        }
    }
    DTriAttr ta;
    ta.normal[0] = v[0];
    ta.normal[1] = v[1];
    ta.normal[2] = v[2];
    ta.material = w.material[k];
    ta.xform = w.xform;
    w.attr[w.tri_base + slot] = ta;

    double lo[3], hi[3];
    tri_box(t, lo, hi);
    double* lb = w.leaf_box + 6 * (size_t)slot;
    for (int a = 0; a < 3; a++) {
        lb[a] = lo[a];
        lb[3 + a] = hi[a];
    }
    int32_t p = w.leaf_parent[slot];
    while (p >= 0) {
        if (atomic_inc_u32(w.arrive + p) == 0) return;  // the sibling subtree is not fitted yet; its thread continues
        double a[6], b[6];
        child_box(w, w.left[p], a);
        child_box(w, w.right[p], b);
        double* out = w.box + 6 * (size_t)p;
        for (int c = 0; c < 3; c++) {
            out[c] = fmin(a[c], b[c]);
            out[3 + c] = fmax(a[3 + c], b[3 + c]);
        }
        p = w.parent[p];
    }
}

// ---- 5 ---------------------------------------------------------------------------------------------------------------
LBVH_HD float f32_below(double x) {
    float f = (float)x;
    if ((double)f > x) f = nextafterf(f, -INFINITY);
    return nextafterf(f, -INFINITY);
}
LBVH_HD float f32_above(double x) {
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return nextafterf(f, INFINITY);
}
LBVH_HD int32_t covered(const Work& w, int32_t child) { return child < 0 ? 1 : w.last[child] - w.first[child] + 1; }
LBVH_HD void emit(const Work& w, uint32_t node) {
    const int32_t i = (int32_t)node;
    if (w.last[i] - w.first[i] + 1 <= kLeafRun) return;  // folded into its parent as a leaf run
    const double max_abs = order_value(w.gbox[6]);
    const double pad = kPad * fmax(max_abs, DBL_MIN);
    DBvhNode nd;
    for (int side = 0; side < 2; side++) {
        const int32_t c = side ? w.right[i] : w.left[i];
        double b[6];
        child_box(w, c, b);
        float* lo = side ? nd.lo1 : nd.lo0;
        float* hi = side ? nd.hi1 : nd.hi0;
        for (int a = 0; a < 3; a++) {
            lo[a] = f32_below(b[a] - pad);
            hi[a] = f32_above(b[3 + a] + pad);
        }
        const int32_t cnt = covered(w, c);
        int32_t child, count;
        if (cnt <= kLeafRun) {
            child = w.tri_base + (c < 0 ? ~c : w.first[c]);
            count = cnt;
        } else {
            child = w.node_base + c;
            count = 0;
        }
        if (side) {
            nd.child1 = child;
            nd.count1 = count;
        } else {
            nd.child0 = child;
            nd.count0 = count;
        }
    }
    w.nodes[w.node_base + i] = nd;
    int depth = 0;
    for (int32_t p = w.parent[i]; p >= 0; p = w.parent[p]) depth++;
    atomic_max_i32(w.depth_max, depth + 2);  // this node's leaf runs sit one level below it; counted from 1 like bvh.hpp
    if (i == 0) {
        w.mesh->root = w.node_base;
        w.mesh->extent = f32_above(max_abs * (1.0 + 2.0 * kPad));
        if (w.gate_out)
            for (int a = 0; a < 3; a++) {
                w.gate_out->lo[a] = order_value(w.gate_acc[a]);
                w.gate_out->hi[a] = order_value(w.gate_acc[3 + a]);
            }
    }
}

}  // namespace lbvh
}  // namespace rtc
