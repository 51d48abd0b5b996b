// bvh.hpp — object-space BVH over the triangles of one mesh (host, built once per scene).
//
// The reference has no spatial structure: a Group tests its box and then every child (shape.rs:399-436).  This BVH sits
// BENEATH the reference's exact gate and may only skip triangle tests whose exact result is a miss, so its boxes are
// padded (kPadRel of the mesh's largest coordinate, ~1e9 ulps) and the device test against them is conservative; which
// triangle wins is still decided by the exact Moller-Trumbore arithmetic and the (t, DFS leaf) order.
//
// Binary tree, binned SAH (kBins bins per axis, all three axes), leaves of <= kLeafMax triangles.  Each 64-byte node
// carries the f32 boxes (rounded outward) of BOTH children, so one node fetch decides two subtrees.  Depth is bounded (median splits past
// kSahDepth) so the device's fixed traversal stack (kBvhStackDepth) cannot overflow.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "device_scene.h"

namespace rtc {

constexpr int kLeafMax = 4;
constexpr int kBins = 32;
constexpr int kSahDepth = 24;
constexpr double kPadRel = 1e-7;

struct BvhTri {
    double p[3][3];  // p1, p2, p3
};

namespace detail {

struct Aabb {
    double lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::numeric_limits<double>::infinity();
            hi[a] = -std::numeric_limits<double>::infinity();
        }
    }
    void grow(const Aabb& o) {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], o.lo[a]);
            hi[a] = std::max(hi[a], o.hi[a]);
        }
    }
    void grow(const double* p) {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    double half_area() const {
        double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0.;
        return dx * dy + dy * dz + dz * dx;
    }
};

// the largest float <= x / smallest float >= x, then one more step outward (so the f32 box strictly contains the f64 one)
inline float f32_below(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return std::nextafterf(f, -std::numeric_limits<float>::infinity());
}
inline float f32_above(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return std::nextafterf(f, std::numeric_limits<float>::infinity());
}

struct BvhBuilder {
    const std::vector<BvhTri>& tris;
    std::vector<DBvhNode>& nodes;
    int32_t tri_base;
    std::vector<uint32_t>& order;  // slot -> original triangle
    std::vector<Aabb> boxes;
    std::vector<double> cx, cy, cz;
    double pad = 0.;
    int max_depth = 0;

    struct Sub {
        bool leaf;
        int32_t index, count;
        Aabb box;
    };

    const double* centroid(uint32_t t, int axis) const { return axis == 0 ? &cx[t] : axis == 1 ? &cy[t] : &cz[t]; }

    Sub build(uint32_t begin, uint32_t end, int depth) {
        if (depth > max_depth) max_depth = depth;
        Aabb box, cbox;
        box.reset();
        cbox.reset();
        for (uint32_t i = begin; i < end; i++) {
            box.grow(boxes[order[i]]);
            double c[3] = {cx[order[i]], cy[order[i]], cz[order[i]]};
            cbox.grow(c);
        }
        const uint32_t n = end - begin;
        if (n <= (uint32_t)kLeafMax) return Sub{true, tri_base + (int32_t)begin, (int32_t)n, box};

        uint32_t mid = 0;
        bool have_split = false;
        if (depth < kSahDepth) {
            double best_cost = std::numeric_limits<double>::infinity();
            int best_axis = -1, best_bin = -1;
            for (int axis = 0; axis < 3; axis++) {
                const double lo = cbox.lo[axis], ext = cbox.hi[axis] - cbox.lo[axis];
                if (!(ext > 0.)) continue;
                Aabb bb[kBins];
                uint32_t cnt[kBins];
                for (int b = 0; b < kBins; b++) {
                    bb[b].reset();
                    cnt[b] = 0;
                }
                const double scale = kBins / ext;
                for (uint32_t i = begin; i < end; i++) {
                    int b = (int)((*centroid(order[i], axis) - lo) * scale);
                    b = std::max(0, std::min(kBins - 1, b));
                    bb[b].grow(boxes[order[i]]);
                    cnt[b]++;
                }
                double right_area[kBins];
                uint32_t right_cnt[kBins];
                Aabb acc;
                acc.reset();
                uint32_t c = 0;
                for (int b = kBins - 1; b >= 1; b--) {
                    acc.grow(bb[b]);
                    c += cnt[b];
                    right_area[b] = acc.half_area();
                    right_cnt[b] = c;
                }
                acc.reset();
                c = 0;
                for (int b = 0; b < kBins - 1; b++) {
                    acc.grow(bb[b]);
                    c += cnt[b];
                    if (c == 0 || right_cnt[b + 1] == 0) continue;
                    double cost = acc.half_area() * c + right_area[b + 1] * right_cnt[b + 1];
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = axis;
                        best_bin = b;
                    }
                }
            }
            if (best_axis >= 0) {
                const double lo = cbox.lo[best_axis], ext = cbox.hi[best_axis] - cbox.lo[best_axis];
                const double scale = kBins / ext;
                auto it = std::partition(order.begin() + begin, order.begin() + end, [&](uint32_t t) {
                    int b = (int)((*centroid(t, best_axis) - lo) * scale);
                    b = std::max(0, std::min(kBins - 1, b));
                    return b <= best_bin;
                });
                mid = (uint32_t)(it - order.begin());
                have_split = (mid > begin && mid < end);
            }
        }
        if (!have_split) {  // median split along the widest centroid axis (also the depth-bounded fallback)
            int axis = 0;
            double e0 = cbox.hi[0] - cbox.lo[0], e1 = cbox.hi[1] - cbox.lo[1], e2 = cbox.hi[2] - cbox.lo[2];
            if (e1 > e0 && e1 >= e2) axis = 1;
            else if (e2 > e0 && e2 > e1) axis = 2;
            mid = begin + n / 2;
            std::nth_element(order.begin() + begin, order.begin() + mid, order.begin() + end, [&](uint32_t a, uint32_t b) {
                double ca = *centroid(a, axis), cb = *centroid(b, axis);
                return ca < cb || (ca == cb && a < b);
            });
        }
        const int32_t me = (int32_t)nodes.size();
        nodes.emplace_back();
        Sub l = build(begin, mid, depth + 1);
        Sub r = build(mid, end, depth + 1);
        DBvhNode& nd = nodes[me];
        for (int a = 0; a < 3; a++) {
            nd.lo0[a] = f32_below(l.box.lo[a] - pad); nd.hi0[a] = f32_above(l.box.hi[a] + pad);
            nd.lo1[a] = f32_below(r.box.lo[a] - pad); nd.hi1[a] = f32_above(r.box.hi[a] + pad);
        }
        nd.child0 = l.index; nd.count0 = l.leaf ? l.count : 0;
        nd.child1 = r.index; nd.count1 = r.leaf ? r.count : 0;
        return Sub{false, me, 0, box};
    }
};

}  // namespace detail

// Appends the mesh's nodes to `nodes`; fills `order` (slot -> input triangle; the caller stores triangles in this order
// starting at global slot `tri_base`).  Returns the root node index, or -1 when the mesh is small enough to scan.
inline int32_t build_bvh(const std::vector<BvhTri>& tris, std::vector<DBvhNode>& nodes, int32_t tri_base,
                         std::vector<uint32_t>& order, int* max_depth, double* max_abs_out = nullptr) {
    const uint32_t n = (uint32_t)tris.size();
    order.resize(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    if (max_depth) *max_depth = 0;
    if (max_abs_out) {
        double m = 0.;
        for (uint32_t i = 0; i < n; i++)
            for (int v = 0; v < 3; v++)
                for (int a = 0; a < 3; a++) m = std::max(m, std::fabs(tris[i].p[v][a]));
        *max_abs_out = m;
    }
    if (n <= (uint32_t)kLeafMax) return -1;
    detail::BvhBuilder b{tris, nodes, tri_base, order, {}, {}, {}, {}, 0., 0};
    b.boxes.resize(n);
    b.cx.resize(n);
    b.cy.resize(n);
    b.cz.resize(n);
    double max_abs = 0.;
    for (uint32_t i = 0; i < n; i++) {
        b.boxes[i].reset();
        for (int v = 0; v < 3; v++) {
            b.boxes[i].grow(tris[i].p[v]);
            for (int a = 0; a < 3; a++) max_abs = std::max(max_abs, std::fabs(tris[i].p[v][a]));
        }
        b.cx[i] = 0.5 * (b.boxes[i].lo[0] + b.boxes[i].hi[0]);
        b.cy[i] = 0.5 * (b.boxes[i].lo[1] + b.boxes[i].hi[1]);
        b.cz[i] = 0.5 * (b.boxes[i].lo[2] + b.boxes[i].hi[2]);
    }
    b.pad = kPadRel * std::max(max_abs, std::numeric_limits<double>::min());
    detail::BvhBuilder::Sub root = b.build(0, n, 0);
    if (max_depth) *max_depth = b.max_depth;
    return root.index;
}

}  // namespace rtc
