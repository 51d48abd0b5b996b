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
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "device_scene.h"

namespace rtc {

#ifndef RTC_BVH_LEAF_MAX
#define RTC_BVH_LEAF_MAX 4
#endif
constexpr int kLeafMax = RTC_BVH_LEAF_MAX;
constexpr int kBins = 32;
constexpr int kSahDepth = 24;
constexpr int kSweepMax = 8;  // ranges of <= kSweepMax triangles are split by an exact SAH sweep
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

// One triangle while building: its box, its centroid and where it came from.  Items are physically partitioned
// (not index lists), so every pass over a range is a linear sweep.
struct Item {
    double lo[3], hi[3], c[3];
    uint32_t id;
    uint32_t pad;
};

struct Sub {
    bool leaf;
    int32_t index, count;  // leaf: first item slot (relative to the builder's range), count; inner: node index
    Aabb box;
};

// Builds the subtree over items[begin, end) into `nodes` (indices local to `nodes`); leaf slots are item positions.
struct BvhBuilder {
    std::vector<Item>& items;
    std::vector<DBvhNode>& nodes;
    double pad;
    int leaf_max = kLeafMax;  // items per leaf (triangles: kLeafMax; clustered primitives: 1, their exact tests are dear)
    int max_depth = 0;

    static int bin_of(double c, double lo, double scale, int nb) {
        int b = (int)((c - lo) * scale);
        return b < 0 ? 0 : (b >= nb ? nb - 1 : b);
    }

    static void grow_item(Aabb& b, const Item& it) {
        for (int d = 0; d < 3; d++) {
            b.lo[d] = std::min(b.lo[d], it.lo[d]);
            b.hi[d] = std::max(b.hi[d], it.hi[d]);
        }
    }
    // insertion sort of a few items by centroid along `axis`, ties by input index (deterministic)
    static void sort_by_centroid(Item* v, uint32_t n, int axis) {
        for (uint32_t i = 1; i < n; i++) {
            const Item x = v[i];
            uint32_t j = i;
            while (j > 0 && (v[j - 1].c[axis] > x.c[axis] || (v[j - 1].c[axis] == x.c[axis] && v[j - 1].id > x.id))) {
                v[j] = v[j - 1];
                j--;
            }
            v[j] = x;
        }
    }

    // chooses the split of [begin, end) and partitions the items; returns mid (begin < mid < end)
    uint32_t split(uint32_t begin, uint32_t end, int depth, const Aabb& cbox) {
        const uint32_t n = end - begin;
        uint32_t mid = 0;
        bool have_split = false;
        // bins scale with the range: evaluating 3 x 32 bins for a handful of triangles costs more than it finds
        const int nb = n >= 512 ? kBins : (n >= 64 ? kBins / 2 : kBins / 4);
        if (depth < kSahDepth && n <= (uint32_t)kSweepMax) {
            // a handful of triangles just above the leaf size: every split position on every axis, exactly (binning
            // cannot tell such few items apart, and a median split here cost 8-12 % more triangle tests per ray)
            Item tmp[kSweepMax];
            double right_area[kSweepMax];
            double best_cost = std::numeric_limits<double>::infinity();
            int best_axis = -1;
            uint32_t best_k = 0;
            for (int a = 0; a < 3; a++) {
                for (uint32_t k = 0; k < n; k++) tmp[k] = items[begin + k];
                sort_by_centroid(tmp, n, a);
                Aabb acc;
                acc.reset();
                for (uint32_t k = n - 1; k >= 1; k--) {
                    grow_item(acc, tmp[k]);
                    right_area[k] = acc.half_area();
                }
                acc.reset();
                for (uint32_t k = 1; k < n; k++) {
                    grow_item(acc, tmp[k - 1]);
                    const double cost = acc.half_area() * k + right_area[k] * (n - k);
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = a;
                        best_k = k;
                    }
                }
            }
            if (best_axis >= 0) {
                sort_by_centroid(&items[begin], n, best_axis);
                mid = begin + best_k;
                have_split = true;
            }
        } else if (depth < kSahDepth && n > 8) {
            // one sweep bins all three axes
            Aabb bb[3][kBins];
            uint32_t cnt[3][kBins];
            double lo[3], scale[3];
            bool usable[3];
            for (int a = 0; a < 3; a++) {
                const double ext = cbox.hi[a] - cbox.lo[a];
                usable[a] = ext > 0.;
                lo[a] = cbox.lo[a];
                scale[a] = usable[a] ? nb / ext : 0.;
                for (int k = 0; k < nb; k++) {
                    bb[a][k].reset();
                    cnt[a][k] = 0;
                }
            }
            for (uint32_t i = begin; i < end; i++) {
                const Item& it = items[i];
                for (int a = 0; a < 3; a++) {
                    if (!usable[a]) continue;
                    const int k = bin_of(it.c[a], lo[a], scale[a], nb);
                    Aabb& q = bb[a][k];
                    for (int d = 0; d < 3; d++) {
                        q.lo[d] = std::min(q.lo[d], it.lo[d]);
                        q.hi[d] = std::max(q.hi[d], it.hi[d]);
                    }
                    cnt[a][k]++;
                }
            }
            double best_cost = std::numeric_limits<double>::infinity();
            int best_axis = -1, best_bin = -1;
            for (int a = 0; a < 3; a++) {
                if (!usable[a]) continue;
                double right_area[kBins];
                uint32_t right_cnt[kBins];
                Aabb acc;
                acc.reset();
                uint32_t c = 0;
                for (int k = nb - 1; k >= 1; k--) {
                    acc.grow(bb[a][k]);
                    c += cnt[a][k];
                    right_area[k] = acc.half_area();
                    right_cnt[k] = c;
                }
                acc.reset();
                c = 0;
                for (int k = 0; k < nb - 1; k++) {
                    acc.grow(bb[a][k]);
                    c += cnt[a][k];
                    if (c == 0 || right_cnt[k + 1] == 0) continue;
                    const double cost = acc.half_area() * c + right_area[k + 1] * right_cnt[k + 1];
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = a;
                        best_bin = k;
                    }
                }
            }
            if (best_axis >= 0) {
                const double l = lo[best_axis], sc = scale[best_axis];
                auto it = std::partition(items.begin() + begin, items.begin() + end,
                                         [&](const Item& t) { return bin_of(t.c[best_axis], l, sc, nb) <= best_bin; });
                mid = (uint32_t)(it - items.begin());
                have_split = (mid > begin && mid < end);
            }
        }
        if (!have_split) {  // median split along the widest centroid axis (also the depth-bounded fallback)
            int axis = 0;
            const double e0 = cbox.hi[0] - cbox.lo[0], e1 = cbox.hi[1] - cbox.lo[1], e2 = cbox.hi[2] - cbox.lo[2];
            if (e1 > e0 && e1 >= e2) axis = 1;
            else if (e2 > e0 && e2 > e1) axis = 2;
            mid = begin + n / 2;
            std::nth_element(items.begin() + begin, items.begin() + mid, items.begin() + end,
                             [&](const Item& x, const Item& y) {
                                 return x.c[axis] < y.c[axis] || (x.c[axis] == y.c[axis] && x.id < y.id);
                             });
        }
        return mid;
    }

    void range_boxes(uint32_t begin, uint32_t end, Aabb& box, Aabb& cbox) const {
        box.reset();
        cbox.reset();
        for (uint32_t i = begin; i < end; i++) {
            const Item& it = items[i];
            for (int d = 0; d < 3; d++) {
                box.lo[d] = std::min(box.lo[d], it.lo[d]);
                box.hi[d] = std::max(box.hi[d], it.hi[d]);
                cbox.lo[d] = std::min(cbox.lo[d], it.c[d]);
                cbox.hi[d] = std::max(cbox.hi[d], it.c[d]);
            }
        }
    }

    void fill(DBvhNode& nd, const Sub& l, const Sub& r) const {
        for (int a = 0; a < 3; a++) {
            nd.lo0[a] = f32_below(l.box.lo[a] - pad); nd.hi0[a] = f32_above(l.box.hi[a] + pad);
            nd.lo1[a] = f32_below(r.box.lo[a] - pad); nd.hi1[a] = f32_above(r.box.hi[a] + pad);
        }
        nd.child0 = l.index; nd.count0 = l.leaf ? l.count : 0;
        nd.child1 = r.index; nd.count1 = r.leaf ? r.count : 0;
    }

    Sub build(uint32_t begin, uint32_t end, int depth) {
        if (depth > max_depth) max_depth = depth;
        Aabb box, cbox;
        range_boxes(begin, end, box, cbox);
        const uint32_t n = end - begin;
        if (n <= (uint32_t)leaf_max) return Sub{true, (int32_t)begin, (int32_t)n, box};
        const uint32_t mid = split(begin, end, depth, cbox);
        const int32_t me = (int32_t)nodes.size();
        nodes.emplace_back();
        const Sub l = build(begin, mid, depth + 1);
        const Sub r = build(mid, end, depth + 1);
        fill(nodes[me], l, r);
        return Sub{false, me, 0, box};
    }
};

}  // namespace detail

// Appends the mesh's nodes to `nodes`; fills `order` (slot -> input triangle; the caller stores triangles in this order
// starting at global slot `tri_base`).  Returns the root node index, or -1 when the mesh is small enough to scan.
// The top kParallelDepth levels fork: a range is split on the thread that owns it, its right half goes to a new thread
// and its left half stays, so level d runs 2^d splits at once and the critical path is n + n/2 + n/4 + the deepest
// subtree.  Every range builds into its own node array and the arrays are spliced parent, left, right — the result does
// not depend on thread timing.
constexpr int kParallelDepth = 3;
constexpr uint32_t kParallelMin = 2048;

namespace detail {

struct Built {
    std::vector<DBvhNode> nodes;  // indices local to this array; leaf slots are item positions
    Sub sub;
    int depth = 0;
};

inline void build_range(std::vector<Item>& items, double pad, int leaf_max, uint32_t begin, uint32_t end, int depth,
                        Built& out) {
    BvhBuilder b{items, out.nodes, pad, leaf_max};
    if (depth >= kParallelDepth || end - begin < kParallelMin) {
        out.sub = b.build(begin, end, depth);
        out.depth = b.max_depth;
        return;
    }
    Aabb box, cbox;
    b.range_boxes(begin, end, box, cbox);
    const uint32_t mid = b.split(begin, end, depth, cbox);
    Built left, right;
    std::thread other([&] { build_range(items, pad, leaf_max, mid, end, depth + 1, right); });
    build_range(items, pad, leaf_max, begin, mid, depth + 1, left);
    other.join();
    out.nodes.reserve(1 + left.nodes.size() + right.nodes.size());
    out.nodes.emplace_back();
    auto append = [&](const Built& c) {
        const int32_t base = (int32_t)out.nodes.size();
        for (DBvhNode nd : c.nodes) {
            if (nd.count0 == 0) nd.child0 += base;
            if (nd.count1 == 0) nd.child1 += base;
            out.nodes.push_back(nd);
        }
        Sub s = c.sub;
        if (!s.leaf) s.index += base;
        return s;
    };
    const Sub l = append(left);
    const Sub r = append(right);
    std::memset(&out.nodes[0], 0, sizeof(DBvhNode));
    b.fill(out.nodes[0], l, r);
    out.sub = Sub{false, 0, 0, box};
    out.depth = std::max(left.depth, right.depth);
}

}  // namespace detail

// The tree over ready-made items (boxes + centroids + ids): appends its nodes to `nodes`, leaf slots count from
// `leaf_base`, order[slot] = id of the item stored there.  Returns the root node index (-1: everything fits one leaf).
inline int32_t build_bvh_items(std::vector<detail::Item>& items, double pad, int leaf_max, std::vector<DBvhNode>& nodes,
                               int32_t leaf_base, std::vector<uint32_t>& order, int* max_depth) {
    using namespace detail;
    const uint32_t n = (uint32_t)items.size();
    order.resize(n);
    Built root;
    build_range(items, pad, leaf_max, 0, n, 0, root);
    // global indices: nodes after the ones already in `nodes`, leaf slots after leaf_base
    const int32_t base = (int32_t)nodes.size();
    nodes.reserve(nodes.size() + root.nodes.size());
    for (DBvhNode nd : root.nodes) {
        nd.child0 += nd.count0 > 0 ? leaf_base : base;
        nd.child1 += nd.count1 > 0 ? leaf_base : base;
        nodes.push_back(nd);
    }
    for (uint32_t i = 0; i < n; i++) order[i] = items[i].id;
    if (max_depth) *max_depth = root.depth + 1;
    return root.sub.leaf ? -1 : base + root.sub.index;
}

inline int32_t build_bvh(const std::vector<BvhTri>& tris, std::vector<DBvhNode>& nodes, int32_t tri_base,
                         std::vector<uint32_t>& order, int* max_depth, double* max_abs_out = nullptr,
                         double* phase_ms = nullptr) {  // phase_ms[3]: items, build, splice
    using namespace detail;
    auto lap_from = std::chrono::steady_clock::now();
    auto lap = [&](int k) {
        const auto now = std::chrono::steady_clock::now();
        if (phase_ms) phase_ms[k] += std::chrono::duration<double, std::milli>(now - lap_from).count();
        lap_from = now;
    };
    const uint32_t n = (uint32_t)tris.size();
    order.resize(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    if (max_depth) *max_depth = 0;
    double max_abs = 0.;
    std::vector<Item> items(n);
    for (uint32_t i = 0; i < n; i++) {
        Item& it = items[i];
        for (int a = 0; a < 3; a++) {
            const double x = tris[i].p[0][a], y = tris[i].p[1][a], z = tris[i].p[2][a];
            it.lo[a] = std::min(x, std::min(y, z));
            it.hi[a] = std::max(x, std::max(y, z));
            it.c[a] = 0.5 * (it.lo[a] + it.hi[a]);
            max_abs = std::max(max_abs, std::max(std::fabs(it.lo[a]), std::fabs(it.hi[a])));
        }
        it.id = i;
        it.pad = 0;
    }
    if (max_abs_out) *max_abs_out = max_abs;
    if (n <= (uint32_t)kLeafMax) return -1;
    const double pad = kPadRel * std::max(max_abs, std::numeric_limits<double>::min());
    lap(0);
    int depth = 0;
    const int32_t root = build_bvh_items(items, pad, kLeafMax, nodes, tri_base, order, &depth);
    if (max_depth) *max_depth = depth;
    lap(1);
    return root;
}

}  // namespace rtc
