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
    int max_depth = 0;

    static int bin_of(double c, double lo, double scale, int nb) {
        int b = (int)((c - lo) * scale);
        return b < 0 ? 0 : (b >= nb ? nb - 1 : b);
    }

    // chooses the split of [begin, end) and partitions the items; returns mid (begin < mid < end)
    uint32_t split(uint32_t begin, uint32_t end, int depth, const Aabb& cbox) {
        const uint32_t n = end - begin;
        uint32_t mid = 0;
        bool have_split = false;
        // bins scale with the range: evaluating 3 x 32 bins for a handful of triangles costs more than it finds
        const int nb = n >= 512 ? kBins : (n >= 64 ? kBins / 2 : kBins / 4);
        if (depth < kSahDepth && n > 8) {
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
        if (n <= (uint32_t)kLeafMax) return Sub{true, (int32_t)begin, (int32_t)n, box};
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
// The top kParallelDepth levels are split on the calling thread; the subtrees below them are built concurrently, each into
// its own node array, and spliced in a fixed order — the result does not depend on thread timing.
constexpr int kParallelDepth = 3;
constexpr uint32_t kParallelMin = 2048;

inline int32_t build_bvh(const std::vector<BvhTri>& tris, std::vector<DBvhNode>& nodes, int32_t tri_base,
                         std::vector<uint32_t>& order, int* max_depth, double* max_abs_out = nullptr) {
    using namespace detail;
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

    // phase 1: split the top levels, remembering the open ranges
    struct Open {
        uint32_t begin, end;
        int depth;
        int32_t parent;  // node (in `top`) whose child slot this range fills, -1 for the root
        int side;
    };
    std::vector<DBvhNode> top;
    std::vector<Open> open, pending{{0, n, 0, -1, 0}};
    BvhBuilder tb{items, top, pad};
    int depth_seen = 0;
    while (!pending.empty()) {
        Open o = pending.back();
        pending.pop_back();
        const uint32_t cnt = o.end - o.begin;
        if (o.depth >= kParallelDepth || cnt < kParallelMin) {
            open.push_back(o);
            continue;
        }
        Aabb box, cbox;
        tb.range_boxes(o.begin, o.end, box, cbox);
        const uint32_t mid = tb.split(o.begin, o.end, o.depth, cbox);
        const int32_t me = (int32_t)top.size();
        top.emplace_back();
        std::memset(&top[me], 0, sizeof(DBvhNode));
        if (o.parent >= 0) (o.side ? top[o.parent].child1 : top[o.parent].child0) = me;  // inner child: count stays 0
        pending.push_back(Open{mid, o.end, o.depth + 1, me, 1});
        pending.push_back(Open{o.begin, mid, o.depth + 1, me, 0});
        depth_seen = std::max(depth_seen, o.depth);
    }
    // phase 2: the open ranges, concurrently, each into its own node array
    struct Done {
        std::vector<DBvhNode> nodes;
        Sub sub;
        int depth;
    };
    std::vector<Done> done(open.size());
    auto work = [&](size_t k) {
        BvhBuilder b{items, done[k].nodes, pad};
        done[k].sub = b.build(open[k].begin, open[k].end, open[k].depth);
        done[k].depth = b.max_depth;
    };
    if (open.size() > 1) {
        std::vector<std::thread> th;
        for (size_t k = 1; k < open.size(); k++) th.emplace_back(work, k);
        work(0);
        for (auto& t : th) t.join();
    } else {
        work(0);
    }
    // phase 3: splice.  Global node index = base of its array; leaf slots are item positions + tri_base.
    const int32_t top_base = (int32_t)nodes.size();
    std::vector<int32_t> base(open.size());
    int32_t at = top_base + (int32_t)top.size();
    for (size_t k = 0; k < open.size(); k++) {
        base[k] = at;
        at += (int32_t)done[k].nodes.size();
        depth_seen = std::max(depth_seen, done[k].depth);
    }
    // boxes of the top nodes' children, bottom-up: children were created after their parents, so walk backwards
    std::vector<Aabb> top_box(top.size());
    std::vector<Sub> open_sub(open.size());
    for (size_t k = 0; k < open.size(); k++) {
        Sub s = done[k].sub;
        if (s.leaf) s.index += tri_base;
        else s.index += base[k];
        open_sub[k] = s;
    }
    // child links of top nodes: inner links were stored as top-local indices; open ranges fill the rest
    std::vector<Sub> child[2];
    child[0].resize(top.size());
    child[1].resize(top.size());
    std::vector<char> have[2];
    have[0].assign(top.size(), 0);
    have[1].assign(top.size(), 0);
    for (size_t k = 0; k < open.size(); k++)
        if (open[k].parent >= 0) {
            child[open[k].side][open[k].parent] = open_sub[k];
            have[open[k].side][open[k].parent] = 1;
        }
    for (int32_t t = (int32_t)top.size() - 1; t >= 0; t--) {
        for (int side = 0; side < 2; side++)
            if (!have[side][t]) {  // an inner top node
                const int32_t c = side ? top[t].child1 : top[t].child0;
                child[side][t] = Sub{false, top_base + c, 0, top_box[c]};
            }
        top_box[t] = child[0][t].box;
        top_box[t].grow(child[1][t].box);
        tb.fill(top[t], child[0][t], child[1][t]);
    }
    nodes.insert(nodes.end(), top.begin(), top.end());
    for (size_t k = 0; k < open.size(); k++) {
        for (DBvhNode nd : done[k].nodes) {
            if (nd.count0 > 0) nd.child0 += tri_base; else nd.child0 += base[k];
            if (nd.count1 > 0) nd.child1 += tri_base; else nd.child1 += base[k];
            nodes.push_back(nd);
        }
    }
    for (uint32_t i = 0; i < n; i++) order[i] = items[i].id;
    if (max_depth) *max_depth = depth_seen + 1;
    if (top.empty()) return open_sub[0].leaf ? -1 : open_sub[0].index;
    return top_base;
}

}  // namespace rtc
