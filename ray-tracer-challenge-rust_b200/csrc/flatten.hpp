// flatten.hpp — host side of rtc_scene_create: World (as an rtc_scene_desc pre-order walk) -> flat device tables.
//
//   * validates the description (indices, affine matrices, identity group transforms);
//   * computes every Group's gate box exactly as Bounds::new does per ray in the reference (bounds.rs:11-151,
//     shape.rs:401) — origin-seeded, 8 transformed corners per child, f64::min/max folds — once per scene;
//   * flattens the tree in DFS order into the GATE / PRIM / MESH program of device_scene.h, numbering leaves in the
//     order World::intersect would push them (the tie-break order of the reference's stable sorts);
//   * turns each run of sibling triangles that share one transform into a MESH with a padded object-space BVH
//     (bvh.hpp) and precomputes each triangle's world normal (shape.rs:509-518 is point-independent).
//
// Compiled with -ffp-contract=off: gate boxes and triangle normals reach pixels.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtc.h"
#include "bvh.hpp"
#include "device_scene.h"
#include "flat_scene.hpp"
#include "host_math.hpp"

namespace rtc {

struct FlattenError {
    int code;
    std::string message;
};

// internal: a class of value-equal leaves reaches into a device-built mesh (triangle slots unknown on the host) — flatten
// again with the host build
constexpr int kFlattenNeedsHostBuild = -100;

struct FlattenOptions {
    bool device_mesh_build = false;  // leave meshes of >= kDeviceBuildMin triangles to the device build (lbvh.cuh)
#if defined(RTC_NO_CLUSTERS)  // A/B switch (tools/tune_variants.py)
    bool clusters = false;
#else
    bool clusters = true;            // gather bounded sibling leaves into BVH clusters (off: every leaf is a PRIM entry)
#endif
#if defined(RTC_NO_CLUSTER_HEADERS)  // A/B switch: LIST clusters without header entries (a flat list of leaf boxes)
    bool cluster_headers = false;
#else
    bool cluster_headers = true;
#endif
};
constexpr uint32_t kDeviceBuildMin = 256;

namespace detail {

struct XKey {
    unsigned char b[96];
    bool operator<(const XKey& o) const { return std::memcmp(b, o.b, sizeof(b)) < 0; }
};

struct Box4 {
    Vec4 min, max;
};

class Flattener {
  public:
    Flattener(const rtc_scene_desc& d, FlatScene& out, const FlattenOptions& o = {}) : d_(d), out_(out), opts_(o) {}

    void run() {
        check(d_.shape_count == 0 || d_.shapes, "shapes is NULL");
        check(d_.transform_count == 0 || d_.transforms, "transforms is NULL");
        check(d_.material_count == 0 || d_.materials, "materials is NULL");
        check(d_.triangle_count == 0 || d_.triangles, "triangles is NULL");
        PhaseClock clock;
        // subtree extents of the pre-order walk
        end_.assign(d_.shape_count, 0);
        uint32_t pos = 0;
        for (uint32_t r = 0; r < d_.root_count; r++) pos = scan(pos, 0);
        check(pos == d_.shape_count, "shape_count does not match the pre-order walk of root_count objects");
        for (uint32_t i = 0; i < d_.transform_count; i++) check_affine(d_.transforms[i]);
        for (uint32_t i = 0; i < d_.material_count; i++) out_.materials.push_back(material(d_.materials[i]));
        for (int k = 0; k < 3; k++) {
            out_.light_pos[k] = d_.light_position[k];
            out_.light_int[k] = d_.light_intensity[k];
        }
        for (const DMaterial& m : out_.materials)
            if (m.transparency != 0.0) want_classes_ = true;
        // RECURSION_LIMIT (world.rs:11) as a parameter: the budget is spent three units per bounce, and a limit with
        // limit % 3 == 1 arrives at shade_hit with remaining = 0, whose `remaining - 1` underflows (world.rs:68)
        const uint32_t limit = d_.recursion_limit ? d_.recursion_limit : 5u;
        if (limit % 3 == 1) fail(RTC_ERR_PANIC, "attempt to subtract with overflow (src/world.rs:68): RECURSION_LIMIT % 3 == 1");
        if (limit > (uint32_t)kMaxRecursionLimit) fail(RTC_ERR_UNSUPPORTED, "RECURSION_LIMIT above 18");
        out_.recursion_limit = (int32_t)limit;
        if (limit != 5) out_.feature_mask |= 512;
        clock.lap(out_.phase_ms, FlatScene::T_VALIDATE);
        emit_children(0, d_.shape_count);  // World.objects in order (world.rs:46-50)
        out_.leaf_count = next_leaf_;
        for (const DGate& g : out_.gates) hot_add(g.lo, g.hi);
        // device-built meshes take the table entries after the host-built ones
        for (PendingMesh& p : out_.pending) {
            p.tri_base = (int32_t)(out_.tris.size() + out_.device_tris);
            p.node_base = (int32_t)(out_.bvh.size() + out_.device_nodes);
            out_.meshes[p.mesh_index].tri_base = p.tri_base;
            out_.device_tris += p.n;
            out_.device_nodes += p.n - 1;
        }
        for (const DMesh& m : out_.meshes)
            if (m.xform >= 0) out_.feature_mask |= 32;  // a cluster is a DMesh too (xform -1): not a triangle run
        if (!out_.gates.empty()) out_.feature_mask |= 64;
        for (const DMaterial& m : out_.materials)
            if (m.transparency != 0.0) out_.feature_mask |= 128;
        // n1/n2 are only observable with a transparent material in the scene (world.rs:137-139): no classes otherwise
        if (want_classes_) {
            PhaseClock cc;
            build_classes();
            cc.lap(out_.phase_ms, FlatScene::T_VALIDATE);
        }
    }

  private:
    const rtc_scene_desc& d_;
    FlatScene& out_;
    const FlattenOptions opts_;
    std::vector<uint32_t> end_;  // end_[i] = index just past shape i's subtree
    std::map<XKey, int32_t> xform_ids_;
    std::map<uint32_t, Box4> group_bounds_;
    uint32_t next_leaf_ = 0;
    int32_t parent_gate_ = -1;  // gate of the group being emitted, if this node is its ONLY child
    bool only_child_ = false;
    int group_depth_ = 0;        // 0 while emitting World.objects themselves
    int32_t pending_gate_ = -1;  // gate the next device-built mesh run folds (see device_gate_run)
    int32_t gate_node_ = -1;     // program index of the GATE entry enclosing what is being emitted (-1: World level)
    // where each DFS leaf went: its shape in the description, its program entry and its prims[] index / tris[] slot
    // (slot -1: a triangle of a device-built mesh, placed by the device)
    struct LeafSite {
        uint32_t shape;
        int32_t node, slot;
    };
    std::vector<LeafSite> sites_;
    bool want_classes_ = false;  // some material is transparent: n1/n2 are observable (world.rs:137-139)

    [[noreturn]] static void fail(int code, const std::string& m) { throw FlattenError{code, m}; }
    static void check(bool c, const char* m) {
        if (!c) fail(RTC_ERR_INVALID, m);
    }

    static bool is_tri(int32_t kind) { return kind == RTC_TRIANGLE || kind == RTC_SMOOTH_TRIANGLE; }
    bool has_smooth_ = false;

    uint32_t scan(uint32_t i, int depth) {
        check(i < d_.shape_count, "pre-order walk runs past shape_count");
        check(depth < 64, "group nesting deeper than 64");
        const rtc_shape_desc& s = d_.shapes[i];
        check(s.kind >= RTC_SPHERE && s.kind <= RTC_SMOOTH_TRIANGLE, "unknown shape kind");
        check(s.transform >= 0 && (uint32_t)s.transform < d_.transform_count, "transform index out of range");
        uint32_t next = i + 1;
        if (s.kind == RTC_GROUP) {
            check(s.child_count >= 0, "negative child_count");
            for (int32_t c = 0; c < s.child_count; c++) next = scan(next, depth + 1);
        } else {
            check(s.material >= 0 && (uint32_t)s.material < d_.material_count, "material index out of range");
            if (is_tri(s.kind))
                check(s.triangle >= 0 && (uint32_t)s.triangle < d_.triangle_count, "triangle index out of range");
            if (s.kind == RTC_SMOOTH_TRIANGLE) {
                check(d_.vertex_normals != nullptr, "vertex_normals is NULL in a world with a smooth triangle");
                // vertex normals are placed by the host build only (the device build emits flat triangles)
                if (opts_.device_mesh_build) fail(kFlattenNeedsHostBuild, "smooth triangles need the host mesh build");
                has_smooth_ = true;
            }
        }
        end_[i] = next;
        return next;
    }

    static bool is_pm_zero(double v) { return v == 0.0; }
    void check_affine(const rtc_transform_desc& t) {
        const double* a = t.transform;
        const double* b = t.inverse;
        bool ok = is_pm_zero(a[12]) && is_pm_zero(a[13]) && is_pm_zero(a[14]) && a[15] == 1.0 && is_pm_zero(b[12]) &&
                  is_pm_zero(b[13]) && is_pm_zero(b[14]) && b[15] == 1.0;
        if (!ok) fail(RTC_ERR_UNSUPPORTED, "non-affine shape transform (bottom row must be 0 0 0 1)");
    }

    DMaterial material(const rtc_material& m) {
        DMaterial o;
        std::memset(&o, 0, sizeof(o));
        for (int k = 0; k < 3; k++) {
            o.color[k] = m.color[k];
            o.pa[k] = m.pattern_a[k];
            o.pb[k] = m.pattern_b[k];
        }
        o.ambient = m.ambient; o.diffuse = m.diffuse; o.specular = m.specular; o.shininess = m.shininess;
        o.reflective = m.reflective; o.transparency = m.transparency; o.refractive_index = m.refractive_index;
        check(m.pattern_kind >= RTC_PATTERN_NONE && m.pattern_kind <= RTC_PATTERN_TEST, "unknown pattern kind");
        o.pattern_kind = m.pattern_kind;
        if (m.pattern_kind >= 0) {
            const double* b = m.pattern_inverse;
            if (!(is_pm_zero(b[12]) && is_pm_zero(b[13]) && is_pm_zero(b[14]) && b[15] == 1.0))
                fail(RTC_ERR_UNSUPPORTED, "non-affine pattern transform");
            std::memcpy(o.pinv, b, sizeof(double) * 12);
        } else {
            Mat4 id = Mat4::identity();
            std::memcpy(o.pinv, id.m, sizeof(double) * 12);
        }
        return o;
    }

    int32_t xform_id(int32_t transform_index) {
        XKey k;
        std::memcpy(k.b, d_.transforms[transform_index].inverse, sizeof(k.b));
        auto it = xform_ids_.find(k);
        if (it != xform_ids_.end()) return it->second;
        DXform x;
        std::memcpy(x.m, d_.transforms[transform_index].inverse, sizeof(x.m));
        int32_t id = (int32_t)out_.xforms.size();
        out_.xforms.push_back(x);
        xform_ids_.emplace(k, id);
        return id;
    }

    // ---- bounds.rs:11-151 -------------------------------------------------------------------------------------------
    static void add(Box4& b, const Vec4& p) {  // bounds.rs:142-151
        if (!(p.w == 1.0)) fail(RTC_ERR_PANIC, "assertion failed: point.is_point() (src/bounds.rs:143) — a group holds a "
                                               "shape with an unbounded box (uncapped cylinder/cone)");
        b.min.x = min_(b.min.x, p.x); b.min.y = min_(b.min.y, p.y); b.min.z = min_(b.min.z, p.z);
        b.max.x = max_(b.max.x, p.x); b.max.y = max_(b.max.y, p.y); b.max.z = max_(b.max.z, p.z);
    }
    // f64::min / f64::max (a NaN operand loses), inline: libm's fmin/fmax calls were half of a mesh group's box time.
    // On a +0 / -0 tie either may come back (as with f64::min); no comparison a gate makes can tell them apart.
    static double min_(double a, double b) { return a != a ? b : (b < a ? b : a); }
    static double max_(double a, double b) { return a != a ? b : (b > a ? b : a); }
    void fold_child(Box4& out, uint32_t c) {  // bounds.rs:52-124: the child's box, eight corners through its transform
        const Box4 pb = bounds_of(c);
        const Mat4 tr = Mat4::from(d_.transforms[d_.shapes[c].transform].transform);
        add(out, mul(tr, point(pb.min.x, pb.min.y, pb.min.z)));
        add(out, mul(tr, point(pb.min.x, pb.min.y, pb.max.z)));
        add(out, mul(tr, point(pb.min.x, pb.max.y, pb.min.z)));
        add(out, mul(tr, point(pb.min.x, pb.max.y, pb.max.z)));
        add(out, mul(tr, point(pb.max.x, pb.min.y, pb.min.z)));
        add(out, mul(tr, point(pb.max.x, pb.min.y, pb.max.z)));
        add(out, mul(tr, point(pb.max.x, pb.max.y, pb.min.z)));
        add(out, mul(tr, pb.max));
    }
    Box4 bounds_of(uint32_t i) {
        const rtc_shape_desc& s = d_.shapes[i];
        const double inf = std::numeric_limits<double>::infinity();
        switch (s.kind) {
            case RTC_SPHERE:
            case RTC_CUBE: return {point(-1., -1., -1.), point(1., 1., 1.)};
            case RTC_PLANE: return {point(-1., -1., 0.), point(1., 1., 0.)};  // sic, bounds.rs:21-24
            case RTC_CYLINDER:
            case RTC_CONE:
                if (s.capped) return {point(-1., s.minimum, -1.), point(1., s.maximum, 1.)};
                return {point(-1., -inf, -1.), point(1., inf, 1.)};
            case RTC_SMOOTH_TRIANGLE:  // the book bounds a smooth triangle like a triangle
            case RTC_TRIANGLE: {  // bounds.rs:126-137: seeded with the origin
                const rtc_triangle_desc& t = d_.triangles[s.triangle];
                Box4 b{point(0., 0., 0.), point(0., 0., 0.)};
                add(b, point(t.p1[0], t.p1[1], t.p1[2]));
                add(b, point(t.p2[0], t.p2[1], t.p2[2]));
                add(b, point(t.p3[0], t.p3[1], t.p3[2]));
                return b;
            }
            default: {  // group, bounds.rs:50-125 (memoised: a nested group's box is asked for by its parent and by itself)
                auto hit = group_bounds_.find(i);
                if (hit != group_bounds_.end()) return hit->second;
                Box4 out{point(0., 0., 0.), point(0., 0., 0.)};
                const uint32_t first = i + 1, last = end_[i];
                bool folded = false;
                if ((uint32_t)s.child_count >= kParallelMin * 2 && last - first == (uint32_t)s.child_count) {
                    // a long run of leaves (an OBJ mesh): slices fold on their own threads, each from the origin seed the
                    // whole fold starts from, and min/max merge them — the same box.  A slice that would panic leaves
                    // the serial loop below to raise it at the right child.
                    const uint32_t slices = std::min<uint32_t>(8, std::max(1u, std::thread::hardware_concurrency()));
                    std::vector<Box4> part(slices, out);
                    std::vector<char> ok(slices, 1);
                    std::vector<std::thread> th;
                    auto run = [&](uint32_t k) {
                        const uint32_t a = first + (uint64_t)(last - first) * k / slices;
                        const uint32_t b = first + (uint64_t)(last - first) * (k + 1) / slices;
                        try {
                            for (uint32_t c = a; c < b; c++) fold_child(part[k], c);
                        } catch (const FlattenError&) {
                            ok[k] = 0;
                        }
                    };
                    for (uint32_t k = 1; k < slices; k++) th.emplace_back(run, k);
                    run(0);
                    for (auto& t : th) t.join();
                    folded = true;
                    for (uint32_t k = 0; k < slices; k++) folded = folded && ok[k];
                    if (folded)
                        for (const Box4& p : part) {
                            add(out, p.min);
                            add(out, p.max);
                        }
                }
                if (!folded)
                    for (uint32_t c = first; c < last; c = end_[c]) fold_child(out, c);
                group_bounds_.emplace(i, out);
                return out;
            }
        }
    }

    static float f32_above_(double x) {
        float f = (float)x;
        if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
        return std::nextafterf(f, std::numeric_limits<float>::infinity());
    }

    // ---- clusters: a BVH over the world boxes of sibling spheres / cubes / finite cylinders (device_scene.h DMesh) ------
    // Own structure, not reference arithmetic: it may only skip exact leaf tests whose result is "no intersection".
    //   sphere / cube : the unit cube [-1,1]^3 through the leaf's transform (shape.rs:258-319)
    //   cylinder      : radius 1 walls for min < y < max; caps accept x^2 + z^2 <= |y| (shape.rs:584), i.e. radius
    //                   sqrt(|y|) — the box takes the larger of the two; needs finite min and max
    //   cone, plane   : never clustered (the cone's a ~ 0 branch reports an intersection without a y-range check,
    //                   shape.rs:363-368)
    // The cube's check_axis treats |direction| < EPSILON as parallel (shape.rs:593-599) and then reports intersections that
    // have drifted up to EPSILON * t outside the true cube along that axis: a cube's box is padded by EPSILON * (the longest
    // ray the fast path admits) * (its largest column norm).
    static constexpr uint32_t kClusterMin = 5;     // fewer bounded siblings are cheaper to test outright (hexagon: 2 per group)
    static constexpr double kClusterReachR = 8.0;  // fast path: rays that start within 8 R + 10 of the cluster's centre
    static constexpr double kClusterReachAdd = 10.0;
    static constexpr double kClusterMaxScale = 5000.0;
    struct ClusterItem {
        DPrim prim;
        uint32_t shape;
        double lo[3], hi[3];
        double scale;  // cubes: largest column norm of the transform's 3x3; others: 0 (no EPSILON drift)
    };
    // world box of a bounded leaf; false: not bounded (or not finite) — stays a PRIM entry
    bool leaf_world_box(const rtc_shape_desc& s, ClusterItem& it) const {
        double r = 1.0, ylo = -1.0, yhi = 1.0;
        if (s.kind == RTC_SPHERE || s.kind == RTC_CUBE) {
        } else if (s.kind == RTC_CYLINDER && std::isfinite(s.minimum) && std::isfinite(s.maximum)) {
            ylo = std::fmin(s.minimum, s.maximum);
            yhi = std::fmax(s.minimum, s.maximum);
            if (s.capped) r = std::fmax(1.0, std::sqrt(std::fmax(std::fabs(s.minimum), std::fabs(s.maximum))));
        } else {
            return false;
        }
        const rtc_transform_desc& td = d_.transforms[s.transform];
        const Mat4 t = Mat4::from(td.transform);
        for (int a = 0; a < 3; a++) {
            it.lo[a] = 1e300;
            it.hi[a] = -1e300;
        }
        for (int c = 0; c < 8; c++) {
            Vec4 w = mul(t, point((c & 1) ? r : -r, (c & 2) ? yhi : ylo, (c & 4) ? r : -r));
            const double v[3] = {w.x, w.y, w.z};
            for (int a = 0; a < 3; a++) {
                if (!std::isfinite(v[a]) || std::fabs(v[a]) > 1e30) return false;
                it.lo[a] = std::fmin(it.lo[a], v[a]);
                it.hi[a] = std::fmax(it.hi[a], v[a]);
            }
        }
        it.scale = 0.;
        if (s.kind == RTC_CUBE) {
            for (int c = 0; c < 3; c++) {
                const double x = td.transform[c], y = td.transform[4 + c], z = td.transform[8 + c];
                it.scale = std::fmax(it.scale, std::sqrt(x * x + y * y + z * z));
            }
            if (!(it.scale <= kClusterMaxScale)) return false;
        }
        return true;
    }
    uint32_t cluster_candidates(uint32_t begin, uint32_t end) const {
        uint32_t n = 0;
        ClusterItem tmp;
        for (uint32_t c = begin; c < end; c = end_[c])
            if (d_.shapes[c].kind != RTC_GROUP && !is_tri(d_.shapes[c].kind) && leaf_world_box(d_.shapes[c], tmp)) n++;
        return n;
    }
    void push_prim(const DPrim& p) { out_.prims.push_back(p); }
    void hot_add(const double* lo, const double* hi) {
        for (int a = 0; a < 3; a++) {
            if (!(std::isfinite(lo[a]) && std::isfinite(hi[a]))) return;
        }
        for (int a = 0; a < 3; a++) {
            out_.hot_lo[a] = std::fmin(out_.hot_lo[a], lo[a]);
            out_.hot_hi[a] = std::fmax(out_.hot_hi[a], hi[a]);
        }
    }

    // Relabels the nodes out_.bvh[first ..) of one mesh in breadth-first order from `root`: the top levels of the tree — the
    // nodes every ray visits — become the first entries of the mesh's range (contiguous in L1/L2; a kernel that stages the
    // top of the tree in shared memory, RTC_STAGE_BVH_TOP, copies one range).  Returns the new root index (= first).
    int32_t breadth_first(size_t first, int32_t root) {
        const size_t n = out_.bvh.size() - first;
        std::vector<int32_t> order;  // new position -> old index
        order.reserve(n);
        std::vector<int32_t> where(n, -1);  // old index - first -> new position
        order.push_back(root);
        for (size_t head = 0; head < order.size(); head++) {
            const DBvhNode& nd = out_.bvh[order[head]];
            where[order[head] - first] = (int32_t)head;
            if (nd.count0 == 0) order.push_back(nd.child0);
            if (nd.count1 == 0) order.push_back(nd.child1);
        }
        if (order.size() != n) return root;  // (unreachable nodes would be a builder bug: leave the layout alone)
        std::vector<DBvhNode> moved(n);
        for (size_t k = 0; k < n; k++) {
            DBvhNode nd = out_.bvh[order[k]];
            if (nd.count0 == 0) nd.child0 = (int32_t)first + where[nd.child0 - first];
            if (nd.count1 == 0) nd.child1 = (int32_t)first + where[nd.child1 - first];
            moved[k] = nd;
        }
        std::copy(moved.begin(), moved.end(), out_.bvh.begin() + first);
        return (int32_t)first;
    }

    // ---- the skip list of a LIST cluster (device_scene.h DBox32) -------------------------------------------------------
    // A small exact-sweep SAH tree over the leaves' padded boxes, written out in depth-first order; an inner node becomes a
    // HEADER entry only where that lowers the expected number of box tests (surface-area heuristic: a ray that enters a box
    // of area A enters a box of area a inside it with probability a / A).  The root is never a header: the scenes this serves
    // have the camera inside the cluster.
    struct SkipNode {
        double lo[3], hi[3];
        int left = -1, right = -1;  // inner
        int item = -1;              // leaf: index into the cluster's run
        int leaves = 1;
    };
    static double box_area(const double* lo, const double* hi) {
        const double x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return 2.0 * (x * y + y * z + z * x);
    }
    // (Every split position of every axis is tried with its exact cost: boxes of the two sides from prefix / suffix unions —
    // min and max are exact, so these are the boxes a fresh fold over each side would give — and the first strictly cheaper
    // candidate in (axis, cut) order wins.  At most kClusterListMax items: fixed arrays, an insertion sort (stable, like the
    // std::stable_sort this replaced: the order is the same), no allocation.  The table scene's 17-cube cluster took
    // 0.14 ms of every scene creation when each cut re-folded both sides and copied index vectors.)
    static int skip_build(std::vector<SkipNode>& nodes, const std::vector<detail::Item>& items, const int* idx, int n) {
        SkipNode nd;
        for (int a = 0; a < 3; a++) {
            nd.lo[a] = 1e300;
            nd.hi[a] = -1e300;
        }
        for (int k = 0; k < n; k++)
            for (int a = 0; a < 3; a++) {
                nd.lo[a] = std::fmin(nd.lo[a], items[idx[k]].lo[a]);
                nd.hi[a] = std::fmax(nd.hi[a], items[idx[k]].hi[a]);
            }
        nd.leaves = n;
        if (n == 1) {
            nd.item = idx[0];
            nodes.push_back(nd);
            return (int)nodes.size() - 1;
        }
        constexpr int kMax = kClusterListMax;
        double best = 1e300;
        int best_order[kMax], best_cut = 0;
        for (int axis = 0; axis < 3; axis++) {
            int o[kMax];
            for (int k = 0; k < n; k++) {  // stable insertion sort by the centre's coordinate
                const int v = idx[k];
                int j = k;
                while (j > 0 && items[v].c[axis] < items[o[j - 1]].c[axis]) {
                    o[j] = o[j - 1];
                    j--;
                }
                o[j] = v;
            }
            double slo[kMax + 1][3], shi[kMax + 1][3];  // suffix unions: items o[k .. n)
            for (int a = 0; a < 3; a++) {
                slo[n][a] = 1e300;
                shi[n][a] = -1e300;
            }
            for (int k = n - 1; k >= 0; k--)
                for (int a = 0; a < 3; a++) {
                    slo[k][a] = std::fmin(slo[k + 1][a], items[o[k]].lo[a]);
                    shi[k][a] = std::fmax(shi[k + 1][a], items[o[k]].hi[a]);
                }
            double plo[3] = {1e300, 1e300, 1e300}, phi[3] = {-1e300, -1e300, -1e300};  // prefix union: items o[0 .. cut)
            for (int cut = 1; cut < n; cut++) {
                for (int a = 0; a < 3; a++) {
                    plo[a] = std::fmin(plo[a], items[o[cut - 1]].lo[a]);
                    phi[a] = std::fmax(phi[a], items[o[cut - 1]].hi[a]);
                }
                const double cost = box_area(plo, phi) * (double)cut + box_area(slo[cut], shi[cut]) * (double)(n - cut);
                if (cost < best) {
                    best = cost;
                    best_cut = cut;
                    std::memcpy(best_order, o, sizeof(int) * (size_t)n);
                }
            }
        }
        const int l = skip_build(nodes, items, best_order, best_cut);
        const int r = skip_build(nodes, items, best_order + best_cut, n - best_cut);
        nd.left = l;
        nd.right = r;
        nodes.push_back(nd);
        return (int)nodes.size() - 1;
    }
    // Expected box tests of subtree n when the nearest tested box above it is node `anc` (-1: none — the walk starts here);
    // memoised per (n, anc).  .second: make n a header in that context.
    struct SkipMemo {  // a table over (node, ancestor + 1): at most 63 x 64 entries
        int nn;
        std::vector<std::pair<double, bool>> value;
        std::vector<char> known;
        explicit SkipMemo(int nodes) : nn(nodes), value((size_t)nodes * (nodes + 1)), known((size_t)nodes * (nodes + 1), 0) {}
        size_t at(int n, int anc) const { return (size_t)n * (nn + 1) + (size_t)(anc + 1); }
    };
    static std::pair<double, bool> skip_cost(const std::vector<SkipNode>& nodes, int n, int anc, SkipMemo& memo) {
        const SkipNode& nd = nodes[n];
        if (nd.item >= 0) return {1.0, false};
        const size_t key = memo.at(n, anc);
        if (memo.known[key]) return memo.value[key];
        const double plain = skip_cost(nodes, nd.left, anc, memo).first + skip_cost(nodes, nd.right, anc, memo).first;
        const double area = box_area(nd.lo, nd.hi);
        const double within = anc >= 0 ? box_area(nodes[anc].lo, nodes[anc].hi) : 0.0;
        const double p = within > 0.0 ? std::fmin(1.0, area / within) : 1.0;
        const double head = 1.0 + p * (skip_cost(nodes, nd.left, n, memo).first + skip_cost(nodes, nd.right, n, memo).first);
        const bool use = anc >= 0 && nd.leaves >= 3 && head < plain;  // anc < 0: the root is never a header
        const std::pair<double, bool> r{use ? head : plain, use};
        memo.value[key] = r;
        memo.known[key] = 1;
        return r;
    }
    void skip_emit(const std::vector<SkipNode>& nodes, int n, int anc, SkipMemo& memo, bool headers, double pad,
                   int32_t prim_base) {
        const SkipNode& nd = nodes[n];
        const bool head = nd.item < 0 && headers && skip_cost(nodes, n, anc, memo).second;
        if (nd.item < 0 && !head) {
            skip_emit(nodes, nd.left, anc, memo, headers, pad, prim_base);
            skip_emit(nodes, nd.right, anc, memo, headers, pad, prim_base);
            return;
        }
        DBox32 b;
        std::memset(&b, 0, sizeof(b));
        for (int a = 0; a < 3; a++) {
            b.lo[a] = detail::f32_below(nd.lo[a] - pad);
            b.hi[a] = detail::f32_above(nd.hi[a] + pad);
        }
        b.skip = -1;
        b.prim = nd.item >= 0 ? prim_base + nd.item : -1;
        const size_t at = out_.cluster_entries.size();
        out_.cluster_entries.push_back(b);
        if (nd.item >= 0) return;
        skip_emit(nodes, nd.left, n, memo, headers, pad, prim_base);
        skip_emit(nodes, nd.right, n, memo, headers, pad, prim_base);
        out_.cluster_entries[at].skip = (int32_t)(out_.cluster_entries.size() - at - 1);
    }
    void set_site(uint32_t leaf, const LeafSite& st) {
        if (!want_classes_) return;
        if (sites_.size() <= leaf) sites_.resize(leaf + 1);
        sites_[leaf] = st;
    }
    void emit_cluster(std::vector<ClusterItem>& run) {
        const uint32_t n = (uint32_t)run.size();
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, smax = 0.;
        for (const ClusterItem& it : run) {
            for (int a = 0; a < 3; a++) {
                lo[a] = std::fmin(lo[a], it.lo[a]);
                hi[a] = std::fmax(hi[a], it.hi[a]);
            }
            smax = std::fmax(smax, it.scale);
        }
        hot_add(lo, hi);
        double centre[3], r2 = 0.;
        for (int a = 0; a < 3; a++) {
            centre[a] = 0.5 * (lo[a] + hi[a]);
            r2 += 0.25 * (hi[a] - lo[a]) * (hi[a] - lo[a]);
        }
        const double radius = std::sqrt(r2);
        const double reach = kClusterReachR * radius + kClusterReachAdd;
        // the longest fast-path ray to any point of a (padded) box: origin within `reach`, direction length within 1 %, and
        // the pad itself (at most sqrt(3) * EPSILON * scale <= 0.09 of the length, kClusterMaxScale)
        const double longest = 1.25 * (reach + radius);
        std::vector<detail::Item> items(n);
        double max_abs = 0.;
        for (uint32_t k = 0; k < n; k++) {
            const ClusterItem& ci = run[k];
            const double drift = 1.0001e-5 * longest * ci.scale;
            detail::Item& it = items[k];
            for (int a = 0; a < 3; a++) {
                it.lo[a] = ci.lo[a] - drift;
                it.hi[a] = ci.hi[a] + drift;
                it.c[a] = 0.5 * (it.lo[a] + it.hi[a]);
                max_abs = std::fmax(max_abs, std::fmax(std::fabs(it.lo[a]), std::fabs(it.hi[a])));
            }
            it.id = k;
            it.pad = 0;
        }
        DMesh m;
        std::memset(&m, 0, sizeof(m));
        m.xform = -1;
        m.tri_base = (int32_t)out_.prims.size();
        m.tri_count = (int32_t)n;
        const double pad = kPadRel * std::fmax(max_abs, std::numeric_limits<double>::min());
        m.extent = f32_above_(max_abs * (1.0 + 2.0 * kPadRel));
        m.cx = (float)centre[0];
        m.cy = (float)centre[1];
        m.cz = (float)centre[2];
        // the device measures the origin's distance in f32 from the f32 centre: both roundings are far inside the 1.25
        m.rfast2 = detail::f32_below(reach * reach);
        const int32_t node = (int32_t)out_.program.size();
        if (n <= (uint32_t)kClusterListMax) {  // a LIST: leaves in DFS order in prims[], the skip list over their boxes
            m.root = -1;
            m.entry_base = (int32_t)out_.cluster_entries.size();
            std::vector<SkipNode> nodes;
            nodes.reserve(2 * n);
            int all[kClusterListMax];
            for (uint32_t k = 0; k < n; k++) all[k] = (int)k;
            const int root = skip_build(nodes, items, all, (int)n);
            SkipMemo memo((int)nodes.size());
            // the root is never a header, but its box is what the headers beneath it are measured against
            skip_emit(nodes, nodes[root].left, root, memo, opts_.cluster_headers, pad, m.tri_base);
            skip_emit(nodes, nodes[root].right, root, memo, opts_.cluster_headers, pad, m.tri_base);
            m.entry_count = (int32_t)out_.cluster_entries.size() - m.entry_base;
            for (uint32_t k = 0; k < n; k++) {
                const ClusterItem& ci = run[k];
                set_site((uint32_t)ci.prim.leaf, LeafSite{ci.shape, node, (int32_t)out_.prims.size()});
                push_prim(ci.prim);
            }
        } else {
            std::vector<uint32_t> order;
            int depth = 0;
            const size_t nodes_before = out_.bvh.size();
            m.root = build_bvh_items(items, pad, 1, out_.bvh, m.tri_base, order, &depth);
            if (depth + 2 > kClusterStackDepth) {  // a degenerate layout (the builder bounds depth by ~log2 n otherwise)
                out_.bvh.resize(nodes_before);
                for (const ClusterItem& ci : run) {
                    set_site((uint32_t)ci.prim.leaf, LeafSite{ci.shape, (int32_t)out_.program.size(), (int32_t)out_.prims.size()});
                    out_.program.push_back(DProgramNode{NODE_PRIM, (int32_t)out_.prims.size(), 0, gate_node_});
                    push_prim(ci.prim);
                }
                return;
            }
            if (depth > out_.bvh_max_depth) out_.bvh_max_depth = depth;
            out_.feature_mask |= 1024;
            for (uint32_t slot = 0; slot < n; slot++) {
                const ClusterItem& ci = run[order[slot]];
                set_site((uint32_t)ci.prim.leaf, LeafSite{ci.shape, node, (int32_t)out_.prims.size()});
                push_prim(ci.prim);
            }
        }
        out_.program.push_back(DProgramNode{NODE_CLUSTER, (int32_t)out_.meshes.size(), 0, gate_node_});
        out_.meshes.push_back(m);
        out_.feature_mask |= 256;
    }

    // ---- classes of value-equal leaves (shape.rs:638-646) -------------------------------------------------------------
    // Shape == compares kind (derived: payload included), transform and material; tuples, matrices and colours compare with
    // is_almost_equal (|a - b| < 1e-5: utils.rs:4-6, tuple.rs:93-100, matrix.rs:174-185, color.rs:47-53), plain f64 fields
    // exactly (material.rs:3, shape.rs:13).  prepare_computations finds its containers with it (intersection.rs:33,42).
    static bool almost(double a, double b) { return std::fabs(a - b) < 0.00001; }
    template <int N>
    static bool almost_n(const double* a, const double* b) {
        for (int k = 0; k < N; k++)
            if (!almost(a[k], b[k])) return false;
        return true;
    }
    static bool material_eq(const rtc_material& a, const rtc_material& b) {
        if (!almost_n<3>(a.color, b.color)) return false;
        if (!(a.ambient == b.ambient && a.diffuse == b.diffuse && a.specular == b.specular && a.shininess == b.shininess &&
              a.reflective == b.reflective && a.transparency == b.transparency && a.refractive_index == b.refractive_index))
            return false;
        const bool pa = a.pattern_kind >= 0, pb = b.pattern_kind >= 0;
        if (pa != pb) return false;
        if (!pa) return true;
        if (a.pattern_kind != b.pattern_kind) return false;
        if (a.pattern_kind != RTC_PATTERN_TEST && !(almost_n<3>(a.pattern_a, b.pattern_a) && almost_n<3>(a.pattern_b, b.pattern_b)))
            return false;
        return almost_n<16>(a.pattern_transform, b.pattern_transform) && almost_n<16>(a.pattern_inverse, b.pattern_inverse);
    }
    bool leaf_eq(uint32_t ia, uint32_t ib) const {
        const rtc_shape_desc& a = d_.shapes[ia];
        const rtc_shape_desc& b = d_.shapes[ib];
        if (a.kind != b.kind) return false;
        if (a.kind == RTC_CYLINDER || a.kind == RTC_CONE) {
            if (!(a.minimum == b.minimum && a.maximum == b.maximum && (a.capped != 0) == (b.capped != 0))) return false;
        } else if (is_tri(a.kind)) {
            if (a.kind == RTC_SMOOTH_TRIANGLE) {
                const rtc_vertex_normals& p = d_.vertex_normals[a.triangle];
                const rtc_vertex_normals& q = d_.vertex_normals[b.triangle];
                if (!(almost_n<3>(p.n1, q.n1) && almost_n<3>(p.n2, q.n2) && almost_n<3>(p.n3, q.n3))) return false;
            }
            const rtc_triangle_desc& x = d_.triangles[a.triangle];
            const rtc_triangle_desc& y = d_.triangles[b.triangle];
            if (!(almost_n<3>(x.p1, y.p1) && almost_n<3>(x.p2, y.p2) && almost_n<3>(x.p3, y.p3) && almost_n<3>(x.e1, y.e1) &&
                  almost_n<3>(x.e2, y.e2) && almost_n<3>(x.normal, y.normal)))
                return false;
        }
        if (a.transform != b.transform &&
            !almost_n<16>(d_.transforms[a.transform].transform, d_.transforms[b.transform].transform))
            return false;
        return a.material == b.material || material_eq(d_.materials[a.material], d_.materials[b.material]);
    }
    // Equal leaves can only be found among leaves of one kind whose first compared coordinate (a triangle's p1.x, any other
    // leaf's x translation) lies within 1e-5: sort by (kind, that coordinate) and compare inside the window.
    void build_classes() {
        const uint32_t n = (uint32_t)sites_.size();
        if (n < 2) return;
        std::vector<double> key(n);
        std::vector<uint32_t> idx(n);
        for (uint32_t l = 0; l < n; l++) {
            const rtc_shape_desc& s = d_.shapes[sites_[l].shape];
            const double c = is_tri(s.kind) ? d_.triangles[s.triangle].p1[0] : d_.transforms[s.transform].transform[3];
            // kinds are 1e6 apart on the key axis only if coordinates are small; compare kinds explicitly in the window
            key[l] = c;
            idx[l] = l;
        }
        std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b] || (key[a] == key[b] && a < b); });
        std::vector<uint32_t> parent(n);
        std::iota(parent.begin(), parent.end(), 0u);
        auto find = [&](uint32_t x) {
            while (parent[x] != x) x = parent[x] = parent[parent[x]];
            return x;
        };
        std::vector<std::pair<uint32_t, uint32_t>> edges;
        for (uint32_t a = 0; a < n; a++) {
            const uint32_t la = idx[a];
            if (!(key[la] == key[la])) continue;  // NaN: equal to nothing, not even itself
            for (uint32_t b = a + 1; b < n && key[idx[b]] - key[la] < 0.00001; b++) {
                const uint32_t lb = idx[b];
                if (leaf_eq(sites_[la].shape, sites_[lb].shape)) {
                    edges.emplace_back(la, lb);
                    const uint32_t ra = find(la), rb = find(lb);
                    if (ra != rb) parent[ra < rb ? rb : ra] = ra < rb ? ra : rb;  // root = lowest leaf
                }
            }
        }
        if (edges.empty()) return;
        // == must be an equivalence on this scene's leaves for "one container per class" to be what the reference's
        // containers.position(|o| o == i.object) does; with a chain a == b == c, a != c the reference's answer depends on
        // which of them happens to sit in the list — no fixed class structure reproduces that
        std::vector<uint32_t> size(n, 0), links(n, 0);
        for (uint32_t l = 0; l < n; l++) size[find(l)]++;
        for (const auto& e : edges) links[find(e.first)]++;
        for (uint32_t l = 0; l < n; l++)
            if (size[l] > 1 && (uint64_t)links[l] * 2 != (uint64_t)size[l] * (size[l] - 1))
                fail(RTC_ERR_UNSUPPORTED, "shapes that equal a common shape within 1e-5 but not each other (shape.rs:638-646): "
                                          "the reference's n1/n2 for this world depend on intersection order");
        std::vector<int32_t> cls(n, -1);
        for (uint32_t l = 0; l < n; l++) {  // classes numbered by their lowest leaf; members in leaf order
            const uint32_t r = find(l);
            if (size[r] < 2) continue;
            if (sites_[l].slot < 0) fail(kFlattenNeedsHostBuild, "a class of value-equal leaves inside a device-built mesh");
            if (cls[r] < 0) {
                cls[r] = (int32_t)out_.class_offsets.size();
                out_.class_offsets.push_back(0);
            }
            cls[l] = cls[r];
        }
        const size_t nc = out_.class_offsets.size();
        std::vector<int32_t> count(nc, 0);
        for (uint32_t l = 0; l < n; l++)
            if (cls[l] >= 0) count[cls[l]]++;
        out_.class_offsets.assign(nc + 1, 0);
        for (size_t c = 0; c < nc; c++) out_.class_offsets[c + 1] = out_.class_offsets[c] + count[c];
        out_.class_members.resize(out_.class_offsets[nc]);
        std::vector<int32_t> fill(out_.class_offsets.begin(), out_.class_offsets.end() - 1);
        for (uint32_t l = 0; l < n; l++) {
            if (cls[l] < 0) continue;
            const LeafSite& st = sites_[l];
            out_.class_members[fill[cls[l]]++] = DClassMember{st.node, st.slot};
            if (out_.program[st.node].type != NODE_MESH) out_.prims[st.slot].cls = cls[l];
            else out_.tris[st.slot].cls = cls[l];
        }
    }

    // ---- emission ---------------------------------------------------------------------------------------------------
    bool same_inverse(uint32_t a, uint32_t b) const {
        return std::memcmp(d_.transforms[d_.shapes[a].transform].inverse, d_.transforms[d_.shapes[b].transform].inverse,
                           sizeof(double) * 16) == 0;
    }

    void emit_children(uint32_t begin, uint32_t end) {
        const bool cluster = opts_.clusters && cluster_candidates(begin, end) >= kClusterMin;
        std::vector<ClusterItem> run;  // this parent's bounded leaves, when there are enough of them for a cluster
        uint32_t i = begin;
        while (i < end) {
            const rtc_shape_desc& s = d_.shapes[i];
            if (s.kind == RTC_GROUP) {
                emit_group(i);
                i = end_[i];
            } else if (is_tri(s.kind)) {
                uint32_t j = i + 1;
                while (j < end && is_tri(d_.shapes[j].kind) && same_inverse(i, j)) j++;
                emit_mesh(i, j);
                i = j;
            } else {
                ClusterItem ci;
                DPrim& p = ci.prim;
                std::memset(&p, 0, sizeof(p));
                p.kind = s.kind;
                p.material = s.material;
                p.xform = xform_id(s.transform);
                p.capped = s.capped ? 1 : 0;
                p.minimum = s.minimum;
                p.maximum = s.maximum;
                p.leaf = (int32_t)next_leaf_++;  // DFS order whatever the leaf's place in the tables
                p.cls = -1;
                out_.feature_mask |= 1 << s.kind;
                ci.shape = i;
                if (cluster && leaf_world_box(s, ci)) {
                    run.push_back(ci);  // hit selection is min over (t, DFS leaf): the order of the tests is free
                } else {
                    ClusterItem bounded;
                    if (leaf_world_box(s, bounded)) hot_add(bounded.lo, bounded.hi);  // (a plane is not: it stays "the rest")
                    set_site((uint32_t)p.leaf, LeafSite{i, (int32_t)out_.program.size(), (int32_t)out_.prims.size()});
                    out_.program.push_back(DProgramNode{NODE_PRIM, (int32_t)out_.prims.size(), 0, gate_node_});
                    push_prim(p);
                }
                i++;
            }
        }
        if (!run.empty()) emit_cluster(run);
    }

    void require_identity(uint32_t i) const {
        // the reference intersects a group in the space of its own transform, which set_transform never changes from
        // the identity (shape.rs:203-217); anything else has no reference behaviour to match
        const double* t = d_.transforms[d_.shapes[i].transform].transform;
        Mat4 id = Mat4::identity();
        for (int k = 0; k < 16; k++)
            if (!(t[k] == id.m[k])) fail(RTC_ERR_UNSUPPORTED, "a group's own transform must be the identity (shape.rs:203-217)");
    }
    // group g holds nothing but one run of >= kDeviceBuildMin triangles that share one transform entry
    bool device_gate_run(uint32_t g) const {
        const rtc_shape_desc& s = d_.shapes[g];
        if (s.kind != RTC_GROUP || (uint32_t)s.child_count < kDeviceBuildMin || end_[g] - g - 1 != (uint32_t)s.child_count)
            return false;
        const int32_t tr = d_.shapes[g + 1].transform;
        for (uint32_t c = g + 1; c < end_[g]; c++)
            if (!is_tri(d_.shapes[c].kind) || d_.shapes[c].transform != tr) return false;
        return true;
    }

    void emit_group(uint32_t i) {
        require_identity(i);
        if (opts_.device_mesh_build && group_depth_ == 0) {
            // A World object that is a mesh — group{ triangles } or, as Parser::obj_to_group builds it,
            // group{ default_group{ triangles } } — leaves its gate box to the device build along with the mesh: the fold
            // over 8 corners per triangle is the largest host cost left (bounds.rs:50-151).  The outer group's box is
            // the inner one's (identity child transform, both origin-seeded), so one gate stands for both, as below.
            uint32_t g = i;
            const bool nested = d_.shapes[i].child_count == 1 && d_.shapes[i + 1].kind == RTC_GROUP;
            if (nested) g = i + 1;
            if (device_gate_run(g)) {
                if (nested) {
                    require_identity(g);
                    out_.merged_gates++;
                }
                const size_t at = out_.program.size();
                out_.program.push_back(DProgramNode{NODE_GATE, (int32_t)out_.gates.size(), 0, gate_node_});
                DGate zero;
                std::memset(&zero, 0, sizeof(zero));
                out_.gates.push_back(zero);
                pending_gate_ = (int32_t)out_.gates.size() - 1;
                const int32_t saved_node = gate_node_;
                gate_node_ = (int32_t)at;
                group_depth_++;
                emit_children(g + 1, end_[g]);  // one run -> one pending mesh, which takes pending_gate_
                group_depth_--;
                gate_node_ = saved_node;
                if (pending_gate_ != -1) fail(RTC_ERR_INVALID, "internal: device gate not taken by its mesh");
                out_.program[at].skip = (int32_t)out_.program.size();
                return;
            }
        }
        PhaseClock clock;
        Box4 b = bounds_of(i);
        clock.lap(out_.phase_ms, FlatScene::T_BOUNDS);
        DGate g;
        g.lo[0] = b.min.x; g.lo[1] = b.min.y; g.lo[2] = b.min.z;
        g.hi[0] = b.max.x; g.hi[1] = b.max.y; g.hi[2] = b.max.z;
        // group{ group{...} } with bit-identical boxes (what Parser::obj_to_group builds for an OBJ without `g` lines,
        // obj_file.rs:120-128: the inner box already contains the origin the outer one is seeded with): both gates give
        // the same verdict for every ray, so one test stands for both
        if (parent_gate_ >= 0 && only_child_ && std::memcmp(&out_.gates[parent_gate_], &g, sizeof(g)) == 0) {
            const int32_t saved_parent = parent_gate_;
            const bool saved_only = only_child_;
            only_child_ = (d_.shapes[i].child_count == 1);
            group_depth_++;
            emit_children(i + 1, end_[i]);
            group_depth_--;
            parent_gate_ = saved_parent;
            only_child_ = saved_only;
            out_.merged_gates++;
            return;
        }
        size_t at = out_.program.size();
        out_.program.push_back(DProgramNode{NODE_GATE, (int32_t)out_.gates.size(), 0, gate_node_});
        out_.gates.push_back(g);
        const int32_t saved_parent = parent_gate_;
        const bool saved_only = only_child_;
        const int32_t saved_node = gate_node_;
        parent_gate_ = (int32_t)out_.gates.size() - 1;
        only_child_ = (d_.shapes[i].child_count == 1);
        gate_node_ = (int32_t)at;
        group_depth_++;
        emit_children(i + 1, end_[i]);
        group_depth_--;
        parent_gate_ = saved_parent;
        only_child_ = saved_only;
        gate_node_ = saved_node;
        out_.program[at].skip = (int32_t)out_.program.size();
    }

    void emit_mesh(uint32_t begin, uint32_t end) {
        const int32_t xf = xform_id(d_.shapes[begin].transform);
        const rtc_transform_desc& td = d_.transforms[d_.shapes[begin].transform];
        const Mat4 inv_t = transpose(Mat4::from(td.inverse));  // shape.rs:216
        const uint32_t n = end - begin;
        PhaseClock clock;
        if (opts_.device_mesh_build && n >= kDeviceBuildMin) {
            PendingMesh p;
            p.mesh_index = (int32_t)out_.meshes.size();
            p.n = n;
            p.xform = xf;
            p.leaf0 = (int32_t)next_leaf_;
            p.tri_base = p.node_base = -1;
            std::memcpy(p.inv_t, inv_t.m, sizeof(p.inv_t));
            p.input_offset = out_.pending_material.size();
            out_.pending_material.resize(p.input_offset + n);
            const int32_t t0 = d_.shapes[begin].triangle;
            bool consecutive = true;
            for (uint32_t k = 0; k < n; k++) {
                const rtc_shape_desc& s = d_.shapes[begin + k];
                out_.pending_material[p.input_offset + k] = s.material;
                consecutive = consecutive && s.triangle == t0 + (int32_t)k;
            }
            p.direct = consecutive ? d_.triangles + t0 : nullptr;
            p.gate_index = pending_gate_;
            pending_gate_ = -1;
            std::memcpy(p.transform, td.transform, sizeof(p.transform));
            if (!consecutive) {  // gathered copy, index-aligned with pending_material
                out_.pending_tri.resize(p.input_offset + n);
                for (uint32_t k = 0; k < n; k++)
                    out_.pending_tri[p.input_offset + k] = d_.triangles[d_.shapes[begin + k].triangle];
            }
            next_leaf_ += n;
            DMesh m;
            m.xform = xf;
            m.root = -1;  // written by the device build, like extent
            m.tri_base = -1;
            m.tri_count = (int32_t)n;
            m.extent = 0.f;
            m.cx = m.cy = m.cz = m.rfast2 = 0.f;
            m.entry_base = m.entry_count = m.pad = 0;
            out_.pending.push_back(p);
            for (uint32_t k = 0; k < n; k++) set_site(p.leaf0 + k, LeafSite{begin + k, (int32_t)out_.program.size(), -1});
            out_.program.push_back(DProgramNode{NODE_MESH, (int32_t)out_.meshes.size(), 0, gate_node_});
            out_.meshes.push_back(m);
            clock.lap(out_.phase_ms, FlatScene::T_TRIANGLES);
            return;
        }
        std::vector<BvhTri> bt(n);
        for (uint32_t k = 0; k < n; k++) {
            const rtc_triangle_desc& t = d_.triangles[d_.shapes[begin + k].triangle];
            for (int a = 0; a < 3; a++) {
                bt[k].p[0][a] = t.p1[a];
                bt[k].p[1][a] = t.p2[a];
                bt[k].p[2][a] = t.p3[a];
            }
        }
        // normal_at for a triangle (shape.rs:509-518) is point-independent: invT * normal, w = 0, normalize, w = 0,
        // normalize.  Computed per input triangle on a second thread while the BVH builds (it is a third of the serial
        // time of a mesh otherwise), placed in leaf order afterwards.
        std::vector<DTriAttr> attr_in(n);
        auto attrs = [&] {
            for (uint32_t k = 0; k < n; k++) {
                const rtc_shape_desc& s = d_.shapes[begin + k];
                const rtc_triangle_desc& t = d_.triangles[s.triangle];
                Vec4 wn = mul(inv_t, vector(t.normal[0], t.normal[1], t.normal[2]));
                wn.w = 0.;
                wn = normalize(wn);
                wn.w = 0.;
                wn = normalize(wn);
                DTriAttr& ta = attr_in[k];
                ta.normal[0] = wn.x; ta.normal[1] = wn.y; ta.normal[2] = wn.z;
                ta.material = s.material;
                ta.xform = xf;
            }
        };
        std::thread attr_thread;
        if (n >= kParallelMin) attr_thread = std::thread(attrs);
        else attrs();
        DMesh m;
        m.xform = xf;
        m.tri_base = (int32_t)out_.tris.size();
        m.tri_count = (int32_t)n;
        std::vector<uint32_t> order;
        int depth = 0;
        double max_abs = 0.;
        clock.lap(out_.phase_ms, FlatScene::T_BVH_ITEMS);
        const size_t nodes_before = out_.bvh.size();
        m.root = build_bvh(bt, out_.bvh, m.tri_base, order, &depth, &max_abs, out_.phase_ms + FlatScene::T_BVH_ITEMS);
        if (m.root >= 0) m.root = breadth_first(nodes_before, m.root);
        clock = PhaseClock();  // build_bvh booked its own phases
        if (attr_thread.joinable()) attr_thread.join();
        m.extent = f32_above_(max_abs * (1.0 + 2.0 * kPadRel));
        m.cx = m.cy = m.cz = m.rfast2 = 0.f;
        m.entry_base = m.entry_count = m.pad = 0;
        if (depth > out_.bvh_max_depth) out_.bvh_max_depth = depth;
        if (depth + 2 > kBvhStackDepth)
            fail(RTC_ERR_UNSUPPORTED, "mesh BVH deeper than the device traversal stack (more than ~16M triangles)");
        const uint32_t leaf0 = next_leaf_;
        next_leaf_ += n;
        const size_t at = out_.tris.size();
        out_.tris.resize(at + n);
        out_.tri_attr.resize(at + n);
        for (uint32_t slot = 0; slot < n; slot++) {
            const uint32_t k = order[slot];
            const rtc_triangle_desc& t = d_.triangles[d_.shapes[begin + k].triangle];
            DTri& dt = out_.tris[at + slot];
            std::memset(&dt, 0, sizeof(dt));
            for (int a = 0; a < 3; a++) {
                dt.p1[a] = t.p1[a];
                dt.e1[a] = t.e1[a];
                dt.e2[a] = t.e2[a];
            }
            dt.leaf = (int32_t)(leaf0 + k);
            dt.cls = -1;
            out_.tri_attr[at + slot] = attr_in[k];
            if (has_smooth_) {
                out_.tri_smooth.resize(at + n);  // value-initialised: smooth = 0
                const rtc_shape_desc& sd = d_.shapes[begin + k];
                if (sd.kind == RTC_SMOOTH_TRIANGLE) {
                    DTriSmooth& sm = out_.tri_smooth[at + slot];
                    const rtc_vertex_normals& vn = d_.vertex_normals[sd.triangle];
                    for (int a = 0; a < 3; a++) {
                        sm.n1[a] = vn.n1[a];
                        sm.n2[a] = vn.n2[a];
                        sm.n3[a] = vn.n3[a];
                    }
                    sm.smooth = 1;
                    out_.feature_mask |= 2048;
                }
            }
        }
        for (uint32_t slot = 0; slot < n; slot++)
            set_site(leaf0 + order[slot], LeafSite{begin + order[slot], (int32_t)out_.program.size(), (int32_t)(at + slot)});
        clock.lap(out_.phase_ms, FlatScene::T_TRIANGLES);
        out_.program.push_back(DProgramNode{NODE_MESH, (int32_t)out_.meshes.size(), 0, gate_node_});
        out_.meshes.push_back(m);
    }
};

}  // namespace detail

// Returns RTC_OK or a negative code with *err set.
inline int flatten_scene(const rtc_scene_desc& desc, FlatScene& out, std::string* err, const FlattenOptions& opts = {}) {
    try {
        detail::Flattener f(desc, out, opts);
        f.run();
        return RTC_OK;
    } catch (const FlattenError& e) {
        if (e.code == kFlattenNeedsHostBuild && opts.device_mesh_build) {
            FlattenOptions host = opts;
            host.device_mesh_build = false;
            out = FlatScene();
            return flatten_scene(desc, out, err, host);
        }
        if (err) *err = e.message;
        return e.code;
    }
}

}  // namespace rtc
