// flat_scene.hpp — the flattened World as host vectors (output of flatten.hpp, input of the device upload).
#pragma once
#include <chrono>
#include <cstdint>
#include <vector>

#include "../../include/rtc.h"
#include "device_scene.h"

namespace rtc {

// A mesh left for the device to build (lbvh.cuh): its input triangles, and where its slots and nodes go — after the
// host-built entries of the tris / tri_attr / bvh tables.
struct PendingMesh {
    int32_t mesh_index;  // into FlatScene::meshes
    uint32_t n;
    int32_t xform, leaf0;
    int32_t tri_base, node_base;  // assigned when flattening ends
    double inv_t[16];             // transpose of the transform's inverse (shape.rs:216)
    size_t input_offset;          // first entry in pending_material (and in pending_tri when `direct` is null)
    // the run's triangles where they lie in the caller's description, when their indices are consecutive (the usual
    // case: Parser::obj_to_group pushes them in order) — borrowed for the duration of rtc_scene_create, saves a gather
    const rtc_triangle_desc* direct;
    // >= 0: this run is all its group holds and the group is a World object (or the only child of one): the device folds
    // the group's gate box too (bounds.rs:50-151) into gates[gate_index], through the children's `transform`
    int32_t gate_index;
    double transform[16];
};

struct FlatScene {
    std::vector<DProgramNode> program;
    std::vector<DXform> xforms;
    std::vector<DPrim> prims;
    std::vector<DTriSmooth> tri_smooth;  // empty, or same length as tris
    std::vector<DBox32> cluster_entries;  // skip lists of the LIST clusters (device_scene.h DBox32)
    std::vector<DGate> gates;
    std::vector<DMesh> meshes;
    std::vector<DBvhNode> bvh;
    std::vector<DTri> tris;
    std::vector<DTriAttr> tri_attr;
    std::vector<DMaterial> materials;
    std::vector<int32_t> class_offsets;        // n_classes + 1 entries (empty: no class of value-equal leaves)
    std::vector<DClassMember> class_members;
    double light_pos[3] = {0, 0, 0}, light_int[3] = {0, 0, 0};
    // union of the world boxes of everything bounded (group gates, clusters, bounded leaves): where the expensive pixels are
    double hot_lo[3] = {1e300, 1e300, 1e300}, hot_hi[3] = {-1e300, -1e300, -1e300};
    uint64_t leaf_count = 0;
    int32_t recursion_limit = 5;  // world.rs:11
    int32_t feature_mask = 0;  // bit k: leaves of ShapeKind k; 32 meshes; 64 gates; 128 a transparent material; 256 clusters; 512 a RECURSION_LIMIT other than 5;
                               // 1024 a cluster that is a BVH (more than kClusterListMax leaves); 2048 a smooth triangle
    int32_t merged_gates = 0;  // nested single-child groups whose identical box shares the parent's gate
    int bvh_max_depth = 0;
    // device-built meshes (flatten option device_mesh_build)
    std::vector<PendingMesh> pending;
    std::vector<rtc_triangle_desc> pending_tri;
    std::vector<int32_t> pending_material;
    uint32_t device_tris = 0, device_nodes = 0;  // table entries appended after the host-built ones
    // where rtc_scene_create's host time went (ms), printed under RTC_B200_TRACE=1
    enum { T_VALIDATE, T_BOUNDS, T_BVH_ITEMS, T_BVH_BUILD, T_BVH_SPLICE, T_TRIANGLES, T_UPLOAD, T_COUNT };
    double phase_ms[T_COUNT] = {0, 0, 0, 0, 0, 0, 0};
};

struct PhaseClock {  // adds the time since construction (or the last lap) to a FlatScene phase
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(double* phase_ms, int phase) {
        const auto t1 = std::chrono::steady_clock::now();
        phase_ms[phase] += std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
    }
};

}  // namespace rtc
