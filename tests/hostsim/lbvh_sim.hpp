// tests/hostsim/lbvh_sim.hpp — TEST INFRASTRUCTURE ONLY.
//
// The device mesh build (csrc/lbvh.cuh) run on the host: each kernel body is a plain function of a thread index, so the
// simulation is a loop per launch and std::stable_sort for the radix sort.  It lets the CPU tests check, against the
// oracle and bit for bit, exactly the tables the GPU build writes (the steps are deterministic: stable sort, order-
// independent min/max merges).
#pragma once
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>
#include "../../ray-tracer-challenge-rust_b200/csrc/flat_scene.hpp"
#include "../../ray-tracer-challenge-rust_b200/csrc/lbvh.cuh"
namespace rtc {
// Builds every pending mesh of `f` into its tables, as lbvh_build_device does; returns the deepest tree's depth.
inline int lbvh_build_sim(FlatScene& f) {
    int depth_all = 0;
    f.tris.resize(f.tris.size() + f.device_tris);
    f.tri_attr.resize(f.tri_attr.size() + f.device_tris);
    const size_t host_nodes = f.bvh.size();
    f.bvh.resize(host_nodes + f.device_nodes);
    if (f.device_nodes) std::memset(f.bvh.data() + host_nodes, 0, sizeof(DBvhNode) * f.device_nodes);
    for (const PendingMesh& p : f.pending) {
        const uint32_t n = p.n;
        lbvh::Work w{};
        w.tri = p.direct ? p.direct : f.pending_tri.data() + p.input_offset;
        w.material = f.pending_material.data() + p.input_offset;
        w.n = n;
        w.xform = p.xform;
        w.leaf0 = p.leaf0;
        w.tri_base = p.tri_base;
        w.node_base = p.node_base;
        std::memcpy(w.inv_t, p.inv_t, sizeof(w.inv_t));
        std::vector<unsigned long long> gbox(7), keys(n);
        std::vector<uint32_t> order(n), arrive(n - 1, 0);
        std::vector<int32_t> left(n - 1), right(n - 1), parent(n - 1), first(n - 1), last(n - 1), leaf_parent(n);
        std::vector<double> box(6 * (size_t)(n - 1)), leaf_box(6 * (size_t)n);
        int32_t depth_max = 0, error = 0;
        unsigned long long gate_acc[6];
        w.gate_acc = gate_acc;
        w.error = &error;
        w.gate_out = p.gate_index >= 0 ? &f.gates[p.gate_index] : nullptr;
        std::memcpy(w.tr, p.transform, sizeof(w.tr));
        w.gbox = gbox.data(); w.keys = keys.data(); w.order = order.data(); w.left = left.data(); w.right = right.data();
        w.parent = parent.data(); w.first = first.data(); w.last = last.data(); w.leaf_parent = leaf_parent.data();
        w.arrive = arrive.data(); w.box = box.data(); w.leaf_box = leaf_box.data(); w.depth_max = &depth_max;
        w.nodes = f.bvh.data(); w.tris = f.tris.data(); w.attr = f.tri_attr.data(); w.mesh = &f.meshes[p.mesh_index];
        lbvh::gbox_init(w);
        for (uint32_t k = 0; k < n; k++) lbvh::tri_bounds(w, k);
        if (w.gate_out)
            for (uint32_t k = 0; k < n; k++) lbvh::gate_fold(w, k);
        for (uint32_t k = 0; k < n; k++) lbvh::tri_morton(w, k);
        std::vector<uint32_t> perm(n);
        std::iota(perm.begin(), perm.end(), 0u);
        std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
        std::vector<unsigned long long> ks(n);
        for (uint32_t s = 0; s < n; s++) {
            ks[s] = keys[perm[s]];
            order[s] = perm[s];
        }
        keys = ks;
        w.keys = keys.data();
        for (uint32_t i = 0; i + 1 < n; i++) lbvh::hierarchy(w, i);
        for (uint32_t s = 0; s < n; s++) lbvh::fit(w, s);
        for (uint32_t i = 0; i + 1 < n; i++) lbvh::emit(w, i);
        if (depth_max > depth_all) depth_all = depth_max;
        if (error) depth_all = 1 << 20;  // "rebuild on the host", as the product does
    }
    f.pending.clear();
    f.device_tris = f.device_nodes = 0;
    return depth_all;
}
}  // namespace rtc
