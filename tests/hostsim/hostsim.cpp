// tests/hostsim/hostsim.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the product's per-ray device program (csrc/rt_core.cuh, written __host__ __device__) and its flattener with
// g++ so the `-m "not gpu"` tests can single-step exactly the code the GPU runs — same tables, same BVH, same traversal
// and tie-break logic — against the oracle on this GPU-less box.  It is built into tests/_build/, is not linked into
// librtc_b200.so, and nothing in the product imports it: the product has no CPU rendering path.
#include <atomic>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtc.h"
#include "../../ray-tracer-challenge-rust_b200/csrc/flatten.hpp"
#include "../../ray-tracer-challenge-rust_b200/csrc/rt_core.cuh"
#include "lbvh_sim.hpp"

using namespace rtc;
using namespace rtc::core;

namespace {
thread_local std::string g_err;
struct Sim {
    FlatScene flat;
    DScene view{};
};
}  // namespace

extern "C" {
const char* sim_last_error() { return g_err.c_str(); }

// device_build != 0: meshes go through the simulated device build (lbvh_sim.hpp) instead of the host SAH builder
int sim_scene_create_ex(const rtc_scene_desc* desc, int device_build, void** out, int* depth_out) {
    Sim* s = new Sim();
    std::string e;
    FlattenOptions opts;
    opts.device_mesh_build = (device_build & 1) != 0;
    opts.clusters = (device_build & 2) == 0;  // bit 1: no clusters (every bounded leaf stays a PRIM entry)
    int rc = flatten_scene(*desc, s->flat, &e, opts);
    if (rc != RTC_OK) {
        g_err = e;
        delete s;
        return rc;
    }
    const int depth = lbvh_build_sim(s->flat);
    if (depth_out) *depth_out = (device_build & 1) ? depth : s->flat.bvh_max_depth;
    DScene& v = s->view;
    v.program = s->flat.program.data();
    v.xforms = s->flat.xforms.data();
    v.prims = s->flat.prims.data();
    v.gates = s->flat.gates.data();
    v.meshes = s->flat.meshes.data();
    v.bvh = s->flat.bvh.data();
    v.tris = s->flat.tris.data();
    v.tri_attr = s->flat.tri_attr.data();
    v.materials = s->flat.materials.data();
    v.cluster_entries = s->flat.cluster_entries.data();
    v.tri_smooth = s->flat.tri_smooth.empty() ? nullptr : s->flat.tri_smooth.data();
    v.class_offsets = s->flat.class_offsets.data();
    v.class_members = s->flat.class_members.data();
    v.n_classes = s->flat.class_offsets.empty() ? 0 : (int32_t)s->flat.class_offsets.size() - 1;
    v.program_count = (int32_t)s->flat.program.size();
    v.recursion_limit = s->flat.recursion_limit;
    for (int k = 0; k < 3; k++) {
        v.light_pos[k] = s->flat.light_pos[k];
        v.light_int[k] = s->flat.light_int[k];
    }
    *out = s;
    return 0;
}
int sim_scene_create(const rtc_scene_desc* desc, void** out) { return sim_scene_create_ex(desc, 0, out, nullptr); }
void sim_scene_destroy(void* s) { delete (Sim*)s; }
// the gate boxes (6 doubles each: lo xyz, hi xyz); returns how many the scene has
uint64_t sim_scene_gates(void* scene, double* out, uint64_t cap) {
    Sim* s = (Sim*)scene;
    for (uint64_t i = 0; i < s->flat.gates.size() && i < cap; i++) {
        std::memcpy(out + 6 * i, s->flat.gates[i].lo, 24);
        std::memcpy(out + 6 * i + 3, s->flat.gates[i].hi, 24);
    }
    return s->flat.gates.size();
}
// table sizes and a content hash of the mesh tables (FNV-1a over bvh, tris, tri_attr), to compare builds
// the skip lists of the LIST clusters: per entry (skip, prim); returns the number of entries
uint64_t sim_scene_cluster_entries(void* scene, int32_t* out, uint64_t cap) {
    Sim* s = (Sim*)scene;
    const uint64_t n = s->flat.cluster_entries.size();
    for (uint64_t i = 0; i < n && i < cap; i++) {
        out[2 * i] = s->flat.cluster_entries[i].skip;
        out[2 * i + 1] = s->flat.cluster_entries[i].prim;
    }
    return n;
}
void sim_scene_tables(void* scene, uint64_t n[4]) {
    Sim* s = (Sim*)scene;
    n[0] = s->flat.bvh.size();
    n[1] = s->flat.tris.size();
    n[2] = s->flat.meshes.size();
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t bytes) {
        const unsigned char* b = (const unsigned char*)p;
        for (size_t i = 0; i < bytes; i++) h = (h ^ b[i]) * 1099511628211ull;
    };
    mix(s->flat.bvh.data(), s->flat.bvh.size() * sizeof(DBvhNode));
    mix(s->flat.tris.data(), s->flat.tris.size() * sizeof(DTri));
    mix(s->flat.tri_attr.data(), s->flat.tri_attr.size() * sizeof(DTriAttr));
    n[3] = h;
}

// pixel_xy == NULL: whole frame.  out_rgb: 3 f64 per pixel; counters[4] = primary, shadow, reflect, refract.
int sim_render(void* scene, const rtc_camera_desc* cam, const uint32_t* pixel_xy, uint64_t npixels, int nthreads,
               double* out_rgb, uint8_t* out_rgba8, uint64_t* counters) {
    Sim* s = (Sim*)scene;
    DCamera dc;
    dc.hsize = cam->hsize;
    dc.vsize = cam->vsize;
    std::memcpy(dc.inv, cam->inverse, sizeof(dc.inv));
    dc.half_width = cam->half_width;
    dc.half_height = cam->half_height;
    dc.pixel_size = cam->pixel_size;
    const uint64_t total = pixel_xy ? npixels : (uint64_t)cam->hsize * cam->vsize;
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint64_t> next{0};
    std::vector<RayCounters> rcs(nthreads);
    auto worker = [&](int tid) {
        for (;;) {
            uint64_t b = next.fetch_add(256);
            if (b >= total) break;
            uint64_t e = b + 256 < total ? b + 256 : total;
            for (uint64_t i = b; i < e; i++) {
                uint32_t x, y;
                if (pixel_xy) {
                    x = pixel_xy[2 * i];
                    y = pixel_xy[2 * i + 1];
                } else {
                    x = (uint32_t)(i % cam->hsize);
                    y = (uint32_t)(i / cam->hsize);
                }
                Ray r = ray_for_pixel(dc, x, y);
                Tally tl;
                V3 c = color_at_any(s->view, r, rcs[tid], tl);
                if (out_rgb) {
                    out_rgb[3 * i] = c.x;
                    out_rgb[3 * i + 1] = c.y;
                    out_rgb[3 * i + 2] = c.z;
                }
                if (out_rgba8) {
                    out_rgba8[4 * i] = (uint8_t)quantise(c.x);
                    out_rgba8[4 * i + 1] = (uint8_t)quantise(c.y);
                    out_rgba8[4 * i + 2] = (uint8_t)quantise(c.z);
                    out_rgba8[4 * i + 3] = 255;
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
    if (counters) {
        counters[0] = total;
        counters[1] = counters[2] = counters[3] = 0;
        for (auto& rc : rcs) {
            counters[1] += rc.shadow;
            counters[2] += rc.reflect;
            counters[3] += rc.refract;
        }
    }
    return 0;
}

int sim_color_at(void* scene, const double* rays, uint64_t n, double* rgb) {
    Sim* s = (Sim*)scene;
    RayCounters rc;
    for (uint64_t i = 0; i < n; i++) {
        Ray r{v3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
        Tally tl;
        V3 c = color_at_any(s->view, r, rc, tl);
        rgb[3 * i] = c.x;
        rgb[3 * i + 1] = c.y;
        rgb[3 * i + 2] = c.z;
    }
    return 0;
}
}
