// device_scene.h — plain-old-data tables shared by the host flattener and the CUDA kernels.
//
// HBM layout (everything f64 or i32, uploaded once per scene; a whole scene is < 2 MB so it lives in L2/L1 and the
// only streaming HBM traffic of a frame is the framebuffer store):
//   program[]  : the World's Shape tree flattened in DFS order.  A GATE entry is a reference Group's bounding-box test
//                (shape.rs:399-425) with a skip link past its subtree; a PRIM entry is one non-triangle leaf; a MESH
//                entry is a run of sibling triangles that share one transform, traversed through a BVH.
//   xforms[]   : rows 0..2 of each distinct Shape.transform_inverse (row 3 is (0,0,0,1) for affine input — checked).
//   prims[]    : tag-switched non-triangle leaves (sphere/plane/cube/cylinder/cone).
//   gates[]    : world-space boxes exactly as Bounds::new computes them (bounds.rs:50-125).
//   meshes[]   : per mesh: transform, BVH root, triangle range.
//   bvh[]      : 64-byte nodes holding BOTH children's boxes (f32, padded and rounded outward) so one fetch tests two boxes.
//   tris[]     : p1,e1,e2 of each triangle in BVH-leaf order + its DFS leaf index (the reference's tie-break order).
//   tri_attr[] : per triangle (same order): precomputed world normal (shape.rs:509-518 is point-independent) + material.
//   tri_smooth[]: per triangle (same order), only in scenes with smooth triangles: the three vertex normals.
//   materials[]: material.rs:4-14 + flattened pattern with rows 0..2 of the pattern inverse.
#pragma once
#include <stdint.h>

namespace rtc {

enum : int32_t { NODE_GATE = 0, NODE_PRIM = 1, NODE_MESH = 2, NODE_CLUSTER = 3 };

struct DProgramNode {
    int32_t type;    // NODE_*
    int32_t index;   // into gates / prims / meshes
    int32_t skip;    // GATE: program index just past the group's subtree
    int32_t parent;  // program index of the enclosing GATE entry, -1 at World level (the class pass of the n1/n2 walk)
};
struct DXform {
    double m[12];  // inverse, rows 0..2
};
// A non-triangle leaf.  Bounded leaves with enough bounded siblings are gathered into a CLUSTER — a small BVH over their
// padded world boxes (see DMesh) — and stored in its leaf order; the others are PRIM entries of the program.
struct alignas(16) DPrim {
    int32_t kind, material, xform, capped;
    double minimum, maximum;
    int32_t leaf;  // DFS leaf index
    int32_t cls;   // >= 0: member of a class of value-equal leaves (shape.rs:638-646), see DClassMember
    int32_t pad[2];
};
static_assert(sizeof(DPrim) == 48, "DPrim layout: read with 16-byte loads");
// A cluster of up to kClusterListMax leaves is a SKIP LIST of padded f32 world boxes (rounded outward), walked front to
// back with no stack: a LEAF entry (skip < 0) is one leaf's box — the exact test of prims[prim] runs if the ray passes it;
// a HEADER entry (skip >= 0) is the box around the next `skip` entries — a ray that misses it jumps past them.  Headers are
// the inner nodes of a SAH tree over the leaves, kept only where the expected number of box tests falls (flatten.hpp).
// Measured against a stack-based BVH walk over the same few boxes (table scene, 18 cubes: 0.93 ms) and against the flat
// list without headers (0.62 ms): profiles/r02c_variants.json, r02p_variants.json.  Larger clusters get the BVH.
struct alignas(16) DBox32 {
    float lo[3], hi[3];
    int32_t skip;  // >= 0: header over the next `skip` entries; -1: leaf
    int32_t prim;  // leaf: index into prims[]
};
static_assert(sizeof(DBox32) == 32, "DBox32 layout: read with two 16-byte loads");
constexpr int kClusterListMax = 32;
struct DGate {
    double lo[3], hi[3];
};
// MESH: a run of sibling triangles sharing one transform, traversed in the mesh's object space through a BVH.
// CLUSTER (xform = -1): sibling spheres / cubes / finite cylinders under one parent, traversed with the WORLD ray through a
// BVH over their world boxes, each leaf run then tested exactly in its own object space.  A cube's check_axis treats an
// object-space direction component below EPSILON as parallel (shape.rs:593-599) and then reports intersections that have
// drifted up to EPSILON * t outside the true cube, so cluster boxes of cubes are padded by EPSILON * kClusterReach * (the
// cube's largest column norm): valid for unit-length rays that start within `reach` of the cluster's centre (rfast2 =
// reach^2), which every camera, shadow, reflection and refraction ray of a scene around the cluster does; any other ray
// takes the exact linear scan of the cluster's leaves.
struct DMesh {
    int32_t xform, root, tri_base, tri_count;  // root < 0: no BVH — a tiny mesh (scan its triangles) or a LIST cluster
    float extent;                              // max |coordinate| of the boxes (f32 slab error bound)
    float cx, cy, cz, rfast2;                  // CLUSTER: centre and squared reach of the fast path
    int32_t entry_base, entry_count;           // LIST cluster: its skip list is cluster_entries[entry_base .. + entry_count)
    int32_t pad;
};
static_assert(sizeof(DMesh) == 48, "DMesh layout");
// f32 boxes rounded OUTWARD from the padded f64 boxes: 64 bytes hold both children, one fetch decides two subtrees.
struct alignas(64) DBvhNode {
    float lo0[3], hi0[3], lo1[3], hi1[3];
    int32_t child0, count0, child1, count1;  // count > 0: leaf, child = first triangle slot; count == 0: inner node
};
struct alignas(16) DTri {
    double p1[3], e1[3], e2[3];
    int32_t leaf;
    int32_t cls;  // >= 0: member of a class of value-equal leaves, else -1
};
// Leaves the reference's Shape equality (shape.rs:638-646: kind payload, transform and material, 1e-5 tolerance on
// tuples / matrices / colours) cannot tell apart form a CLASS; the n1/n2 container walk (intersection.rs:29-62) treats a
// class as ONE container.  Only classes of two or more leaves are listed: members of class c are
// class_members[class_offsets[c] .. class_offsets[c + 1]).
struct DClassMember {
    int32_t node;  // program index of the member's PRIM entry, or of the MESH entry its triangle belongs to
    int32_t slot;  // prims[] index, or tris[] slot
};
struct alignas(16) DTriAttr {
    double normal[3];
    int32_t material;
    int32_t xform;  // the mesh's transform (patterns evaluate in the leaf's object space, pattern.rs:99)
};
// Vertex normals of a smooth triangle (the book's SmoothTriangle; rtc.h RTC_SMOOTH_TRIANGLE), object space, indexed like
// tris[]; the table exists only in scenes that hold one (smooth = 0: a flat triangle of such a scene).
struct alignas(16) DTriSmooth {
    double n1[3], n2[3], n3[3];
    int32_t smooth;
    int32_t pad;
};
static_assert(sizeof(DTriSmooth) == 80, "DTriSmooth layout: read with 16-byte loads");
struct DMaterial {
    double color[3];
    double ambient, diffuse, specular, shininess, reflective, transparency, refractive_index;
    double pa[3], pb[3];
    double pinv[12];  // pattern inverse rows 0..2
    int32_t pattern_kind;
    int32_t pad;
};
struct DScene {
    const DProgramNode* program;
    const DXform* xforms;
    const DPrim* prims;
    const DGate* gates;
    const DMesh* meshes;
    const DBvhNode* bvh;
    const DTri* tris;
    const DTriAttr* tri_attr;
    const DMaterial* materials;
    const DBox32* cluster_entries;
    const DTriSmooth* tri_smooth;  // null: no smooth triangle in the scene
    const int32_t* class_offsets;
    const DClassMember* class_members;
    int32_t n_classes;
    int32_t pad1;
    int32_t program_count;
    int32_t recursion_limit;  // World's RECURSION_LIMIT (world.rs:11): 5 in the reference
    int32_t n_bvh;            // host-built BVH nodes (bvh[0 .. n_bvh); device-built ones follow)
    int32_t pad0;
    uint32_t n_prims, n_xforms, n_gates, n_materials;  // table lengths (shared-memory staging)
    double light_pos[3];
    double light_int[3];
};
struct DCamera {
    uint32_t hsize, vsize;
    double inv[12];  // transform_inverse rows 0..2
    double half_width, half_height, pixel_size;
};
struct DRows {
    uint32_t band_rows, band_first, band_stride, local_rows;  // local_rows = rows this call renders
    uint32_t frame_layout;                                    // 1: outputs are addressed by FRAME row, not by local row
    uint32_t row_begin, row_count;                            // the local rows THIS launch renders (a chunk of the call)
    uint32_t pad;
    // 0, or the address of a uint32 counter (device memory of ANY GPU mapped into this device): the launch's last CTA adds 1
    // to it with system scope once every pixel store of the launch is visible system-wide — how a rank tells the frame's
    // owner that its bands have landed (multi-process sharded renders: multi.py)
    unsigned long long notify;
    // Tile ORDER of the queue (never which tiles are rendered): the tiles of the rectangle [hot_x0, hot_x1) x [hot_y0, hot_y1)
    // — where the scene's bounded geometry projects to, the expensive pixels — are handed out first, the rest after them, so
    // the launch drains on cheap background tiles instead of on the last mesh tiles (render_inst.cu tile_of).
    uint32_t hot_x0, hot_y0, hot_x1, hot_y1;
};
// depth of the per-thread traversal stack; flatten.hpp refuses a mesh whose BVH is deeper (the builder bounds depth)
#if !defined(RTC_BVH_STACK_DEPTH)  // A/B switch (tools/tune_variants.py, tools/dram_variants.py)
#define RTC_BVH_STACK_DEPTH 48
#endif
constexpr int kBvhStackDepth = RTC_BVH_STACK_DEPTH;
// the same for a kernel whose scenes hold clusters but no meshes (a cluster tree deeper than this is not built: its leaves
// stay PRIM entries)
constexpr int kClusterStackDepth = 16;
// largest RECURSION_LIMIT the general-depth integrator accepts (six shaded generations, up to 63 shaded hits per pixel)
constexpr int kMaxRecursionLimit = 18;

}  // namespace rtc
