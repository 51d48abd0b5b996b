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
//   bvh[]      : 128-byte nodes holding BOTH children's boxes (padded, conservative) so one fetch tests two boxes.
//   tris[]     : p1,e1,e2 of each triangle in BVH-leaf order + its DFS leaf index (the reference's tie-break order).
//   tri_attr[] : per triangle (same order): precomputed world normal (shape.rs:509-518 is point-independent) + material.
//   materials[]: material.rs:4-14 + flattened pattern with rows 0..2 of the pattern inverse.
#pragma once
#include <stdint.h>

namespace rtc {

enum : int32_t { NODE_GATE = 0, NODE_PRIM = 1, NODE_MESH = 2 };

struct DProgramNode {
    int32_t type;   // NODE_*
    int32_t index;  // into gates / prims / meshes
    int32_t skip;   // GATE: program index just past the group's subtree
    int32_t pad;
};
struct DXform {
    double m[12];  // inverse, rows 0..2
};
struct DPrim {
    int32_t kind, material, xform, capped;
    double minimum, maximum;
    int32_t leaf;  // DFS leaf index
    int32_t pad;
};
struct DGate {
    double lo[3], hi[3];
};
struct DMesh {
    int32_t xform, root, tri_base, tri_count;  // root < 0: no BVH (tiny mesh), scan [tri_base, tri_base + tri_count)
};
struct alignas(128) DBvhNode {
    double lo0[3], hi0[3], lo1[3], hi1[3];
    int32_t child0, count0, child1, count1;  // count > 0: leaf, child = first triangle slot; count == 0: inner node
    double pad[2];
};
struct alignas(16) DTri {
    double p1[3], e1[3], e2[3];
    int32_t leaf;
    int32_t pad;
};
struct alignas(16) DTriAttr {
    double normal[3];
    int32_t material;
    int32_t xform;  // the mesh's transform (patterns evaluate in the leaf's object space, pattern.rs:99)
};
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
    int32_t program_count;
    int32_t pad;
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
};
struct DStats {
    unsigned long long primary, shadow, reflect, refract;
};

constexpr int kBvhStackDepth = 48;

}  // namespace rtc
