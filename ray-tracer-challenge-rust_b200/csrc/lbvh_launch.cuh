// lbvh_launch.cuh — host-callable entry of the device mesh build (lbvh.cu); internal to librtc_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "device_scene.h"

namespace rtc {

struct FlatScene;

// pinned staging bytes the build needs for its inputs
size_t lbvh_staging_bytes(const FlatScene& f);
// Builds every FlatScene::pending mesh into the scene tables (device pointers to table starts), in stream order; returns
// after the stream has drained with the deepest tree's depth (counted as bvh.hpp does).  0 or -3 (*err set).
// *would_panic: a gate fold met a coordinate the reference's Bounds::add would panic on (the caller rebuilds on the host,
// which raises it).
int lbvh_build_device(const FlatScene& f, unsigned char* pinned, DBvhNode* d_nodes, DTri* d_tris, DTriAttr* d_attr,
                      DMesh* d_meshes, DGate* d_gates, cudaStream_t st, int* max_depth, bool* would_panic,
                      std::string* err);

}  // namespace rtc
