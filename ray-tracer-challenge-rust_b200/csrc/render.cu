// render.cu — the CUDA side of librtc_b200.so (sm_100a, compiled with -fmad=false; see rt_core.cuh for why).
//
// Camera::render's `for y { for x { ... } }` (camera.rs:70-76) is ONE kernel launch: a persistent grid (a fixed number
// of CTAs per SM) whose warps pull 8x4-pixel tiles from an atomic work queue, run World::color_at per lane, quantise
// as canvas.rs:61-63 does and store one uchar4 (and optionally the f64 Canvas colour) per pixel.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "flat_scene.hpp"
#include "lbvh_launch.cuh"
#include "render.cuh"
#include "device_scene_impl.cuh"
#include "render_launch.cuh"
#include "rt_core.cuh"

namespace rtc {
// The smallest instantiation whose feature mask covers the scene's (render_launch.cuh lists them smallest first; the
// last one covers everything).
namespace {
struct RenderInstance {
    int mask;
    RenderLaunchFn fn;
};
RenderInstance pick_instance(int feature_mask) {
    static const RenderInstance kInstances[] = {
#define RTC_TABLE_ENTRY(mask) {mask, launch_render_##mask},
        RTC_RENDER_INSTANCES(RTC_TABLE_ENTRY)
#undef RTC_TABLE_ENTRY
    };
    for (const RenderInstance& inst : kInstances)
        if ((inst.mask & feature_mask) == feature_mask) return inst;
    return RenderInstance{-1, nullptr};
}
}  // namespace
int render_instance_mask(int feature_mask) { return pick_instance(feature_mask).mask; }


using namespace core;

namespace {


__global__ void __launch_bounds__(kBlockThreads) color_at_kernel(const __grid_constant__ DScene s,
                                                                 const double* __restrict__ rays, uint64_t n,
                                                                 double* __restrict__ rgb) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r{v3(rays[6 * i + 0], rays[6 * i + 1], rays[6 * i + 2]), v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
    RayCounters rc;
    Tally tl;
    V3 c = color_at_any(s, r, rc, tl);
    rgb[3 * i + 0] = c.x;
    rgb[3 * i + 1] = c.y;
    rgb[3 * i + 2] = c.z;
}

// FP64 issue-rate probes: 8 independent dependency chains per thread
template <bool kFma>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (kFma) a[k] = __fma_rn(a[k], b, c);
            else a[k] = __dadd_rn(__dmul_rn(a[k], b), c);
        }
    }
    double sum = 0.;
#pragma unroll
    for (int k = 0; k < 8; k++) sum += a[k];
    if (sum == 123.456) out[0] = sum;  // keep the chains alive
}

std::string cuda_err(const char* what, cudaError_t e) {
    return std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
}
#define RTC_CUDA(call)                                         \
    do {                                                       \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) {                               \
            if (err) *err = cuda_err(#call, e_);               \
            return -3;                                         \
        }                                                      \
    } while (0)

template <class T>
size_t slab_bytes(const std::vector<T>& v) {
    return (v.size() * sizeof(T) + 255) & ~size_t(255);
}

}  // namespace


int enable_peer_access(int device, int peer, std::string* err) {
    if (device == peer) return 0;
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(device));
    int can = 0;
    RTC_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) {
        if (err) *err = "devices cannot access each other's memory";
        return -3;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return 0;
    }
    RTC_CUDA(e);
    return 0;
}

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
int frame_share_create(int device, uint64_t bytes, void** d_ptr, unsigned char* handle64, std::string* err) {
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(device));
    void* p = nullptr;
    RTC_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        if (err) *err = cuda_err("cudaIpcGetMemHandle", e);
        return -3;
    }
    std::memcpy(handle64, &h, 64);
    *d_ptr = p;
    return 0;
}
int frame_share_open(int device, const unsigned char* handle64, void** d_ptr, std::string* err) {
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    // opened with the CALLER's device current: the mapping lands in that device's address space and peer access to the
    // owning device is enabled on the way
    RTC_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int frame_share_close(int device, void* d_ptr, int owner, std::string* err) {
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(device));
    if (owner) RTC_CUDA(cudaFree(d_ptr));
    else RTC_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

int cuda_device_count(std::string* err) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        if (err) *err = cuda_err("cudaGetDeviceCount", e);
        return 0;
    }
    return n;
}

// one context per device, created on first use
static DeviceContext* context_for(int device, std::string* err) {
    static std::mutex gmu;
    static DeviceContext* table[64] = {};
    std::lock_guard<std::mutex> lk(gmu);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) {
        if (err) *err = cuda_err("cudaGetDeviceCount(&ndev)", e);
        return nullptr;
    }
    if (device < 0 || device >= ndev || device >= 64) {
        if (err) *err = "no such CUDA device";
        return nullptr;
    }
    if (table[device]) return table[device];
    DeviceContext* c = new DeviceContext();
    c->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->queues, sizeof(DQueue) * kQueueSlots);
    if (e == cudaSuccess) e = cudaMemset(c->queues, 0, sizeof(DQueue) * kQueueSlots);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->shared_slot_free, cudaEventDisableTiming);
    if (e == cudaSuccess) {  // keep freed scene slabs in the stream-ordered pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if (e != cudaSuccess) {
        if (err) *err = cuda_err("creating the device context", e);
        delete c;
        return nullptr;
    }
    table[device] = c;
    return c;
}

int device_scene_create(const FlatScene& f, int device, DeviceScene** out, std::string* err, int* built_depth) {
    DeviceGuard guard_;
    DeviceContext* ctx = context_for(device, err);
    if (!ctx) return -3;
    std::lock_guard<std::mutex> lk(ctx->mu);
    RTC_CUDA(cudaSetDevice(device));
    // device-built meshes (lbvh.cu) own the table entries after the host-built ones: allocated here, never uploaded
    const auto with_extra = [](size_t raw_bytes, size_t extra_bytes) { return (raw_bytes + extra_bytes + 255) & ~(size_t)255; };
    constexpr int kTables = 13;
    const size_t sizes[kTables] = {slab_bytes(f.program), slab_bytes(f.xforms),   slab_bytes(f.prims),
                             slab_bytes(f.gates),   slab_bytes(f.meshes),
                             with_extra(f.bvh.size() * sizeof(DBvhNode), (size_t)f.device_nodes * sizeof(DBvhNode)),
                             with_extra(f.tris.size() * sizeof(DTri), (size_t)f.device_tris * sizeof(DTri)),
                             with_extra(f.tri_attr.size() * sizeof(DTriAttr), (size_t)f.device_tris * sizeof(DTriAttr)),
                             slab_bytes(f.materials), slab_bytes(f.class_offsets), slab_bytes(f.class_members),
                             slab_bytes(f.cluster_entries), slab_bytes(f.tri_smooth)};
    size_t total = 256;
    for (size_t b : sizes) total += b;
    // pinned staging mirrors the slab's layout (one copy when nothing is device-built), followed by the device build's inputs
    const size_t staging_need = total + lbvh_staging_bytes(f) + 256;
    if (ctx->staging_size < staging_need) {
        if (ctx->staging) cudaFreeHost(ctx->staging);
        ctx->staging = nullptr;
        ctx->staging_size = 0;
        RTC_CUDA(cudaHostAlloc(&ctx->staging, staging_need * 2, cudaHostAllocDefault));
        ctx->staging_size = staging_need * 2;
    }
    unsigned char* host = (unsigned char*)ctx->staging;
    size_t off[kTables], at = 0;
    const void* src[kTables] = {f.program.data(), f.xforms.data(),   f.prims.data(),    f.gates.data(), f.meshes.data(),
                          f.bvh.data(),     f.tris.data(),     f.tri_attr.data(), f.materials.data(),
                          f.class_offsets.data(), f.class_members.data(), f.cluster_entries.data(), f.tri_smooth.data()};
    const size_t raw[kTables] = {f.program.size() * sizeof(DProgramNode), f.xforms.size() * sizeof(DXform),
                           f.prims.size() * sizeof(DPrim),          f.gates.size() * sizeof(DGate),
                           f.meshes.size() * sizeof(DMesh),         f.bvh.size() * sizeof(DBvhNode),
                           f.tris.size() * sizeof(DTri),            f.tri_attr.size() * sizeof(DTriAttr),
                           f.materials.size() * sizeof(DMaterial),  f.class_offsets.size() * sizeof(int32_t),
                           f.class_members.size() * sizeof(DClassMember), f.cluster_entries.size() * sizeof(DBox32),
                           f.tri_smooth.size() * sizeof(DTriSmooth)};
    for (int k = 0; k < kTables; k++) {
        off[k] = at;
        if (raw[k]) std::memcpy(host + at, src[k], raw[k]);
        at += sizes[k];
    }
    void* slab = nullptr;
    RTC_CUDA(cudaMallocAsync(&slab, total, ctx->stream));
    cudaError_t e = cudaSuccess;
    if (f.pending.empty()) {
        e = cudaMemcpyAsync(slab, host, at, cudaMemcpyHostToDevice, ctx->stream);
    } else {  // the tails of the bvh / tris / tri_attr tables are written by the device build: copy host-built bytes only
        for (int k = 0; k < kTables && e == cudaSuccess; k++)
            if (raw[k])
                e = cudaMemcpyAsync((unsigned char*)slab + off[k], host + off[k], raw[k], cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess && !f.pending.empty()) {
        unsigned char* base = (unsigned char*)slab;
        int depth = 0;
        bool would_panic = false;
        const int rc = lbvh_build_device(f, host + ((at + 255) & ~(size_t)255), (DBvhNode*)(base + off[5]), (DTri*)(base + off[6]),
                                         (DTriAttr*)(base + off[7]), (DMesh*)(base + off[4]), (DGate*)(base + off[3]),
                                         ctx->stream, &depth, &would_panic, err);
        if (rc != 0) {
            cudaFreeAsync(slab, ctx->stream);
            return rc;
        }
        // a pathological key distribution, or a coordinate the reference panics on: the caller rebuilds on the host
        if (depth + 2 > kBvhStackDepth || would_panic) {
            cudaFreeAsync(slab, ctx->stream);
            if (err) *err = would_panic ? "device gate fold met a non-finite coordinate" : "device-built BVH deeper than the traversal stack";
            return kDeviceBuildTooDeep;
        }
        if (built_depth) *built_depth = depth;
    } else if (e == cudaSuccess) {
        e = cudaStreamSynchronize(ctx->stream);  // the tables are now visible to every stream
    }
    if (e != cudaSuccess) {
        cudaFreeAsync(slab, ctx->stream);
        if (err) *err = cuda_err("uploading the scene", e);
        return -3;
    }
    DeviceScene* s = new DeviceScene();
    s->ctx = ctx;
    s->device = device;
    s->sm_count = ctx->sm_count;
    s->stream = ctx->stream;
    s->slab = slab;
    s->slab_size = total;
    s->upload_bytes = at;
    if (!f.pending.empty()) {
        s->upload_bytes = f.pending_material.size() * (sizeof(rtc_triangle_desc) + sizeof(int32_t));
        for (size_t b : raw) s->upload_bytes += b;
    }
    unsigned char* base = (unsigned char*)s->slab;
    s->view.program = (const DProgramNode*)(base + off[0]);
    s->view.xforms = (const DXform*)(base + off[1]);
    s->view.prims = (const DPrim*)(base + off[2]);
    s->view.gates = (const DGate*)(base + off[3]);
    s->view.meshes = (const DMesh*)(base + off[4]);
    s->view.bvh = (const DBvhNode*)(base + off[5]);
    s->view.tris = (const DTri*)(base + off[6]);
    s->view.tri_attr = (const DTriAttr*)(base + off[7]);
    s->view.materials = (const DMaterial*)(base + off[8]);
    s->view.class_offsets = (const int32_t*)(base + off[9]);
    s->view.class_members = (const DClassMember*)(base + off[10]);
    s->view.cluster_entries = (const DBox32*)(base + off[11]);
    s->view.tri_smooth = f.tri_smooth.empty() ? nullptr : (const DTriSmooth*)(base + off[12]);
    s->view.n_classes = f.class_offsets.empty() ? 0 : (int32_t)f.class_offsets.size() - 1;
    s->view.pad1 = 0;
    s->view.program_count = (int32_t)f.program.size();
    s->view.recursion_limit = f.recursion_limit;
    s->view.n_bvh = (int32_t)f.bvh.size();
    s->view.pad0 = 0;
    s->view.n_prims = (uint32_t)f.prims.size();
    s->view.n_xforms = (uint32_t)f.xforms.size();
    s->view.n_gates = (uint32_t)f.gates.size();
    s->view.n_materials = (uint32_t)f.materials.size();
    s->feature_mask = f.feature_mask;
    s->hot_valid = true;
    for (int k = 0; k < 3; k++) {
        s->hot_lo[k] = f.hot_lo[k];
        s->hot_hi[k] = f.hot_hi[k];
        if (!(f.hot_lo[k] <= f.hot_hi[k]) || !std::isfinite(f.hot_lo[k]) || !std::isfinite(f.hot_hi[k])) s->hot_valid = false;
    }
    if (!f.pending.empty()) s->hot_valid = false;  // a device-built mesh folds its gate box on the GPU: not known here
    for (int k = 0; k < 3; k++) {
        s->view.light_pos[k] = f.light_pos[k];
        s->view.light_int[k] = f.light_int[k];
    }
    *out = s;
    return 0;
}

void device_scene_destroy(DeviceScene* s) {
    if (!s) return;
    {
        std::lock_guard<std::mutex> lk(s->ctx->mu);
        DeviceGuard guard_;
        cudaSetDevice(s->device);
        // stream-ordered free: frames already queued on the library stream finish first.  Frames the caller queued on
        // its OWN stream must be synchronised by the caller before destroying the scene (see rtc_render_device).
        if (s->slab) cudaFreeAsync(s->slab, s->ctx->stream);
    }
    delete s;
}
uint64_t device_scene_bytes(const DeviceScene* s) { return s->slab_size; }
uint64_t device_scene_upload_bytes(const DeviceScene* s) { return s->upload_bytes; }
int device_scene_device(const DeviceScene* s) { return s->device; }

// The work queue a launch on stream `st` uses (DQueue, kQueueSlots).  *shared: the slot is the shared one.
static DQueue* queue_for(DeviceContext* ctx, cudaStream_t st, bool* shared) {
    *shared = false;
    for (int k = 0; k < ctx->slots_taken; k++)
        if (ctx->slot_stream[k] == st) return ctx->queues + k;
    if (ctx->slots_taken < kQueueSlots - 1) {
        ctx->slot_stream[ctx->slots_taken] = st;
        return ctx->queues + ctx->slots_taken++;
    }
    *shared = true;
    return ctx->queues + (kQueueSlots - 1);
}

// Where the scene's bounded geometry lands in this launch's tile grid (DRows.hot_*): its world box through the camera
// (camera.rs:48-65 inverted: pixel = (half - c / -z) / pixel_size - 0.5 for a camera-space point (c, z), z < 0).  Only the
// ORDER of the tile queue depends on it, so anything doubtful — a corner beside or behind the camera plane, a device-built
// mesh, an empty box — simply means "every tile is hot" (the plain row-major order).
static void hot_rectangle(const DeviceScene* s, const DCamera& cam, DRows* rows) {
    const uint32_t tiles_x = (cam.hsize + kTileW - 1) / kTileW, tiles_y = (rows->row_count + kTileH - 1) / kTileH;
    rows->hot_x0 = rows->hot_y0 = 0;
    rows->hot_x1 = tiles_x;
    rows->hot_y1 = tiles_y;
    static const bool off = std::getenv("RTC_B200_NO_HOT_TILES") != nullptr;  // A/B switch
    if (!s->hot_valid || off || tiles_x == 0 || tiles_y == 0) return;
    // camera-to-world is [A | t] (rows 0..2 of transform_inverse): world-to-camera is [A^-1 | -A^-1 t]
    const double* m = cam.inv;
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (!(std::fabs(det) > 1e-300)) return;
    const double r[9] = {(e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det,
                         (f * g - d * i) / det, (a * i - c * g) / det, (c * d - a * f) / det,
                         (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det};
    double px0 = 1e300, px1 = -1e300, py0 = 1e300, py1 = -1e300;
    for (int k = 0; k < 8; k++) {
        const double w[3] = {((k & 1) ? s->hot_hi[0] : s->hot_lo[0]) - m[3], ((k & 2) ? s->hot_hi[1] : s->hot_lo[1]) - m[7],
                             ((k & 4) ? s->hot_hi[2] : s->hot_lo[2]) - m[11]};
        const double cx = r[0] * w[0] + r[1] * w[1] + r[2] * w[2], cy = r[3] * w[0] + r[4] * w[1] + r[5] * w[2],
                     cz = r[6] * w[0] + r[7] * w[1] + r[8] * w[2];
        if (!(cz < -1e-6)) return;  // beside or behind the camera plane: the box's image is not a rectangle
        const double x = (cam.half_width - cx / -cz) / cam.pixel_size - 0.5, y = (cam.half_height - cy / -cz) / cam.pixel_size - 0.5;
        if (!(std::isfinite(x) && std::isfinite(y))) return;
        px0 = std::fmin(px0, x); px1 = std::fmax(px1, x);
        py0 = std::fmin(py0, y); py1 = std::fmax(py1, y);
    }
    px0 = std::floor(px0) - 1.0; py0 = std::floor(py0) - 1.0;
    px1 = std::ceil(px1) + 2.0;  py1 = std::ceil(py1) + 2.0;
    if (px1 <= 0.0 || py1 <= 0.0 || px0 >= (double)cam.hsize || py0 >= (double)cam.vsize) {  // nothing bounded on screen
        rows->hot_x1 = rows->hot_y1 = 0;
        return;
    }
    const double fx0 = std::fmax(px0, 0.0), fx1 = std::fmin(px1, (double)cam.hsize);
    double fy0 = std::fmax(py0, 0.0), fy1 = std::fmin(py1, (double)cam.vsize);
    // frame rows -> this launch's local rows: local band k holds frame band band_first + k * band_stride
    const double br = (double)rows->band_rows, st = (double)rows->band_stride;
    double ly0 = std::floor((std::floor(fy0 / br) - (double)rows->band_first) / st) * br;
    double ly1 = (std::floor((std::floor(fy1 / br) - (double)rows->band_first) / st) + 1.0) * br;
    if (rows->band_stride == 1 && rows->band_first == 0) {  // contiguous rows: exact
        ly0 = fy0;
        ly1 = fy1;
    }
    ly0 = std::fmax(ly0 - (double)rows->row_begin, 0.0);
    ly1 = std::fmin(std::fmax(ly1 - (double)rows->row_begin, 0.0), (double)rows->row_count);
    if (!(ly1 > ly0)) {
        rows->hot_x1 = rows->hot_y1 = 0;
        return;
    }
    rows->hot_x0 = (uint32_t)(fx0 / kTileW);
    rows->hot_x1 = std::min<uint32_t>(tiles_x, (uint32_t)((fx1 + kTileW - 1) / kTileW));
    rows->hot_y0 = (uint32_t)(ly0 / kTileH);
    rows->hot_y1 = std::min<uint32_t>(tiles_y, (uint32_t)((ly1 + kTileH - 1) / kTileH));
    if (rows->hot_x1 <= rows->hot_x0 || rows->hot_y1 <= rows->hot_y0) rows->hot_x1 = rows->hot_y1 = rows->hot_x0 = rows->hot_y0 = 0;
}

// Enqueues one render kernel on `st`.  timed: bracket it with the context's events (one timed launch in flight per device).
static int launch(DeviceScene* s, const DCamera& cam, const DRows& rows, void* d8, void* d64, cudaStream_t st, bool timed,
                  DQueue** used, std::string* err) {
    if (used) *used = nullptr;
    if (rows.row_count == 0 || cam.hsize == 0) return 0;
    DeviceContext* ctx = s->ctx;
    bool shared = false;
    DQueue* queue = queue_for(ctx, st, &shared);
    if (shared) RTC_CUDA(cudaStreamWaitEvent(st, ctx->shared_slot_free, 0));
    const uint64_t tiles = (uint64_t)((cam.hsize + kTileW - 1) / kTileW) * ((rows.row_count + kTileH - 1) / kTileH);
    const uint64_t warps_per_block = kBlockThreads / 32;
    uint64_t blocks = (tiles + warps_per_block - 1) / warps_per_block;
    const uint64_t cap = (uint64_t)s->sm_count * blocks_per_sm_for(pick_instance(s->feature_mask).mask);
    if (blocks > cap) blocks = cap;
    // (Launching fewer CTAs for a small slice — a minimum number of tiles per warp, or exactly ceil(tiles / resident warps)
    // tiles for every warp — was measured and changes nothing: one rank's eighth of the 1080p table frame takes 0.121 ms with
    // any of them, against 0.574 / 8 = 0.072 ms; every slice costs about 0.05 ms on top of its share, the time the last tiles
    // (32 pixels x up to six dependent walks) take to drain.  profiles/r02u_slice_sweep.md)
    if (timed) RTC_CUDA(cudaEventRecord(ctx->ev0, st));
    const RenderLaunchFn fn = pick_instance(s->feature_mask).fn;
    DRows ordered = rows;
    hot_rectangle(s, cam, &ordered);
    fn((unsigned)blocks, st, s->view, cam, ordered, (uint32_t*)d8, (double*)d64, queue);
    RTC_CUDA(cudaGetLastError());
    if (timed) RTC_CUDA(cudaEventRecord(ctx->ev1, st));
    if (shared) RTC_CUDA(cudaEventRecord(ctx->shared_slot_free, st));
    if (used) *used = queue;
    return 0;
}

// Waits for a timed launch and reads its counters and kernel time.
static int launch_stats(DeviceScene* s, DQueue* queue, cudaStream_t st, LaunchStats* stats, std::string* err) {
    *stats = LaunchStats{};
    if (!queue) return 0;
    unsigned long long h[4];
    RTC_CUDA(cudaMemcpyAsync(h, queue->result, sizeof(h), cudaMemcpyDeviceToHost, st));
    RTC_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    RTC_CUDA(cudaEventElapsedTime(&ms, s->ctx->ev0, s->ctx->ev1));
    stats->primary = h[0];
    stats->shadow = h[1];
    stats->reflect = h[2];
    stats->refract = h[3];
    stats->launches = 1;
    stats->device_ms = ms;
    return 0;
}

int render_device(DeviceScene* s, const DCamera& cam, const DRows& rows, void* d_rgba8, void* d_rgb_f64, void* stream,
                  LaunchStats* stats, std::string* err) {
    std::lock_guard<std::mutex> lk(s->mu());
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(s->device));
    DQueue* q = nullptr;
    int rc = launch(s, cam, rows, d_rgba8, d_rgb_f64, (cudaStream_t)stream, stats != nullptr, &q, err);
    if (rc == 0 && stats) rc = launch_stats(s, q, (cudaStream_t)stream, stats, err);
    return rc;
}

// Two halves of render_device for callers that keep several devices busy at once (multi.cu): enqueue on every device
// first, then collect.
int render_device_begin(DeviceScene* s, const DCamera& cam, const DRows& rows, void* d_rgba8, void* d_rgb_f64, void* stream,
                        void** token, std::string* err) {
    std::lock_guard<std::mutex> lk(s->mu());
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(s->device));
    DQueue* q = nullptr;
    const int rc = launch(s, cam, rows, d_rgba8, d_rgb_f64, stream ? (cudaStream_t)stream : s->stream, true, &q, err);
    *token = q;
    return rc;
}
int render_device_end(DeviceScene* s, void* stream, void* token, LaunchStats* stats, std::string* err) {
    std::lock_guard<std::mutex> lk(s->mu());
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(s->device));
    LaunchStats local;
    return launch_stats(s, (DQueue*)token, stream ? (cudaStream_t)stream : s->stream, stats ? stats : &local, err);
}
void* device_scene_stream(const DeviceScene* s) { return s->stream; }

// Local rows [r0, r1) of a compact device buffer -> the host buffer.  Compact host buffer: the same offsets.  Frame layout
// (the host pointer addresses the WHOLE frame): local band j lands at frame band band_first + j * band_stride — one
// strided copy for the whole bands of the range, one more for the frame's ragged last band; r0 is a multiple of band_rows.
static int copy_rows_out(const DCamera& cam, const DRows& rows, bool to_frame, unsigned char* host, const unsigned char* dev,
                         size_t bytes_per_pixel, uint32_t r0, uint32_t r1, cudaStream_t st, std::string* err) {
    if (r1 <= r0 || cam.hsize == 0) return 0;
    const size_t row_bytes = (size_t)cam.hsize * bytes_per_pixel;
    if (!to_frame) {
        RTC_CUDA(cudaMemcpyAsync(host + r0 * row_bytes, dev + r0 * row_bytes, (size_t)(r1 - r0) * row_bytes,
                                 cudaMemcpyDeviceToHost, st));
        return 0;
    }
    const size_t band_bytes = row_bytes * rows.band_rows;
    const uint32_t b0 = r0 / rows.band_rows, full = (r1 - r0) / rows.band_rows;
    const uint32_t rest = (r1 - r0) - full * rows.band_rows;
    const auto frame_band = [&](uint32_t local_band) {
        return host + ((size_t)rows.band_first + (size_t)local_band * rows.band_stride) * band_bytes;
    };
    if (full)
        RTC_CUDA(cudaMemcpy2DAsync(frame_band(b0), band_bytes * rows.band_stride, dev + r0 * row_bytes, band_bytes, band_bytes,
                                   full, cudaMemcpyDeviceToHost, st));
    if (rest)
        RTC_CUDA(cudaMemcpyAsync(frame_band(b0 + full), dev + ((size_t)r0 + (size_t)full * rows.band_rows) * row_bytes,
                                 (size_t)rest * row_bytes, cudaMemcpyDeviceToHost, st));
    return 0;
}

int render_host(DeviceScene* s, const DCamera& cam, const DRows& rows_in, uint8_t* rgba8, double* rgb_f64,
                LaunchStats* stats, std::string* err) {
    std::lock_guard<std::mutex> lk(s->mu());
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(s->device));
    // RTC_ROWS_FRAME with host buffers: the kernel still renders into COMPACT device buffers; the copies place each band
    // at its frame position in the host frame (e.g. a canvas in host memory shared by one process per GPU: every rank's
    // copy engine writes its own bands over its own PCIe link).  The whole frame is its own compact form.
    const bool to_frame = rows_in.frame_layout != 0 && !(rows_in.band_first == 0 && rows_in.band_stride == 1);
    DRows rows = rows_in;
    rows.frame_layout = 0;
    const size_t px = (size_t)rows.local_rows * cam.hsize;
    if (rgba8 && s->ctx->out8_size < px * 4) {
        if (s->ctx->out8) cudaFree(s->ctx->out8);
        s->ctx->out8 = nullptr;
        s->ctx->out8_size = 0;
        RTC_CUDA(cudaMalloc(&s->ctx->out8, px * 4));
        s->ctx->out8_size = px * 4;
    }
    if (rgb_f64 && s->ctx->out64_size < px * 24) {
        if (s->ctx->out64) cudaFree(s->ctx->out64);
        s->ctx->out64 = nullptr;
        s->ctx->out64_size = 0;
        RTC_CUDA(cudaMalloc(&s->ctx->out64, px * 24));
        s->ctx->out64_size = px * 24;
    }
    // Overlap the device->host copies with rendering: the call's rows are rendered in a few launches and each chunk's
    // pixels cross PCIe on a second stream while the next chunk renders.
    //   RGBA8 only (4 B/px): two launches, 3/4 then 1/4 — more, smaller chunks were measured slower at every frame size
    //   (each launch pays its own ramp-up and tail: 8K pumpkin 11.5 ms with 2 chunks, 11.7-12.2 ms with 4-16;
    //   profiles/r01n_host_chunk_sweep.json);
    //   with the f64 Canvas colours (24 B/px more: the copy, not the kernel, is the long pole — 50 MB at 1080p) four or
    //   eight launches (see the weights below): the copy engine starts after a fraction of the frame and hardly waits.
    // Ray counters are per launch; with `stats` requested the frame is rendered in one launch so that the reported kernel
    // time is one kernel's.
    const bool split = !stats && (rgba8 || rgb_f64) && rows.local_rows >= 256;
    DQueue* q = nullptr;
    if (!split) {
        int rc = launch(s, cam, rows, rgba8 ? s->ctx->out8 : nullptr, rgb_f64 ? s->ctx->out64 : nullptr, s->stream,
                        stats != nullptr, &q, err);
        if (rc) return rc;
        if (rgba8 && (rc = copy_rows_out(cam, rows, to_frame, rgba8, (const unsigned char*)s->ctx->out8, 4, 0, rows.local_rows,
                                         s->stream, err)))
            return rc;
        if (rgb_f64 && (rc = copy_rows_out(cam, rows, to_frame, (unsigned char*)rgb_f64, (const unsigned char*)s->ctx->out64, 24,
                                           0, rows.local_rows, s->stream, err)))
            return rc;
    } else {
        DeviceContext* ctx = s->ctx;
        if (!ctx->copy_stream) {
            RTC_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            for (cudaEvent_t& e : ctx->chunk_done) RTC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            RTC_CUDA(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
        }
        uint32_t cut[kHostChunks + 1];  // chunk k renders local rows [cut[k], cut[k + 1])
        int nchunks;
        // chunks start on a tile row, and for a frame-layout copy on a band as well
        uint32_t unit = kTileH;
        if (to_frame)
            for (unit = rows.band_rows; unit % kTileH; unit += rows.band_rows) {}
        const auto tile_rows = [&](uint32_t r) { return std::min(rows.local_rows, (r + unit - 1) / unit * unit); };
        if (rgb_f64) {
            // Relative sizes of the chunks.  The copy (24 B/px at ~55 GB/s) is slower than the rendering, so the call costs
            // about  first chunk's render + the whole copy + the time the copy engine starves: a small first chunk, then
            // chunks small enough that the next one is rendered before the last one has crossed, and a small last chunk
            // (its copy is the only one nothing overlaps).  Eight chunks {1,2,2,2,2,2,2,1} for calls of >= 2 Mpx: the
            // drop-in call on the 1080p table frame 1.20 -> 1.10 ms, 8K pumpkin 20.3 -> 16.8 ms, teapot and cow & teddy
            // unchanged (their device mesh build and marshalling dominate); smaller calls — a rank's share of a sharded
            // frame — keep four {1,1,2,4}: every launch pays its own ~0.05 ms drain.  profiles/r02zo_chunk_sweep.json;
            // RTC_B200_F64_CHUNKS="w0,w1,..." (up to kHostChunks positive integers) overrides both (tools/chunk_sweep.py).
            static const std::vector<uint32_t> env_weights = [] {
                std::vector<uint32_t> w;
                if (const char* env = std::getenv("RTC_B200_F64_CHUNKS")) {
                    for (const char* p = env; *p && (int)w.size() < kHostChunks;) {
                        char* end = nullptr;
                        const long v = std::strtol(p, &end, 10);
                        if (end == p) break;
                        if (v > 0) w.push_back((uint32_t)v);
                        p = (*end == ',') ? end + 1 : end;
                        if (*end != ',' ) break;
                    }
                }
                return w;
            }();
            static const std::vector<uint32_t> small_call = {1, 1, 2, 4}, large_call = {1, 2, 2, 2, 2, 2, 2, 1};
            const std::vector<uint32_t>& weights = !env_weights.empty() ? env_weights : px >= 2000000 ? large_call : small_call;
            nchunks = (int)weights.size();
            uint64_t total = 0, run = 0;
            for (uint32_t v : weights) total += v;
            cut[0] = 0;
            for (int k = 0; k < nchunks; k++) {
                run += weights[k];
                cut[k + 1] = k + 1 == nchunks ? rows.local_rows : tile_rows((uint32_t)((uint64_t)rows.local_rows * run / total));
            }
        } else {
            nchunks = 2;
            cut[0] = 0;
            cut[1] = tile_rows(rows.local_rows * 3 / 4);
            cut[2] = rows.local_rows;
        }
        for (int k = 0; k < nchunks; k++) {
            DRows part = rows;
            part.row_begin = cut[k];
            part.row_count = cut[k + 1] - cut[k];
            int rc = launch(s, cam, part, rgba8 ? ctx->out8 : nullptr, rgb_f64 ? ctx->out64 : nullptr, s->stream, false,
                            nullptr, err);
            if (rc) return rc;
            RTC_CUDA(cudaEventRecord(ctx->chunk_done[k], s->stream));
            RTC_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_done[k], 0));
            if (rgba8 && (rc = copy_rows_out(cam, rows, to_frame, rgba8, (const unsigned char*)ctx->out8, 4, cut[k], cut[k + 1],
                                             ctx->copy_stream, err)))
                return rc;
            if (rgb_f64 && (rc = copy_rows_out(cam, rows, to_frame, (unsigned char*)rgb_f64, (const unsigned char*)ctx->out64,
                                               24, cut[k], cut[k + 1], ctx->copy_stream, err)))
                return rc;
        }
        RTC_CUDA(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
        RTC_CUDA(cudaStreamWaitEvent(s->stream, ctx->copy_done, 0));
    }
    RTC_CUDA(cudaStreamSynchronize(s->stream));
    if (stats) return launch_stats(s, q, s->stream, stats, err);
    return 0;
}

int color_at_host(DeviceScene* s, const double* rays, uint64_t n, double* rgb, std::string* err) {
    std::lock_guard<std::mutex> lk(s->mu());
    if (n == 0) return 0;
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(s->device));
    double *d_rays = nullptr, *d_rgb = nullptr;
    RTC_CUDA(cudaMalloc((void**)&d_rays, n * 48));
    cudaError_t e = cudaMalloc((void**)&d_rgb, n * 24);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rays, rays, n * 48, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) {
        color_at_kernel<<<(unsigned)((n + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, s->stream>>>(
            s->view, d_rays, n, d_rgb);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(rgb, d_rgb, n * 24, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_rays);
    if (d_rgb) cudaFree(d_rgb);
    if (e != cudaSuccess) {
        if (err) *err = cuda_err("rtc_color_at", e);
        return -3;
    }
    return 0;
}

int measure_fp64_peak(int device, double* nofma_gflops, double* fma_gflops, std::string* err) {
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RTC_CUDA(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    RTC_CUDA(cudaMalloc((void**)&d, 64));
    cudaEvent_t a, b;
    RTC_CUDA(cudaEventCreate(&a));
    RTC_CUDA(cudaEventCreate(&b));
    const int iters = 4096, threads = 256, blocks = prop.multiProcessorCount * 8;
    const double flops = 2.0 * 8.0 * iters * (double)threads * blocks;  // mul + add (or one FMA = 2) per chain step
    double best[2] = {0., 0.};
    for (int mode = 0; mode < 2; mode++)
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(a);
            if (mode == 0) fp64_peak_kernel<false><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            else fp64_peak_kernel<true><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
            cudaEventRecord(b);
            cudaError_t e = cudaEventSynchronize(b);
            if (e != cudaSuccess) {
                if (err) *err = cuda_err("fp64 probe", e);
                cudaFree(d);
                return -3;
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, a, b);
            if (rep > 0 && ms > 0.f) best[mode] = std::max(best[mode], flops / (ms * 1e-3) / 1e9);
        }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    if (nofma_gflops) *nofma_gflops = best[0];
    if (fma_gflops) *fma_gflops = best[1];
    return 0;
}

// ---- stream-ordered counters (render.cuh) ------------------------------------------------------------------------------
namespace {
struct CounterList {
    unsigned int* p[16];
};
__global__ void counters_store_kernel(CounterList c, uint32_t n, uint32_t value, int add) {
    const unsigned k = threadIdx.x;
    if (k >= n) return;
    __threadfence_system();
    if (add) atomicAdd_system(c.p[k], 1u);
    else atomicExch_system(c.p[k], value);
}
__global__ void counter_poll_kernel(const volatile unsigned int* c, uint32_t at_least) {
    while (*c < at_least) __nanosleep(200);
    __threadfence_system();
}
// cuStreamWaitValue32 through the runtime's driver entry point: librtc_b200.so links no libcuda (it must load on a box
// without a driver, where only the host API is used)
typedef int (*StreamWaitValue32Fn)(void* stream, unsigned long long addr, uint32_t value, unsigned int flags);
StreamWaitValue32Fn stream_wait_value32() {
    static StreamWaitValue32Fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        static const bool off = std::getenv("RTC_B200_NO_WAIT_VALUE") != nullptr;  // A/B switch: force the polling kernel
        return off ? nullptr : (StreamWaitValue32Fn)p;
    }();
    return fn;
}
}  // namespace
int stream_counter_wait(int device, void* stream, void* d_counter, uint32_t at_least, std::string* err) {
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(device));
    if (StreamWaitValue32Fn fn = stream_wait_value32()) {
        const int rc = fn(stream, (unsigned long long)(uintptr_t)d_counter, at_least, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
        if (rc == 0) return 0;
        // (e.g. CUDA_ERROR_NOT_SUPPORTED for this address) fall through to the polling kernel
    }
    counter_poll_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const volatile unsigned int*)d_counter, at_least);
    RTC_CUDA(cudaGetLastError());
    return 0;
}
static int counters_store(int device, void* stream, void* const* d_counters, uint32_t n, uint32_t value, int add,
                          std::string* err) {
    if (n == 0) return 0;
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(device));
    CounterList c{};
    for (uint32_t k = 0; k < n && k < 16; k++) c.p[k] = (unsigned int*)d_counters[k];
    counters_store_kernel<<<1, 16, 0, (cudaStream_t)stream>>>(c, n, value, add);
    RTC_CUDA(cudaGetLastError());
    return 0;
}
int stream_counters_set(int device, void* stream, void* const* d_counters, uint32_t n, uint32_t value, std::string* err) {
    return counters_store(device, stream, d_counters, n, value, 0, err);
}
int stream_counters_add(int device, void* stream, void* const* d_counters, uint32_t n, std::string* err) {
    return counters_store(device, stream, d_counters, n, 0, 1, err);
}

// Page-locks host memory the caller mapped itself (a POSIX shared-memory segment: capi.cpp rtc_host_share_*), for every
// CUDA context of the process, so copies into it are asynchronous DMA at PCIe speed.
int host_register(int device, void* p, size_t bytes, std::string* err) {
    DeviceGuard guard_;
    RTC_CUDA(cudaSetDevice(device));
    RTC_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return 0;
}
int host_unregister(void* p, std::string* err) {
    RTC_CUDA(cudaHostUnregister(p));
    return 0;
}

void* pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {  // every device's copy engine may write it
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // namespace rtc
