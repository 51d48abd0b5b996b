/* include/rtc.h — C ABI of librtc_b200.so, the B200 (sm_100a) render path for the Ray-Tracer-Challenge renderer
 * antoinehebert/ray-tracer-challenge-rust.
 *
 * The reference has no FFI or plugin layer; its seam for this path is the pure function
 *     pub fn render(&self, world: &World) -> Canvas            (src/camera.rs:67-79)
 * whose body is `for y {for x { world.color_at(&self.ray_for_pixel(x, y)) }}`.  This header is what a thin Rust
 * `rtc-sys` crate would bind (INTEGRATION.md shows the `extern "C"` block and the patched Camera::render): plain
 * pointers and sizes, no C++ or torch types.  All `file:line` citations are relative to /root/reference/.
 *
 * Two layers:
 *   1. CORE BOUNDARY  (rtc_scene_*, rtc_render*, rtc_color_at) — replaces camera.rs:67-79 + world.rs:80-82.  The caller
 *      hands over the World exactly as its own host code computed it (transforms with their cached inverses,
 *      triangle edges and normals, materials with pattern inverses), so no value that reaches a pixel is recomputed
 *      with a different rounding on this side of the boundary.
 *   2. HOST MIRROR    (rtc_shape_*, rtc_world_*, rtc_camera_*, rtc_canvas_*, rtc_obj_*, rtc_mat_*) — the reference's own
 *      host API (Shape/World/Camera/Canvas/Parser/transformations) restated in C++ behind C entry points, because no
 *      Rust toolchain exists in this image.  It is what tests/ and bench.py drive; rtc_camera_render() marshals its
 *      World into layer 1 exactly as the Rust Camera::render patch would.
 *
 * Errors: the reference panics (expect/panic!/assert!); here every fallible call returns 0 on success or a negative
 * code and leaves a thread-local message for rtc_last_error().  Nothing aborts.  There is NO CPU fallback: without a
 * usable CUDA device every rendering call fails with RTC_ERR_CUDA.
 */
#ifndef RTC_H
#define RTC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTC_OK 0
#define RTC_ERR_INVALID (-1)  /* bad argument / malformed description */
#define RTC_ERR_PANIC (-2)    /* the reference would have panicked (message says where: file:line) */
#define RTC_ERR_CUDA (-3)     /* CUDA runtime failure or no device */
#define RTC_ERR_UNSUPPORTED (-4)
#define RTC_ERR_TIMEOUT (-6)  /* rtc_host_counter_wait: the counter did not get there in time */

/* ShapeKind (src/shape.rs:14-39) */
enum { RTC_SPHERE = 0, RTC_PLANE = 1, RTC_CUBE = 2, RTC_CYLINDER = 3, RTC_CONE = 4, RTC_GROUP = 5, RTC_TRIANGLE = 6,
       /* not in the reference (its smooth-triangle scenarios are commented out, src/intersection.rs:381-386,
        * src/obj_file.rs:295-335): the book's SmoothTriangle — a triangle whose normal is interpolated from three vertex
        * normals with the u, v of the hit */
       RTC_SMOOTH_TRIANGLE = 7 };
/* PatternKind (src/pattern.rs:4-12) */
enum { RTC_PATTERN_NONE = -1, RTC_PATTERN_STRIPE = 0, RTC_PATTERN_GRADIENT = 1, RTC_PATTERN_RING = 2,
       RTC_PATTERN_CHECKERS = 3, RTC_PATTERN_TEST = 4 };

/* Material (src/material.rs:4-14) with its Option<Pattern> (src/pattern.rs:14-19) flattened in.  Matrices are
 * row-major 4x4.  pattern_inverse is Pattern.transform_inverse as the caller cached it (pattern.rs:63-66). */
typedef struct rtc_material {
    double color[3];
    double ambient, diffuse, specular, shininess, reflective, transparency, refractive_index;
    int32_t pattern_kind; /* RTC_PATTERN_* */
    int32_t _pad;
    double pattern_a[3], pattern_b[3];
    double pattern_transform[16], pattern_inverse[16];
} rtc_material;

/* ------------------------------------------------------------------------------------------------------------------
 * 1. CORE BOUNDARY
 * ---------------------------------------------------------------------------------------------------------------- */

/* Shape.transform and Shape.transform_inverse (src/shape.rs:44-45), row-major. */
typedef struct rtc_transform_desc {
    double transform[16];
    double inverse[16];
} rtc_transform_desc;

/* ShapeKind::Triangle payload (src/shape.rs:31-38) as Shape::triangle computed it (src/shape.rs:171-193). */
typedef struct rtc_triangle_desc {
    double p1[3], p2[3], p3[3], e1[3], e2[3], normal[3];
} rtc_triangle_desc;

/* Vertex normals of an RTC_SMOOTH_TRIANGLE (the book's smooth_triangle(p1, p2, p3, n1, n2, n3)). */
typedef struct rtc_vertex_normals {
    double n1[3], n2[3], n3[3];
} rtc_vertex_normals;

/* One Shape.  World.objects is sent as a pre-order walk: a group is followed by its `child_count` children, each
 * followed by its own subtree (src/shape.rs:28-30). */
typedef struct rtc_shape_desc {
    int32_t kind;        /* RTC_SPHERE .. RTC_SMOOTH_TRIANGLE */
    int32_t material;    /* leaves: index into materials[] */
    int32_t transform;   /* index into transforms[] (groups: their own, always identity in the reference) */
    int32_t capped;      /* cylinder / cone */
    double minimum, maximum;
    int32_t child_count; /* groups only */
    int32_t triangle;    /* triangles: index into triangles[] (and vertex_normals[] for smooth triangles) */
} rtc_shape_desc;

/* World (src/world.rs:13-16) + Light (src/light.rs:5-8). */
typedef struct rtc_scene_desc {
    const rtc_shape_desc* shapes;
    uint32_t shape_count;
    uint32_t root_count; /* World.objects.len() */
    const rtc_transform_desc* transforms;
    uint32_t transform_count;
    const rtc_material* materials;
    uint32_t material_count;
    const rtc_triangle_desc* triangles;
    uint32_t triangle_count;
    double light_position[3];
    double light_intensity[3];
    /* n1, n2, n3 of the RTC_SMOOTH_TRIANGLE shapes, indexed like triangles[] (triangle_count entries); may be NULL when
     * the world holds no smooth triangle */
    const rtc_vertex_normals* vertex_normals;
    /* World's RECURSION_LIMIT (src/world.rs:11).  0 means the reference's value, 5.  The budget is spent three units per
     * bounce (world.rs:95, :68-69, :126/:159), so 2-3 render surfaces only, 5-6 one bounce, 8-9 two bounces, ...; values
     * with limit % 3 == 1 underflow the reference's usize arithmetic (world.rs:68) and are refused (RTC_ERR_PANIC). */
    uint32_t recursion_limit;
    uint32_t _reserved;
} rtc_scene_desc;

/* Camera (src/camera.rs:5-12): the values Camera::new / set_transform computed (camera.rs:16-46). */
typedef struct rtc_camera_desc {
    uint32_t hsize, vsize;
    double inverse[16]; /* transform_inverse */
    double half_width, half_height, pixel_size;
} rtc_camera_desc;

/* Which rows of the frame a call renders.  Rows are grouped in bands of `band_rows`; the call renders bands
 * band_first, band_first + band_stride, ... (row-cyclic sharding over GPUs: band_first = rank, band_stride = world
 * size).  The output is COMPACT: local band k holds frame rows [ (band_first + k*band_stride) * band_rows, ... ).
 * {band_rows = vsize, band_first = 0, band_stride = 1} (or a NULL pointer) is the whole frame. */
typedef struct rtc_rows {
    uint32_t band_rows, band_first, band_stride;
    uint32_t layout; /* RTC_ROWS_COMPACT: the output buffers hold only this call's rows, packed;
                        RTC_ROWS_FRAME: the output pointers address the WHOLE frame (vsize rows) and each rendered row is
                        written at its frame position — e.g. a frame buffer on another GPU mapped over NVLink, so the
                        ranks of a sharded render store straight into rank 0's frame and no gather step is needed */
} rtc_rows;
#define RTC_ROWS_COMPACT 0u
#define RTC_ROWS_FRAME 1u

/* Work counters of one render (exact; the numerators of Mrays/s). */
typedef struct rtc_stats {
    uint64_t primary_rays; /* pixels rendered (camera.rs:72) */
    uint64_t shadow_rays;  /* is_shadowed calls = shade_hit calls (world.rs:65) */
    uint64_t reflect_rays; /* world.rs:125 */
    uint64_t refract_rays; /* world.rs:155 */
    uint64_t kernel_launches;
    double device_ms;      /* CUDA-event time of the render kernel(s) on the launch stream */
} rtc_stats;

typedef struct rtc_scene rtc_scene;

/* Validates the description, computes the group gate boxes exactly as Bounds::new would per ray (src/bounds.rs:11-140),
 * flattens the tree into device tables, builds the per-mesh BVH and uploads everything once to `device`.
 * Fails with RTC_ERR_PANIC where the reference would panic while rendering this world (e.g. an uncapped cylinder inside
 * a group: bounds.rs:143), RTC_ERR_UNSUPPORTED for non-affine transforms. */
int rtc_scene_create(const rtc_scene_desc* desc, int device, rtc_scene** out);
/* The same with a choice of mesh build.  Both produce the same pixels (a BVH only decides which exact triangle tests
 * run); they trade build time against traversal work:
 *   RTC_BUILD_HOST_SAH    (rtc_scene_create) binned-SAH BVH built on the host: milliseconds per 10 k triangles, the fewest
 *                         box tests per ray — for scenes that are rendered many times;
 *   RTC_BUILD_DEVICE_LBVH meshes of >= 256 triangles are built on the GPU (Morton order, Karras hierarchy, csrc/lbvh.cuh):
 *                         triangle normals, triangle tables and BVH never touch the host; 8-20 % more box tests per ray
 *                         — for scenes that are built, rendered once and dropped (obj_file.rs -> Camera::render). */
#define RTC_BUILD_HOST_SAH 0u
#define RTC_BUILD_DEVICE_LBVH 1u
int rtc_scene_create_ex(const rtc_scene_desc* desc, int device, uint32_t flags, rtc_scene** out);
void rtc_scene_destroy(rtc_scene* scene);
/* Flattened-scene facts for reports: n[0]=leaves, [1]=gates, [2]=meshes, [3]=mesh triangles, [4]=bvh nodes,
 * [5]=device bytes. */
int rtc_scene_info(const rtc_scene* scene, uint64_t n[6]);
/* Host->device bytes rtc_scene_create(_ex) copied for this scene (tables; for a device build the triangle inputs). */
uint64_t rtc_scene_upload_bytes(const rtc_scene* scene);

/* Camera::render (src/camera.rs:67-79) with HOST output buffers (either may be NULL):
 *   rgba8_out  : rows*hsize*4 bytes, each channel quantised as canvas.rs:61-63 does at PPM time, alpha = 255
 *   rgb_f64_out: rows*hsize*3 doubles, the Canvas colours themselves (canvas.rs:24-26)
 * Includes the device->host copies (rendered in a few launches so that the copies overlap the rendering).  `rows` NULL =
 * whole frame.  With RTC_ROWS_COMPACT the buffers hold only this call's rows, packed; with RTC_ROWS_FRAME they are the WHOLE
 * frame (vsize rows) and each rendered band is copied to its frame position, the other rows are left alone — what one rank
 * of a sharded render passes when the frame lives in host memory shared by the ranks (rtc_host_share_*). */
int rtc_render(const rtc_scene* scene, const rtc_camera_desc* camera, const rtc_rows* rows, uint8_t* rgba8_out,
               double* rgb_f64_out, rtc_stats* stats);

/* Same, DEVICE output buffers on the scene's device, asynchronous on `cuda_stream` (a cudaStream_t, 0 = default
 * stream).  `stats` (nullable) is filled only if `sync_stats` is non-zero, which synchronises the stream. */
int rtc_render_device(const rtc_scene* scene, const rtc_camera_desc* camera, const rtc_rows* rows,
                      void* d_rgba8_out, void* d_rgb_f64_out, void* cuda_stream, int sync_stats, rtc_stats* stats);

/* Sharded renders across PROCESSES (one per GPU) without a collective on the frame's path.  rtc_render_device_notify is
 * rtc_render_device (asynchronous, no stats) whose launch adds 1, with system scope, to the uint32 at d_counter once every
 * pixel it stored is visible system-wide — d_counter may live in another GPU's memory mapped into this device
 * (rtc_frame_share_open), e.g. next to the frame the kernels store into (RTC_ROWS_FRAME).  The frame's owner makes its
 * stream wait for the counter to reach the number of contributions it expects (rtc_stream_wait_counter: a stream wait-value
 * operation, no kernel occupies the GPU while waiting) and later tells the other ranks that a frame buffer may be written
 * again by setting a counter in each rank's own memory (rtc_stream_set_counters, up to 16 per call). */
int rtc_render_device_notify(const rtc_scene* scene, const rtc_camera_desc* camera, const rtc_rows* rows,
                             void* d_rgba8_out, void* d_rgb_f64_out, void* cuda_stream, void* d_counter);
int rtc_stream_wait_counter(int device, void* cuda_stream, void* d_counter, uint32_t at_least);
int rtc_stream_set_counters(int device, void* cuda_stream, void** d_counters, uint32_t n, uint32_t value);

/* ONE frame sharded over the devices 0 .. ngpus-1 of this process — no torch, no NCCL, plain CUDA: the scene is flattened
 * once and uploaded to every device, the frame's 8-row bands are dealt cyclically to the devices (band b -> device b mod
 * ngpus), every device renders its bands with one launch, and the frame is completed in one of two places:
 *   RTC_MULTI_HOST_FRAME    one pinned host frame: every device copies its own bands into it with one strided copy over
 *                           its own PCIe link, all links in parallel (rtc_multi_host_frame; also copied to rgba8_out when
 *                           that is not NULL)
 *   RTC_MULTI_DEVICE_FRAME  a frame in device 0's memory: the kernels of the other devices store their pixels straight
 *                           into it over NVLink peer mappings — no gather step (rtc_multi_device_frame; rgba8_out must
 *                           be NULL)
 * stats: ray counts summed over the devices, kernel_launches = ngpus, device_ms = the slowest device's kernel.
 * rtc_render_multi = create + render (host frame into rgba8_out) + destroy: the sharded Camera::render of a World that is
 * rendered once. */
typedef struct rtc_multi rtc_multi;
#define RTC_MULTI_HOST_FRAME 0u
#define RTC_MULTI_DEVICE_FRAME 1u
int rtc_multi_create(const rtc_scene_desc* desc, int ngpus, uint32_t build_flags, rtc_multi** out);
int rtc_multi_render(rtc_multi* m, const rtc_camera_desc* camera, uint32_t where, uint8_t* rgba8_out, rtc_stats* stats);
/* rtc_render sharded over the devices of `m`, into CALLER host buffers (either may be NULL): rgba8_out (vsize*hsize*4 bytes)
 * and / or rgb_f64_out (vsize*hsize*3 doubles, the Canvas colours) — every device renders its bands and its copy engine
 * writes them to their frame positions in these buffers over its own PCIe link, the copies overlapping the rendering
 * (page-locked buffers — rtc_pinned_alloc — let the copies run at PCIe speed; pageable ones work, staged by the driver).
 * The sharded form of `rtc_scene_create_ex; rtc_render(scene, camera, NULL, rgba8, rgb_f64, stats)`. */
int rtc_multi_render_host(rtc_multi* m, const rtc_camera_desc* camera, uint8_t* rgba8_out, double* rgb_f64_out,
                          rtc_stats* stats);
const uint8_t* rtc_multi_host_frame(const rtc_multi* m);
void* rtc_multi_device_frame(const rtc_multi* m);
void rtc_multi_destroy(rtc_multi* m);
int rtc_render_multi(const rtc_scene_desc* desc, const rtc_camera_desc* camera, int ngpus, uint32_t build_flags,
                     uint8_t* rgba8_out, rtc_stats* stats);

/* Number of frame rows a (camera, rows) selection renders = the row count of the compact output buffers. */
uint32_t rtc_rows_count(const rtc_camera_desc* camera, const rtc_rows* rows);

/* World::color_at (src/world.rs:80-82) for `n` explicit rays (origin xyz, direction xyz; host pointers) -> rgb f64. */
int rtc_color_at(const rtc_scene* scene, const double* rays, uint64_t n, double* rgb_out);

/* Probes of the per-ray program, so the reference's own unit tests can be asked of the CUDA path (never on a frame's
 * path).  A leaf is named by its index in the pre-order walk of the World's leaves (groups do not count).
 *   rtc_intersect             World::intersect (src/world.rs:43-54): every intersection of each ray, in the order the
 *                             reference's stable sorts leave them (t ascending; ties: pre-order of the leaves, then push
 *                             order).  Ray i fills t_out / leaf_out [i*cap, i*cap + min(counts[i], cap)); counts[i] is the
 *                             number found, which may exceed cap.
 *   rtc_prepare_computations  Intersection::hit (src/intersection.rs:79-83) then prepare_computations
 *                             (src/intersection.rs:17-77) and Computations::schlick (:107-128) for the hit of each ray.
 *   rtc_normal_at             Shape::normal_at (src/shape.rs:466-519) of one leaf at n world points.  For an
 *                             RTC_SMOOTH_TRIANGLE leaf a "point" is read as (u, v, unused): the book's normal_at(tri, point,
 *                             intersection_with_uv(t, tri, u, v)), which ignores the point. */
typedef struct rtc_computations {
    int32_t hit;    /* 0: no intersection with t >= 0 (everything else is zero) */
    int32_t leaf;   /* Computations.object */
    int32_t inside;
    int32_t _pad;
    double t;
    double point[3], eyev[3], normalv[3], reflectv[3], over_point[3], under_point[3];
    double n1, n2;
    double reflectance;
} rtc_computations;
int rtc_intersect(const rtc_scene* scene, const double* rays, uint64_t n, uint32_t cap, double* t_out, int32_t* leaf_out,
                  uint32_t* counts);
int rtc_prepare_computations(const rtc_scene* scene, const double* rays, uint64_t n, rtc_computations* out);
int rtc_normal_at(const rtc_scene* scene, int32_t leaf, const double* points, uint64_t n, double* normals_out);

/* Work tallies of one frame for the FP64 roofline: renders the rows with a counting build of the same per-ray program
 * (nothing is stored or timed) and fills counts[rtc_tally_count()] — how many ray transforms, gate tests, leaf tests by
 * kind, triangle tests by outcome, BVH box tests, shaded hits, pattern evaluations, pow calls, refraction set-ups,
 * Schlick evaluations and n1/n2 walks the frame needed (order: TallyIndex in csrc/rt_core.cuh). */
int rtc_tally_count(void);
int rtc_render_tally(const rtc_scene* scene, const rtc_camera_desc* camera, const rtc_rows* rows, uint64_t* counts);

/* Self-measured FP64 issue peaks of the device (independent DADD+DMUL chains, and DFMA chains), in Gflop/s with an FMA
 * counted as 2.  Used as the FP64 roofline denominator (the path must run without FMA contraction). */
int rtc_measure_fp64_peak(int device, double* nofma_gflops, double* fma_gflops);

/* Self-test of the kernels' shared-divisor division (several IEEE quotients over one divisor with the reciprocal refinement
 * done once; csrc/rt_core.cuh SharedDivisor): `pairs` generated operand pairs — raw bit patterns, ordinary magnitudes,
 * special values — each divided both ways on the device and compared bit for bit.  *mismatches must come back 0. */
int rtc_selftest_shared_divisor(int device, uint64_t pairs, uint64_t seed, uint64_t* mismatches);

const char* rtc_last_error(void);
int rtc_device_count(void);
/* cudaDeviceEnablePeerAccess(peer) from `device` (no-op if already enabled): lets kernels launched on `device` store into
 * buffers that live on `peer` (RTC_ROWS_FRAME renders into another GPU's frame). */
int rtc_enable_peer_access(int device, int peer);
/* A frame buffer on `device` that other PROCESSES (one per GPU) can map over NVLink: create() allocates it and returns a
 * 64-byte CUDA IPC handle; open() (called with the caller's own device) maps it into that device's address space with
 * peer access enabled and returns a pointer its kernels can store through (RTC_ROWS_FRAME); close() unmaps (owner = 0) or
 * frees (owner = 1). */
int rtc_frame_share_create(int device, uint64_t bytes, void** d_ptr, uint8_t handle64[64]);
int rtc_frame_share_open(int device, const uint8_t handle64[64], void** d_ptr);
int rtc_frame_share_close(int device, void* d_ptr, int owner);

/* The Canvas of a sharded render in HOST memory shared by the PROCESSES (one per GPU): a POSIX shared-memory segment
 * (`name` starts with '/', see shm_open(3)) that every rank maps and page-locks, so each rank's copy engine writes its own
 * bands of the frame over its own PCIe link, all links in parallel (rtc_render with RTC_ROWS_FRAME and pointers into the
 * segment) — instead of funnelling the frame through one GPU and one link.  create() makes the segment, zero-filled, and
 * reserves its pages (fails cleanly when /dev/shm is too small); open() maps an existing one; close() unmaps and, given a
 * name, unlinks it (the creator, after every rank has closed).  `device`: any CUDA device of the calling process; a negative
 * device maps without page-locking (no CUDA call: copies into it still work, staged by the driver).
 * Completion across processes is a counter in the segment itself: 8-byte aligned uint64 slots, store = release, load / wait =
 * acquire; wait spins (then yields) until the slot is >= at_least or fails with RTC_ERR_TIMEOUT.  rtc_render returns after
 * its copies have landed, so "rank r stores frame number f into its slot after rtc_render; the consumer waits for every
 * slot to reach f" is a complete protocol (ray-tracer-challenge-rust_b200/multi.py SharedCanvasRenderer). */
int rtc_host_share_create(int device, const char* name, uint64_t bytes, void** out);
int rtc_host_share_open(int device, const char* name, uint64_t bytes, void** out);
int rtc_host_share_close(void* p, uint64_t bytes, const char* unlink_name);
void rtc_host_counter_store(void* counter, uint64_t value);
uint64_t rtc_host_counter_load(const void* counter);
int rtc_host_counter_wait(const void* counter, uint64_t at_least, double timeout_s);

/* ------------------------------------------------------------------------------------------------------------------
 * 2. HOST MIRROR of the reference API (C++ inside; the Rust host in the reference)
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct rtc_shape rtc_shape;   /* Shape   src/shape.rs:42-49 */
typedef struct rtc_world rtc_world;   /* World   src/world.rs:13-16 */
typedef struct rtc_camera rtc_camera; /* Camera  src/camera.rs:5-12 */
typedef struct rtc_canvas rtc_canvas; /* Canvas  src/canvas.rs:5-9 */
typedef struct rtc_marshalled rtc_marshalled; /* a World flattened into the arrays of an rtc_scene_desc */

/* transformations.rs:4-93, matrix.rs:29-39,138-157,187-227 — row-major 4x4 in/out */
void rtc_translation(double x, double y, double z, double* out16);
void rtc_scaling(double x, double y, double z, double* out16);
void rtc_rotation_x(double rad, double* out16);
void rtc_rotation_y(double rad, double* out16);
void rtc_rotation_z(double rad, double* out16);
void rtc_shearing(double xy, double xz, double yx, double yz, double zx, double zy, double* out16);
int rtc_view_transform(const double* from3, const double* to3, const double* up3, double* out16);
void rtc_matrix_mul(const double* a16, const double* b16, double* out16);
void rtc_matrix_transpose(const double* a16, double* out16);
int rtc_matrix_inverse(const double* a16, double* out16); /* RTC_ERR_PANIC when |det| < 1e-5 (matrix.rs:140) */
void rtc_matrix_mul_tuple(const double* a16, const double* t4, double* out4);

/* material.rs:17-29 defaults; pattern.rs:63-66 */
void rtc_material_default(rtc_material* m);
int rtc_material_set_pattern_transform(rtc_material* m, const double* t16);

/* shape.rs:52-245.  kind: RTC_SPHERE..RTC_GROUP (triangles via rtc_shape_triangle). */
rtc_shape* rtc_shape_new(int kind, double minimum, double maximum, int capped);
rtc_shape* rtc_shape_triangle(const double* p1, const double* p2, const double* p3);
/* The book's smooth_triangle(p1, p2, p3, n1, n2, n3) (RTC_SMOOTH_TRIANGLE; not in the reference). */
rtc_shape* rtc_shape_smooth_triangle(const double* p1, const double* p2, const double* p3, const double* n1, const double* n2,
                                     const double* n3);
void rtc_shape_free(rtc_shape* s);
int rtc_shape_set_transform(rtc_shape* s, const double* m16);      /* shape.rs:196-218: push-down, once per node */
int rtc_shape_set_material(rtc_shape* s, const rtc_material* m);   /* shape.rs:220-229: push-down */
int rtc_shape_push_shape(rtc_shape* group, rtc_shape* child);      /* shape.rs:528-535; consumes `child` */
uint64_t rtc_shape_leaf_count(const rtc_shape* s);

/* obj_file.rs:23-128: returns Parser::obj_to_group().  Named groups keep first-insertion order (the reference iterates
 * a HashMap, i.e. a random order per process). */
rtc_shape* rtc_obj_parse_file(const char* path, uint64_t* ignored_lines);
rtc_shape* rtc_obj_parse_str(const char* text, uint64_t len, uint64_t* ignored_lines);
/* group{ default_group{ triangles } } for vertex/face arrays (faces 1-based, 3 per triangle) */
rtc_shape* rtc_mesh_from_arrays(const double* verts, uint64_t nverts, const int32_t* faces, uint64_t nfaces);
/* The same for an OBJ with `vn` records and `f v//n` faces: corner k of face f uses vertex faces[3f + k] and normal
 * face_normals[3f + k] (both 1-based); every triangle is an RTC_SMOOTH_TRIANGLE. */
rtc_shape* rtc_smooth_mesh_from_arrays(const double* verts, uint64_t nverts, const double* normals, uint64_t nnormals,
                                       const int32_t* faces, const int32_t* face_normals, uint64_t nfaces);

/* world.rs:18-24, 26-41 */
rtc_world* rtc_world_new(const double* light_position3, const double* light_intensity3);
rtc_world* rtc_world_default(void);
void rtc_world_free(rtc_world* w);
int rtc_world_push(rtc_world* w, rtc_shape* s); /* World.objects.push; consumes `s` */
/* World::color_at for n rays on the GPU (uploads the world on first use; cached until the world changes). */
int rtc_world_color_at(rtc_world* w, const double* rays, uint64_t n, double* rgb_out);
/* The layer-1 scene handle this world marshals into (created on first use on `device`, one per device); owned by the
 * world and valid until the world changes (rtc_world_push, rtc_world_set_build, rtc_world_set_recursion_limit,
 * rtc_world_drop_scenes, rtc_world_free).  Threading: every rtc_world_* call and rtc_camera_render hold the world's lock
 * for their whole duration, so they may be called from several threads; a scene handle obtained here must not be used
 * concurrently with a call that changes the world. */
int rtc_world_scene(rtc_world* w, int device, rtc_scene** out);
/* Drops the cached device scenes: the next render marshals, flattens and uploads the world again (what the patched
 * Camera::render of the Rust host does on every call). */
void rtc_world_drop_scenes(rtc_world* w);
/* World's RECURSION_LIMIT (src/world.rs:11; see rtc_scene_desc.recursion_limit).  0 = the reference's 5. */
int rtc_world_set_recursion_limit(rtc_world* w, uint32_t limit);
/* Which mesh build rtc_world_scene / rtc_camera_render use for this world (RTC_BUILD_*; drops a scene built otherwise). */
int rtc_world_set_build(rtc_world* w, uint32_t flags);

/* The layer-1 description of this world: what rtc_world_scene passes to rtc_scene_create, and what the Rust-side
 * Camera::render patch builds from its &World.  The rtc_scene_desc borrows from the rtc_marshalled. */
int rtc_world_marshal(rtc_world* w, rtc_marshalled** out);
const rtc_scene_desc* rtc_marshalled_desc(const rtc_marshalled* m);
void rtc_marshalled_free(rtc_marshalled* m);
/* Host-only views of what rtc_world_scene would send / build (no device needed; used by reports and CPU tests):
 *   describe     n = {shapes, roots, transforms, materials, triangles} of the marshalled rtc_scene_desc
 *   flatten_info n = {leaves, gates, meshes, mesh triangles, bvh nodes, bvh max depth, program nodes, distinct
 *                transforms}; gates_out (nullable) receives the gate boxes, 6 doubles each (lo xyz, hi xyz). */
int rtc_world_describe(rtc_world* w, uint64_t n[5]);
int rtc_world_flatten_info(rtc_world* w, uint64_t n[8], double* gates_out, uint64_t gates_cap);
/* features = {what the flattened world contains, what the render kernel instantiation chosen for it is compiled for}, both
 * as bit masks: 1 spheres, 2 planes, 4 cubes, 8 cylinders, 16 cones, 32 triangle meshes, 64 groups, 128 a transparent
 * material, 256 clusters of bounded sibling leaves, 512 a RECURSION_LIMIT other than 5, 1024 a cluster large enough to be
 * a tree.  The second covers the first; equal masks mean the world has a kernel of its own (reports and tests). */
int rtc_world_kernel_features(rtc_world* w, uint32_t features[2]);

/* camera.rs:16-46 */
rtc_camera* rtc_camera_new(uint64_t hsize, uint64_t vsize, double field_of_view);
void rtc_camera_free(rtc_camera* c);
int rtc_camera_set_transform(rtc_camera* c, const double* m16);
void rtc_camera_desc_get(const rtc_camera* c, rtc_camera_desc* out);
/* camera.rs:67-79: Camera::render(&World) -> Canvas, on GPU `device`, in pinned host memory (pooled: a freed canvas's
 * buffers serve the next canvas of that size).  `want_f64` != 0: the canvas holds the f64 colours (canvas.rs:8), as the
 * reference's does — 24 bytes per pixel cross PCIe — and its RGBA8 pixels are quantised from them on the host the first
 * time they are asked for (rtc_canvas_pixels_rgba8, rtc_canvas_to_ppm: canvas.rs:61-63 happens at PPM time there too).
 * `want_f64` = 0: only the RGBA8 frame the kernel quantised (4 bytes per pixel; get_pixel then fails).
 * device = RTC_DEVICE_ALL: the frame is sharded over EVERY CUDA device of this process — what the Rust host, one process,
 * gets on a multi-GPU box without changing a line: the World is marshalled and flattened once and uploaded to each device
 * (in parallel), the frame's 8-row bands are dealt cyclically, every device renders its bands and its own copy engine writes
 * them to their frame positions in the one pinned canvas over its own PCIe link.  Same pixels as any single device; stats
 * are summed over the devices, device_ms is the slowest device's kernel.  With one device it is device 0. */
#define RTC_DEVICE_ALL (-1)
int rtc_camera_render(const rtc_camera* c, rtc_world* w, int device, int want_f64, rtc_canvas** out, rtc_stats* stats);

/* canvas.rs:12-58 */
rtc_canvas* rtc_canvas_new(uint64_t width, uint64_t height);
void rtc_canvas_free(rtc_canvas* c);
uint64_t rtc_canvas_width(const rtc_canvas* c);
uint64_t rtc_canvas_height(const rtc_canvas* c);
int rtc_canvas_get_pixel(const rtc_canvas* c, uint64_t x, uint64_t y, double* rgb3);
int rtc_canvas_set_pixel(rtc_canvas* c, uint64_t x, uint64_t y, const double* rgb3);
const double* rtc_canvas_pixels_f64(const rtc_canvas* c); /* NULL if rendered with want_f64 = 0 */
const uint8_t* rtc_canvas_pixels_rgba8(const rtc_canvas* c); /* quantised on first use for a want_f64 canvas */
/* canvas.rs:28-58: P3 text, 70-column wrap.  Returns a malloc'd buffer (free with rtc_free). */
char* rtc_canvas_to_ppm(const rtc_canvas* c, uint64_t* len);
/* Canvas::to_ppm ON THE DEVICE: encodes an RGBA8 frame that is still in HBM (e.g. the output of rtc_render_device, or
 * the frame gathered on rank 0) straight into the reference's P3 text — rows in parallel, 70-column wrap per row, prefix
 * sum of row sizes — and copies only the text to out_host (capacity >= rtc_ppm_max_bytes(w, h) always suffices; pinned
 * memory from rtc_pinned_alloc makes the copy run at PCIe speed).  Byte-identical to rtc_canvas_to_ppm. */
uint64_t rtc_ppm_max_bytes(uint64_t width, uint64_t height);
int rtc_ppm_encode_device(int device, const void* d_rgba8, uint64_t width, uint64_t height, void* cuda_stream,
                          char* out_host, uint64_t capacity, uint64_t* len);
void* rtc_pinned_alloc(uint64_t bytes); /* NULL without a CUDA device */
void rtc_pinned_free(void* p);
/* the same encoder over a caller-owned RGBA8 frame (e.g. a frame gathered from several GPUs) */
char* rtc_ppm_from_rgba8(const uint8_t* rgba8, uint64_t width, uint64_t height, uint64_t* len);
void rtc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* RTC_H */
