//! rtc-sys — raw FFI bindings to `librtc_b200.so`, the B200 (sm_100a) CUDA render path that replaces
//! `Camera::render` (src/camera.rs:67-79) and `World::color_at` (src/world.rs:80-82) of
//! antoinehebert/ray-tracer-challenge-rust.
//!
//! These are the CORE BOUNDARY declarations of `include/rtc.h` (layer 1), field for field and symbol for symbol; the
//! header's HOST MIRROR half (`rtc_shape_*`, `rtc_world_*`, ...) restates the reference's own Rust host API for hosts that
//! have no Rust and is deliberately not bound here — the Rust host keeps its own `Shape / World / Camera / Canvas`.
//!
//! Every fallible call returns `RTC_OK` (0) or a negative code and leaves a thread-local message for
//! [`rtc_last_error`]; nothing aborts.  The reference's convention is to panic, so the patched `Camera::render`
//! (integration/reference.patch) turns a non-zero status into `panic!`.
//!
//! `tests/test_boundary.py` parses this file and checks every `#[repr(C)]` struct (field order, types, size) and every
//! `extern "C"` symbol against `include/rtc.h` and the built library.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const RTC_OK: c_int = 0;
pub const RTC_ERR_INVALID: c_int = -1;
pub const RTC_ERR_PANIC: c_int = -2;
pub const RTC_ERR_CUDA: c_int = -3;
pub const RTC_ERR_UNSUPPORTED: c_int = -4;
pub const RTC_ERR_TIMEOUT: c_int = -6;
/// rtc_camera_render: shard the frame over every CUDA device of the process
pub const RTC_DEVICE_ALL: c_int = -1;

/// ShapeKind (src/shape.rs:14-39)
pub const RTC_SPHERE: i32 = 0;
pub const RTC_PLANE: i32 = 1;
pub const RTC_CUBE: i32 = 2;
pub const RTC_CYLINDER: i32 = 3;
pub const RTC_CONE: i32 = 4;
pub const RTC_GROUP: i32 = 5;
pub const RTC_TRIANGLE: i32 = 6;
pub const RTC_SMOOTH_TRIANGLE: i32 = 7;

/// PatternKind (src/pattern.rs:4-12)
pub const RTC_PATTERN_NONE: i32 = -1;
pub const RTC_PATTERN_STRIPE: i32 = 0;
pub const RTC_PATTERN_GRADIENT: i32 = 1;
pub const RTC_PATTERN_RING: i32 = 2;
pub const RTC_PATTERN_CHECKERS: i32 = 3;
pub const RTC_PATTERN_TEST: i32 = 4;

pub const RTC_ROWS_COMPACT: u32 = 0;
pub const RTC_ROWS_FRAME: u32 = 1;
pub const RTC_BUILD_HOST_SAH: u32 = 0;
pub const RTC_BUILD_DEVICE_LBVH: u32 = 1;

/// Material (src/material.rs:4-14) with its Option<Pattern> (src/pattern.rs:14-19) flattened in; matrices row-major.
#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct rtc_material {
    pub color: [f64; 3],
    pub ambient: f64,
    pub diffuse: f64,
    pub specular: f64,
    pub shininess: f64,
    pub reflective: f64,
    pub transparency: f64,
    pub refractive_index: f64,
    pub pattern_kind: i32,
    pub _pad: i32,
    pub pattern_a: [f64; 3],
    pub pattern_b: [f64; 3],
    pub pattern_transform: [f64; 16],
    pub pattern_inverse: [f64; 16],
}

/// Shape.transform and Shape.transform_inverse (src/shape.rs:44-45)
#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct rtc_transform_desc {
    pub transform: [f64; 16],
    pub inverse: [f64; 16],
}

/// ShapeKind::Triangle payload (src/shape.rs:31-38); n1..n3 are the vertex normals of a smooth triangle
#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct rtc_triangle_desc {
    pub p1: [f64; 3],
    pub p2: [f64; 3],
    pub p3: [f64; 3],
    pub e1: [f64; 3],
    pub e2: [f64; 3],
    pub normal: [f64; 3],
}

/// Vertex normals of a smooth triangle (the book's SmoothTriangle; absent from the reference)
#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct rtc_vertex_normals {
    pub n1: [f64; 3],
    pub n2: [f64; 3],
    pub n3: [f64; 3],
}

/// One Shape of the pre-order walk of World.objects (src/shape.rs:28-30)
#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct rtc_shape_desc {
    pub kind: i32,
    pub material: i32,
    pub transform: i32,
    pub capped: i32,
    pub minimum: f64,
    pub maximum: f64,
    pub child_count: i32,
    pub triangle: i32,
}

/// World (src/world.rs:13-16) + Light (src/light.rs:5-8)
#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct rtc_scene_desc {
    pub shapes: *const rtc_shape_desc,
    pub shape_count: u32,
    pub root_count: u32,
    pub transforms: *const rtc_transform_desc,
    pub transform_count: u32,
    pub materials: *const rtc_material,
    pub material_count: u32,
    pub triangles: *const rtc_triangle_desc,
    pub triangle_count: u32,
    pub light_position: [f64; 3],
    pub light_intensity: [f64; 3],
    pub vertex_normals: *const rtc_vertex_normals,
    pub recursion_limit: u32,
    pub _reserved: u32,
}

/// Camera (src/camera.rs:5-12)
#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct rtc_camera_desc {
    pub hsize: u32,
    pub vsize: u32,
    pub inverse: [f64; 16],
    pub half_width: f64,
    pub half_height: f64,
    pub pixel_size: f64,
}

#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct rtc_rows {
    pub band_rows: u32,
    pub band_first: u32,
    pub band_stride: u32,
    pub layout: u32,
}

#[repr(C)]
#[derive(Debug, Copy, Clone, Default)]
pub struct rtc_stats {
    pub primary_rays: u64,
    pub shadow_rays: u64,
    pub reflect_rays: u64,
    pub refract_rays: u64,
    pub kernel_launches: u64,
    pub device_ms: f64,
}

/// Computations (src/intersection.rs:88-100) + Computations::schlick of the hit of one ray
#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct rtc_computations {
    pub hit: i32,
    pub leaf: i32,
    pub inside: i32,
    pub _pad: i32,
    pub t: f64,
    pub point: [f64; 3],
    pub eyev: [f64; 3],
    pub normalv: [f64; 3],
    pub reflectv: [f64; 3],
    pub over_point: [f64; 3],
    pub under_point: [f64; 3],
    pub n1: f64,
    pub n2: f64,
    pub reflectance: f64,
}

/// Opaque device-resident scene
#[repr(C)]
pub struct rtc_scene {
    _private: [u8; 0],
}

/// Opaque multi-device renderer (one scene per device of this process)
#[repr(C)]
pub struct rtc_multi {
    _private: [u8; 0],
}

pub const RTC_MULTI_HOST_FRAME: u32 = 0;
pub const RTC_MULTI_DEVICE_FRAME: u32 = 1;

extern "C" {
    pub fn rtc_last_error() -> *const c_char;
    pub fn rtc_device_count() -> c_int;

    pub fn rtc_scene_create(desc: *const rtc_scene_desc, device: c_int, out: *mut *mut rtc_scene) -> c_int;
    pub fn rtc_scene_create_ex(desc: *const rtc_scene_desc, device: c_int, flags: u32, out: *mut *mut rtc_scene) -> c_int;
    pub fn rtc_scene_destroy(scene: *mut rtc_scene);
    pub fn rtc_scene_info(scene: *const rtc_scene, n: *mut u64) -> c_int;
    pub fn rtc_scene_upload_bytes(scene: *const rtc_scene) -> u64;

    /// Camera::render (src/camera.rs:67-79), host buffers; either output may be null
    pub fn rtc_render(
        scene: *const rtc_scene,
        camera: *const rtc_camera_desc,
        rows: *const rtc_rows,
        rgba8_out: *mut u8,
        rgb_f64_out: *mut f64,
        stats: *mut rtc_stats,
    ) -> c_int;
    /// the same with device buffers, asynchronous on a cudaStream_t
    pub fn rtc_render_device(
        scene: *const rtc_scene,
        camera: *const rtc_camera_desc,
        rows: *const rtc_rows,
        d_rgba8_out: *mut c_void,
        d_rgb_f64_out: *mut c_void,
        cuda_stream: *mut c_void,
        sync_stats: c_int,
        stats: *mut rtc_stats,
    ) -> c_int;
    /// one frame sharded over `ngpus` devices of this process (row bands, peer stores into device 0's frame)
    pub fn rtc_render_device_notify(
        scene: *const rtc_scene,
        camera: *const rtc_camera_desc,
        rows: *const rtc_rows,
        d_rgba8_out: *mut c_void,
        d_rgb_f64_out: *mut c_void,
        cuda_stream: *mut c_void,
        d_counter: *mut c_void,
    ) -> c_int;
    pub fn rtc_stream_wait_counter(device: c_int, cuda_stream: *mut c_void, d_counter: *mut c_void, at_least: u32) -> c_int;
    pub fn rtc_stream_set_counters(
        device: c_int,
        cuda_stream: *mut c_void,
        d_counters: *mut *mut c_void,
        n: u32,
        value: u32,
    ) -> c_int;
    pub fn rtc_render_multi(
        desc: *const rtc_scene_desc,
        camera: *const rtc_camera_desc,
        ngpus: c_int,
        flags: u32,
        rgba8_out: *mut u8,
        stats: *mut rtc_stats,
    ) -> c_int;
    pub fn rtc_multi_create(desc: *const rtc_scene_desc, ngpus: c_int, build_flags: u32, out: *mut *mut rtc_multi) -> c_int;
    pub fn rtc_multi_render(
        m: *mut rtc_multi,
        camera: *const rtc_camera_desc,
        where_: u32,
        rgba8_out: *mut u8,
        stats: *mut rtc_stats,
    ) -> c_int;
    /// rtc_render sharded over the devices of `m`, into caller host buffers (RGBA8 and / or the f64 Canvas colours)
    pub fn rtc_multi_render_host(
        m: *mut rtc_multi,
        camera: *const rtc_camera_desc,
        rgba8_out: *mut u8,
        rgb_f64_out: *mut f64,
        stats: *mut rtc_stats,
    ) -> c_int;
    pub fn rtc_multi_host_frame(m: *const rtc_multi) -> *const u8;
    pub fn rtc_multi_device_frame(m: *const rtc_multi) -> *mut c_void;
    pub fn rtc_multi_destroy(m: *mut rtc_multi);
    pub fn rtc_rows_count(camera: *const rtc_camera_desc, rows: *const rtc_rows) -> u32;

    /// World::color_at (src/world.rs:80-82) for explicit rays
    pub fn rtc_color_at(scene: *const rtc_scene, rays: *const f64, n: u64, rgb_out: *mut f64) -> c_int;
    /// World::intersect (src/world.rs:43-54), sorted
    pub fn rtc_intersect(
        scene: *const rtc_scene,
        rays: *const f64,
        n: u64,
        cap: u32,
        t_out: *mut f64,
        leaf_out: *mut i32,
        counts: *mut u32,
    ) -> c_int;
    /// Intersection::hit + prepare_computations + schlick (src/intersection.rs:17-128)
    pub fn rtc_prepare_computations(scene: *const rtc_scene, rays: *const f64, n: u64, out: *mut rtc_computations) -> c_int;
    /// Shape::normal_at (src/shape.rs:466-519)
    pub fn rtc_normal_at(scene: *const rtc_scene, leaf: i32, points: *const f64, n: u64, normals_out: *mut f64) -> c_int;

    pub fn rtc_tally_count() -> c_int;
    pub fn rtc_render_tally(
        scene: *const rtc_scene,
        camera: *const rtc_camera_desc,
        rows: *const rtc_rows,
        counts: *mut u64,
    ) -> c_int;
    pub fn rtc_measure_fp64_peak(device: c_int, nofma_gflops: *mut f64, fma_gflops: *mut f64) -> c_int;
    pub fn rtc_selftest_shared_divisor(device: c_int, pairs: u64, seed: u64, mismatches: *mut u64) -> c_int;

    pub fn rtc_enable_peer_access(device: c_int, peer: c_int) -> c_int;
    pub fn rtc_frame_share_create(device: c_int, bytes: u64, d_ptr: *mut *mut c_void, handle64: *mut u8) -> c_int;
    pub fn rtc_frame_share_open(device: c_int, handle64: *const u8, d_ptr: *mut *mut c_void) -> c_int;
    pub fn rtc_frame_share_close(device: c_int, d_ptr: *mut c_void, owner: c_int) -> c_int;
    /// the Canvas of a sharded render in host memory shared by one process per GPU (POSIX shared memory, page-locked)
    pub fn rtc_host_share_create(device: c_int, name: *const c_char, bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn rtc_host_share_open(device: c_int, name: *const c_char, bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn rtc_host_share_close(p: *mut c_void, bytes: u64, unlink_name: *const c_char) -> c_int;
    pub fn rtc_host_counter_store(counter: *mut c_void, value: u64);
    pub fn rtc_host_counter_load(counter: *const c_void) -> u64;
    pub fn rtc_host_counter_wait(counter: *const c_void, at_least: u64, timeout_s: f64) -> c_int;

    /// Canvas::to_ppm (src/canvas.rs:28-58) for an RGBA8 frame in device memory
    pub fn rtc_ppm_max_bytes(width: u64, height: u64) -> u64;
    pub fn rtc_ppm_encode_device(
        device: c_int,
        d_rgba8: *const c_void,
        width: u64,
        height: u64,
        cuda_stream: *mut c_void,
        out_host: *mut c_char,
        capacity: u64,
        len: *mut u64,
    ) -> c_int;
    pub fn rtc_pinned_alloc(bytes: u64) -> *mut c_void;
    pub fn rtc_pinned_free(p: *mut c_void);
}

/// The message of the last failed call on this thread.
pub fn last_error() -> String {
    unsafe {
        let p = rtc_last_error();
        if p.is_null() {
            String::new()
        } else {
            std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}
