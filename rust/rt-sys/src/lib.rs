//! Raw bindings to `librt_b200.so` -- one to one with `include/rt_api.h` (RT_API_VERSION 2).
//!
//! UNCOMPILED in this repository's image (no cargo/rustc): `tests/test_rust_binding.py` checks every `#[repr(C)]` struct,
//! every `extern "C"` prototype and every constant below against the header, and the `size_of` / `offset_of` assertions
//! at the bottom repeat the `static_assert`s of `csrc/rt_api.cu`.
//!
//! The seam this replaces: `pub fn render_scene(scene: &Scene) -> Vec<u8>` (`src/rendering.rs:21`, called at `src/main.rs:55`).
#![allow(non_camel_case_types, non_upper_case_globals)]
use std::os::raw::{c_char, c_int, c_void};

pub const RT_API_VERSION: c_int = 2;

pub const RT_OK: c_int = 0;
pub const RT_ERR_INVALID: c_int = 1;
pub const RT_ERR_IO: c_int = 2;
pub const RT_ERR_FORMAT: c_int = 3;
pub const RT_ERR_CUDA: c_int = 4;
pub const RT_ERR_LIMIT: c_int = 5;

pub const RT_SHAPE_TRIANGLE: c_int = 0;
pub const RT_SHAPE_BOX: c_int = 1;
pub const RT_SHAPE_ELLIPSOID: c_int = 2;
pub const RT_SHAPE_PLANE: c_int = 3;
pub const RT_MATERIAL_PBR: c_int = 0;
pub const RT_MATERIAL_DIELECTRIC: c_int = 1;

pub const RT_FN_BRDF: c_int = 1;
pub const RT_FN_PDF_COSINE: c_int = 2;
pub const RT_FN_PDF_VNDF: c_int = 3;
pub const RT_FN_PDF_LIGHT: c_int = 4;
pub const RT_FN_PDF_MIX: c_int = 5;
pub const RT_FN_SAMPLE_COSINE: c_int = 6;
pub const RT_FN_SAMPLE_VNDF: c_int = 7;
pub const RT_FN_SAMPLE_LIGHT: c_int = 8;
pub const RT_FN_PHILOX: c_int = 9;
pub const RT_FN_SAMPLE_LIGHT_GEN: c_int = 10;
pub const RT_FN_DIELECTRIC: c_int = 11;

/// Opaque scene handle (host copy + device copy + streams), owned by the library until `rt_scene_destroy`.
#[repr(C)]
pub struct RtScene {
    _private: [u8; 0],
}

/// `Scene` (src/scene.rs:22-39) flattened, f64, load order.
#[repr(C)]
pub struct RtSceneDesc {
    pub width: i32,
    pub height: i32,
    pub samples: i32,
    pub ray_depth: i32,
    pub bg_color: [f64; 3],
    pub camera_position: [f64; 3],
    pub camera_forward: [f64; 3],
    pub camera_right: [f64; 3],
    pub camera_up: [f64; 3],
    pub camera_fov_x: f64,
    pub camera_fov_y: f64,
    pub n_tris: i32,
    pub reserved0: i32,
    pub tri_v: *const f64,
    pub tri_n: *const f64,
    pub tri_material: *const f64,
    pub tri_emission: *const f64,
}

/// `Object3D { shape, position, rotation }` (src/geometry.rs:41-46) + `Primitive.ior` (src/scene.rs:18) per primitive.
#[repr(C)]
pub struct RtSceneDesc2 {
    pub base: RtSceneDesc,
    pub shape_kind: *const i32,
    pub position: *const f64,
    pub rotation: *const f64,
    pub ior: *const f64,
    pub material_kind: *const i32,
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct RtSceneInfo {
    pub n_tris: i32,
    pub n_lights: i32,
    pub n_materials: i32,
    pub n_nodes: i32,
    pub n_leaves: i32,
    pub bvh_depth: i32,
    pub max_leaf_size: i32,
    pub bvh_validate_failures: i32,
    pub scene_in_shared_memory: i32,
    pub device: i32,
    pub device_bytes: i64,
    pub bvh_builder: i32,
    pub reserved1: i32,
    pub bvh_build_ms: f64,
    pub n_infinite: i32,
    pub general_primitives: i32,
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct RtRenderParams {
    pub seed: u64,
    pub sample_begin: i32,
    pub sample_end: i32,
    pub max_attempts: i32,
    pub collect_stats: i32,
    pub kernel_variant: i32,
    pub tile_shard_index: i32,
    pub tile_shard_count: i32,
    pub reserved: i32,
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct RtStats {
    pub samples: u64,
    pub segments: u64,
    pub vertices: u64,
    pub attempts: u64,
    pub node_tests: u64,
    pub tri_tests: u64,
    pub light_tri_tests: u64,
    pub attempt_cap_hits: u64,
    pub nonfinite_samples: u64,
    pub kernel_launches: u64,
    pub kernel_ms: f64,
    pub total_ms: f64,
    pub kernel: i32,
    pub block_threads: i32,
    pub blocks_per_sm: i32,
    pub grid_blocks: i32,
    pub regs_per_thread: i32,
    pub smem_bytes_per_block: i32,
    pub scene_in_shared_memory: i32,
    pub n_chunks: i32,
    pub render_ms: f64,
    pub reduce_ms: f64,
    pub resolve_ms: f64,
}

extern "C" {
    pub fn rt_last_error() -> *const c_char;
    pub fn rt_api_version() -> c_int;
    pub fn rt_device_count(count: *mut i32) -> c_int;

    pub fn rt_scene_load_gltf(path: *const c_char, width: i32, height: i32, samples: i32, device: i32, out: *mut *mut RtScene) -> c_int;
    pub fn rt_scene_load_text(path: *const c_char, width: i32, height: i32, samples: i32, device: i32, out: *mut *mut RtScene) -> c_int;
    pub fn rt_scene_load(path: *const c_char, width: i32, height: i32, samples: i32, device: i32, out: *mut *mut RtScene) -> c_int;
    pub fn rt_scene_create(desc: *const RtSceneDesc, device: i32, out: *mut *mut RtScene) -> c_int;
    pub fn rt_scene_create2(desc: *const RtSceneDesc2, device: i32, out: *mut *mut RtScene) -> c_int;
    pub fn rt_scene_destroy(scene: *mut RtScene);
    pub fn rt_release_device_cache() -> c_int;
    pub fn rt_scene_get_desc(scene: *const RtScene, out: *mut RtSceneDesc) -> c_int;
    pub fn rt_scene_get_desc2(scene: *const RtScene, out: *mut RtSceneDesc2) -> c_int;
    pub fn rt_scene_info(scene: *const RtScene, out: *mut RtSceneInfo) -> c_int;
    pub fn rt_scene_set_frame(scene: *mut RtScene, width: i32, height: i32, samples: i32) -> c_int;
    pub fn rt_scene_get_bvh(scene: *const RtScene, nodes: *mut f32, tri_order: *mut i32) -> c_int;
    pub fn rt_scene_get_quantised_bvh(scene: *const RtScene, words: *mut u32, grid6: *mut f32, present: *mut i32) -> c_int;

    pub fn rt_render(scene: *mut RtScene, params: *const RtRenderParams, rgb_out: *mut u8, stats: *mut RtStats) -> c_int;
    pub fn rt_render_linear(scene: *mut RtScene, params: *const RtRenderParams, rgb_linear_out: *mut f32, stats: *mut RtStats) -> c_int;
    pub fn rt_render_accumulate_device(scene: *mut RtScene, params: *const RtRenderParams, accum_dev: *mut f32, stream: *mut c_void, stats: *mut RtStats) -> c_int;
    pub fn rt_resolve_device(accum_dev: *const f32, width: i32, height: i32, rgb_dev: *mut u8, stream: *mut c_void) -> c_int;
    pub fn rt_render_multi(scenes: *const *mut RtScene, n_scenes: i32, params: *const RtRenderParams, rgb_out: *mut u8, stats: *mut RtStats) -> c_int;
    pub fn rt_multi_init(scenes: *const *mut RtScene, n_scenes: i32) -> c_int;

    pub fn rt_trace_primary(scene: *mut RtScene, rays: *const f64, n: i64, precision: i32, tri_id: *mut i32, t: *mut f64) -> c_int;
    pub fn rt_trace_hits(scene: *mut RtScene, rays: *const f64, n: i64, out: *mut f64) -> c_int;
    pub fn rt_primary_rays(scene: *mut RtScene, xy: *const i32, xi: *const f64, n: i64, rays_out: *mut f64) -> c_int;
    pub fn rt_eval(scene_or_null: *mut RtScene, r#fn: i32, r#in: *const f32, n: i64, out: *mut f32) -> c_int;

    pub fn rt_write_ppm(path: *const c_char, width: i32, height: i32, rgb: *const u8, append: i32) -> c_int;
    pub fn rt_write_png(path: *const c_char, width: i32, height: i32, rgb: *const u8) -> c_int;
    pub fn rt_debug_bounds_violations(device: i32, counts: *mut u64, reset: i32) -> c_int;
    pub fn rt_measure_fp32_peak(device: i32, tflops: *mut f64, sm_mhz: *mut f64) -> c_int;
}

// Layout guards: the same numbers as the static_asserts in raytracing-course-2024_b200/csrc/rt_api.cu.
const _: () = assert!(std::mem::size_of::<RtSceneDesc>() == 192);
const _: () = assert!(std::mem::size_of::<RtSceneDesc2>() == 232);
const _: () = assert!(std::mem::size_of::<RtSceneInfo>() == 72);
const _: () = assert!(std::mem::size_of::<RtRenderParams>() == 40);
const _: () = assert!(std::mem::size_of::<RtStats>() == 152);
const _: () = assert!(std::mem::offset_of!(RtSceneDesc, bg_color) == 16);
const _: () = assert!(std::mem::offset_of!(RtSceneDesc, camera_fov_x) == 136);
const _: () = assert!(std::mem::offset_of!(RtSceneDesc, n_tris) == 152);
const _: () = assert!(std::mem::offset_of!(RtSceneDesc, tri_v) == 160);
const _: () = assert!(std::mem::offset_of!(RtSceneDesc2, shape_kind) == 192);
const _: () = assert!(std::mem::offset_of!(RtSceneInfo, device_bytes) == 40);
const _: () = assert!(std::mem::offset_of!(RtSceneInfo, bvh_build_ms) == 56);
const _: () = assert!(std::mem::offset_of!(RtRenderParams, sample_begin) == 8);
const _: () = assert!(std::mem::offset_of!(RtStats, kernel_ms) == 80);
const _: () = assert!(std::mem::offset_of!(RtStats, kernel) == 96);
const _: () = assert!(std::mem::offset_of!(RtStats, render_ms) == 128);
