/* rt_api.h -- C ABI of librt_b200.so, the B200-native replacement for the reference's per-pixel path-tracing
 * loop (metametamoon/raytracing-course-2024).
 *
 * The reference has no FFI layer: its seam is ONE Rust call,
 *     pub fn render_scene(scene: &Scene) -> Vec<u8>            (src/rendering.rs:21, called at src/main.rs:55)
 * fed by  gltf::import + convert_gltf_to_scene                  (src/main.rs:45-47, src/gltf_to_scene.rs:21-79)
 * and followed by dump_rendered_to_ppm                          (src/main.rs:88-95).
 * Every entry point below names the reference interface it replaces.  All signatures are plain C: pointers,
 * sizes and PODs, no C++/torch types.  A Rust `rt-sys` crate binds these 1:1.
 *
 * Conventions: every function returns 0 on success or an RT_ERR_* code; rt_last_error() gives the message of the
 * calling thread's last failure.  Nothing aborts the process (the reference panics instead: unwrap()/assert!).
 * Host buffers are caller-owned.  An RtScene owns its host copy, its device copy and its streams until
 * rt_scene_destroy.  There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * RT_ERR_CUDA.
 */
#ifndef RT_API_H
#define RT_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_API_VERSION 2

enum {
    RT_OK = 0,
    RT_ERR_INVALID = 1,  /* bad argument (samples <= 0, width*height <= 0, null pointer, ...)            */
    RT_ERR_IO = 2,       /* file could not be read / written                                           */
    RT_ERR_FORMAT = 3,   /* glTF the reference would reject (todo!() at gltf_to_scene.rs:131-133,151-153) */
    RT_ERR_CUDA = 4,     /* no device, or a CUDA call failed                                           */
    RT_ERR_LIMIT = 5     /* scene exceeds a build-time limit (e.g. BVH deeper than the traversal stack)  */
};

typedef struct RtScene RtScene; /* opaque */

/* The reference's `Scene` (src/scene.rs:22-39) flattened, in the reference's own numeric type (f64,
 * src/geometry.rs:5) and in LOAD order (the order gltf_to_scene.rs:170-241 pushes primitives).  Only triangles
 * exist at HEAD for glTF input (Shape3D::Triangle, identity rotation, zero position).  Materials are per
 * primitive like scene.rs:13-20.  The BVHs of scene.rs:31,37 are NOT part of the description: rt_scene_create
 * builds its own device BVH (node order is not observable through render_scene). */
typedef struct RtSceneDesc {
    int32_t width, height;            /* scene.rs:23-24                                                   */
    int32_t samples;                  /* scene.rs:35                                                      */
    int32_t ray_depth;                /* scene.rs:33 (6 for glTF scenes, gltf_to_scene.rs:73)             */
    double bg_color[3];               /* scene.rs:25                                                      */
    double camera_position[3];        /* scene.rs:26                                                      */
    double camera_forward[3];         /* scene.rs:27                                                      */
    double camera_right[3];           /* scene.rs:28                                                      */
    double camera_up[3];              /* scene.rs:29                                                      */
    double camera_fov_x, camera_fov_y; /* radians, scene.rs:30-31                                         */
    int32_t n_tris;
    int32_t reserved0;
    const double* tri_v;              /* n_tris x 9: a, b, c                     (geometry.rs:31-38)      */
    const double* tri_n;              /* n_tris x 9: a_norm, b_norm, c_norm                               */
    const double* tri_material;       /* n_tris x 5: base_color rgb, metallic_factor, metallic_roughness (scene.rs:6-11) */
    const double* tri_emission;       /* n_tris x 3                                (scene.rs:19)          */
} RtSceneDesc;

/* General primitives: the reference's Object3D { shape, position, rotation } (src/geometry.rs:41-46) with
 * Shape3D::Box { s } (geometry.rs:27-30) next to Shape3D::Triangle, the dead-but-present Primitive.ior (scene.rs:18), and the
 * shapes / material of the course's text scene format that reference HEAD no longer has (PLANE, ELLIPSOID, DIELECTRIC: this
 * repository's own specification).  Planes are the reference's Scene::infinite_primitives (scene.rs:37,
 * scanned after the BVH, rendering.rs:215-224); every other primitive goes to the BVH (scene.rs:33). */
enum { RT_SHAPE_TRIANGLE = 0, RT_SHAPE_BOX = 1, RT_SHAPE_ELLIPSOID = 2, RT_SHAPE_PLANE = 3 };
enum { RT_MATERIAL_PBR = 0,        /* Material { base_color_factor, metallic_factor, metallic_roughness } (scene.rs:6-11)  */
       RT_MATERIAL_DIELECTRIC = 1  /* own spec: smooth dielectric with Primitive.ior, tinted by base_color on entry         */ };
typedef struct RtSceneDesc2 {
    RtSceneDesc base;                 /* n_tris = number of primitives; tri_v row = a,b,c | s,-,- | radii,-,- | normal,-,- ;
                                         tri_n rows are read for triangles only and are OBJECT-space normals             */
    const int32_t* shape_kind;        /* n: RT_SHAPE_*                                    (geometry.rs:27-39)             */
    const double* position;           /* n x 3: Object3D.position                         (geometry.rs:44)                */
    const double* rotation;           /* n x 4: Object3D.rotation, unit quaternion (i, j, k, w)   (geometry.rs:45)        */
    const double* ior;                /* n: Primitive.ior                                 (scene.rs:18)                   */
    const int32_t* material_kind;     /* n: RT_MATERIAL_*                                                                 */
} RtSceneDesc2;

typedef struct RtSceneInfo {
    int32_t n_tris, n_lights, n_materials;
    int32_t n_nodes;                  /* inner (child-pair) nodes of the device BVH                       */
    int32_t n_leaves, bvh_depth, max_leaf_size;
    int32_t bvh_validate_failures;    /* validate_bvh (bvh.rs:299-322) restated on the flattened BVH      */
    int32_t scene_in_shared_memory;   /* 1 if the render kernels stage the whole scene in shared memory   */
    int32_t device;
    int64_t device_bytes;             /* bytes of the device-resident scene                               */
    int32_t bvh_builder;              /* 0 = host SAH sweep (default), 1 = GPU LBVH builder (env RT_BVH_BUILDER=gpu) */
    int32_t reserved1;
    double  bvh_build_ms;             /* time spent building the finite-primitive BVH (create_bvh_tree, gltf_to_scene.rs:72) */
    int32_t n_infinite;               /* planes = Scene::infinite_primitives (scene.rs:37); n_tris counts the finite primitives */
    int32_t general_primitives;       /* 1: the scene holds boxes / ellipsoids / planes / object transforms / dielectrics    */
} RtSceneInfo;

typedef struct RtRenderParams {
    uint64_t seed;                    /* Philox key; the reference seeds xoshiro per row (rendering.rs:50-51) */
    int32_t sample_begin, sample_end; /* this call renders samples [begin,end) of [0,samples); 0,0 = all  */
    int32_t max_attempts;             /* cap of the rejection loop rendering.rs:102-110 (0 = default 64; clamped to 127) */
    int32_t collect_stats;            /* 1: run the instrumented kernel and fill the work counters         */
    int32_t kernel_variant;           /* 0 = auto; else 10*kernel + placement (benchmarks / A-B tests):
                                         kernel 1 = per-lane megakernel, 2 / 3 = warp-local wavefront with while-while / phased trace bursts;
                                         placement 1 = scene in global memory, 2 = scene in shared memory   */
    int32_t tile_shard_index;         /* with tile_shard_count > 1: render only the 8x4-pixel tiles t with t % count == index  */
    int32_t tile_shard_count;         /* (interleaved tile sharding, the multi-GPU fallback when samples < GPUs); 0 / 1 = whole frame */
    int32_t reserved;
} RtRenderParams;

typedef struct RtStats {
    uint64_t samples;                 /* camera paths started                                              */
    uint64_t segments;                /* nearest-hit queries (rendering.rs:96)                            */
    uint64_t vertices;                /* hits that ran the sampling loop                                   */
    uint64_t attempts;                /* iterations of the rejection loop rendering.rs:102-110             */
    uint64_t node_tests;              /* ray/AABB slab tests                                               */
    uint64_t tri_tests;               /* ray/triangle tests in nearest-hit queries                         */
    uint64_t light_tri_tests;         /* ray/triangle tests of the light pdf (distributions.rs:160-184)    */
    uint64_t attempt_cap_hits;        /* paths cut because max_attempts was reached (reference: spins)     */
    uint64_t nonfinite_samples;       /* paths CUT at a vertex whose throughput became NaN/Inf: the radiance gathered before the cut is
                                         kept and the sample counts (the reference would turn the whole pixel black, NaN -> 0)    */
    uint64_t kernel_launches;         /* CUDA kernels launched by this call                                */
    double   kernel_ms;               /* device time of the render kernels (CUDA events)                   */
    double   total_ms;                /* device time of the whole call incl. copies (CUDA events)          */
    int32_t  kernel;                  /* render kernel that ran: 1 = per-lane megakernel, 2 / 3 = warp-local wavefront (while-while / phased bursts) */
    int32_t  block_threads, blocks_per_sm, grid_blocks, regs_per_thread, smem_bytes_per_block;   /* its launch configuration */
    int32_t  scene_in_shared_memory;  /* 1: the scene blob was staged in shared memory by every block      */
    int32_t  n_chunks;                /* per-chunk float4 layers the render kernel wrote (summed by sum_layers in fixed order)      */
    double   render_ms;               /* device time of the path-tracing kernel + layer sum (slowest device for rt_render_multi) */
    double   reduce_ms;               /* device time of the cross-GPU framebuffer reduce (0 on one GPU)    */
    double   resolve_ms;              /* device time of color_to_pixel                                     */
} RtStats;

/* ---- error reporting ------------------------------------------------------------------------------------ */
const char* rt_last_error(void);
int rt_api_version(void);
int rt_device_count(int32_t* count);

/* ---- scene ingest ---------------------------------------------------------------------------------------- */
/* Replaces gltf::import(path) + convert_gltf_to_scene(&gltf,&buffers,w,h,samples)  (main.rs:45-47,
 * gltf_to_scene.rs:21-79) including create_bvh_tree (gltf_to_scene.rs:72,77): parses the .gltf/.bin on the
 * host, builds the device BVH and uploads everything to `device`.  device = -1 creates a HOST-ONLY scene (loader
 * and BVH inspection through rt_scene_get_desc / rt_scene_info / rt_scene_get_bvh); every compute entry point
 * fails on it with RT_ERR_CUDA. */
int rt_scene_load_gltf(const char* path, int32_t width, int32_t height, int32_t samples, int32_t device, RtScene** out);
/* The course's text scene format (DIMENSIONS, CAMERA_*, RAY_DEPTH, SAMPLES, BG_COLOR, NEW_PRIMITIVE with PLANE / ELLIPSOID / BOX /
 * TRIANGLE, POSITION, ROTATION, COLOR, EMISSION, METALLIC, DIELECTRIC, IOR).  Reference HEAD has NO parser for it
 * (main.rs:48 keeps `// let scene = parse_file_content(file_lines);`): grammar recovered from scenes/practice3_*.txt and
 * scenes/working.txt, semantics = own spec.  width / height / samples > 0 override the file's
 * DIMENSIONS / SAMPLES (the reference CLI passes them, main.rs:39-41); <= 0 keeps the file's values. */
int rt_scene_load_text(const char* path, int32_t width, int32_t height, int32_t samples, int32_t device, RtScene** out);
/* Dispatch on the file extension: ".txt" -> rt_scene_load_text, anything else -> rt_scene_load_gltf (main.rs:45). */
int rt_scene_load(const char* path, int32_t width, int32_t height, int32_t samples, int32_t device, RtScene** out);
/* Replaces constructing a `Scene` by hand (scene.rs:22-39): a caller that already has the flat primitives
 * (e.g. the Rust host after its own convert_gltf_to_scene) hands them over; arrays are copied. */
int rt_scene_create(const RtSceneDesc* desc, int32_t device, RtScene** out);
/* Same with general primitives (Object3D + Shape3D::Box, geometry.rs:27-46; own-spec shapes and material, see RtSceneDesc2).
 * Null extension arrays mean: all triangles / zero positions / identity rotations / ior 1 / RT_MATERIAL_PBR. */
int rt_scene_create2(const RtSceneDesc2* desc, int32_t device, RtScene** out);
void rt_scene_destroy(RtScene* scene);
/* rt_scene_destroy parks the scene's big frame buffers (per-chunk radiance layers, accumulator, result bytes) in a small process-wide
 * cache so that the next scene on the same device does not pay cudaMalloc + cudaFree of ~0.7 GB per frame (at most 8 blocks are
 * kept).  This call frees them all. */
int rt_release_device_cache(void);
/* Host-side view of the flat scene (pointers stay valid until rt_scene_destroy) -- what the loader produced. */
int rt_scene_get_desc(const RtScene* scene, RtSceneDesc* out);
int rt_scene_get_desc2(const RtScene* scene, RtSceneDesc2* out);   /* extension arrays are null for a triangles-only scene */
int rt_scene_info(const RtScene* scene, RtSceneInfo* out);
/* Change samples / image size without re-uploading geometry (main.rs:39-41 are per-run arguments). */
int rt_scene_set_frame(RtScene* scene, int32_t width, int32_t height, int32_t samples);
/* Device BVH dump for the containment check (bvh.rs:299-322).  nodes: n_nodes x 14 floats =
 * child0 min xyz,max xyz, child1 min xyz,max xyz, child0 ref, child1 ref (refs as floats of the int encoding:
 * >= 0 inner node index, < 0 leaf ~((first << 3) | (count-1))); tri_order: n_tris original triangle ids. */
int rt_scene_get_bvh(const RtScene* scene, float* nodes, int32_t* tri_order);
/* The QUANTISED copy of the same tree that triangle scenes too large for shared memory walk on the device (32 bytes per pair node:
 * words 0 / 1 = x minima / maxima of child 0 (low 16 bits) and child 1 (high 16 bits), 2 / 3 = y, 4 / 5 = z, 6 / 7 = the child references
 * with inner nodes as byte offsets index * 32); plane = grid6[axis] + q * grid6[3 + axis].  *present = 0 (and nothing written) for scenes
 * that keep the full-precision nodes only.  For the containment check of the quantisation (every grid box encloses its f32 box). */
int rt_scene_get_quantised_bvh(const RtScene* scene, uint32_t* words, float* grid6, int32_t* present);

/* ---- the hot path ---------------------------------------------------------------------------------------- */
/* Replaces render_scene(&scene) -> Vec<u8>  (rendering.rs:21-69): W*H*3 bytes, RGB8, row-major, row 0 = top.
 * Host output buffer; the device->host copy is inside the call. */
int rt_render(RtScene* scene, const RtRenderParams* params, uint8_t* rgb_out, RtStats* stats);
/* Same estimator, but returns the per-pixel mean linear radiance before color_to_pixel (W*H*3 floats, host).
 * Exists for the statistical parity tests. */
int rt_render_linear(RtScene* scene, const RtRenderParams* params, float* rgb_linear_out, RtStats* stats);
/* Multi-GPU building block (one process per GPU): ADDS the radiance SUMS of samples [begin,end) into a
 * device-resident accumulator of W*H*4 floats (rgb + sample count) on `stream` (a cudaStream_t, may be null).
 * The caller reduces the accumulators across ranks (NCCL sum) and resolves on one rank. */
int rt_render_accumulate_device(RtScene* scene, const RtRenderParams* params, float* accum_dev, void* stream, RtStats* stats);
/* color_to_pixel (rendering.rs:250-262) over a device accumulator: mean = rgb / count, ACES, gamma 1/2.2,
 * round, saturating u8.  rgb_dev: W*H*3 bytes on the device. */
int rt_resolve_device(const float* accum_dev, int32_t width, int32_t height, uint8_t* rgb_dev, void* stream);
/* One frame on n GPUs of one box from ONE process (north_star: "main.rs (device and multi-GPU orchestration)"): scenes[g]
 * holds the same scene on device g (rt_scene_load_gltf / rt_scene_create with device = g); GPU g renders samples
 * [g*S/n, (g+1)*S/n) of every pixel, the W*H*4 float accumulators are summed on scenes[0]'s device with one ncclReduce over
 * NVLink (libnccl is loaded with dlopen; RT_NCCL_LIB overrides its name; without NCCL the sum uses peer copies), resolved
 * there (color_to_pixel) and copied to rgb_out.  n = 1 is rt_render.  Same bytes as rt_render up to FP32 summation order.
 * With fewer samples than GPUs the frame is sharded by interleaved 8x4-pixel tiles instead (RtRenderParams.tile_shard_*). */
int rt_render_multi(RtScene* const* scenes, int32_t n_scenes, const RtRenderParams* params, uint8_t* rgb_out, RtStats* stats);
/* Optional: create the NCCL communicators (seconds) / enable peer access for this device list ahead of the first frame. */
int rt_multi_init(RtScene* const* scenes, int32_t n_scenes);

/* Nearest-hit query of the finite-primitive BVH for caller-supplied rays: replaces
 * interesect_with_bvh_nearest_point (bvh.rs:231-247) for n rays (rays: n x 6 doubles, origin + direction).
 * precision 32 = the production FP32 traversal the renderer uses; 64 = same traversal with f64 triangle tests.
 * tri_id = ORIGINAL (load-order) triangle index or -1; t = hit distance (inf on miss). */
int rt_trace_primary(RtScene* scene, const double* rays, int64_t n, int32_t precision, int32_t* tri_id, double* t);
/* Same query returning what get_ray_color reads from the hit (rendering.rs:96-101,107): out = n x 9 doubles: t,
 * normal_geometry xyz (world, facing the ray), normal_shading xyz (normalised; object space -- the reference does not rotate
 * it back, geometry.rs:245-249), original primitive index (-1 on a miss), is_outer_to_inner.  FP32 traversal. */
int rt_trace_hits(RtScene* scene, const double* rays, int64_t n, double* out);
/* Camera rays of get_ray_to_pixel (rendering.rs:71-84) for explicit jitter: xy n x 2 ints, xi n x 2 doubles ->
 * rays n x 6 doubles, computed on the device in FP32 exactly as the render kernel does. */
int rt_primary_rays(RtScene* scene, const int32_t* xy, const double* xi, int64_t n, double* rays_out);

/* ---- device unit functions (parity tests of the device samplers / BRDF against the oracle) --------------- */
enum {
    RT_FN_BRDF = 1,          /* in: l3 n3 v3 base3 metallic rough (14)   out: rgb (3)   rendering.rs:133-155     */
    RT_FN_PDF_COSINE = 2,    /* in: n3 l3 (6)                            out: 1         distributions.rs:65-67   */
    RT_FN_PDF_VNDF = 3,      /* in: n3 l3 v3 rough (10)                  out: 1         distributions.rs:276-297 */
    RT_FN_PDF_LIGHT = 4,     /* in: point3 l3 (6)                        out: 1         distributions.rs:160-184 */
    RT_FN_PDF_MIX = 5,       /* in: point3 n3 l3 v3 rough (13)           out: 1         distributions.rs:194-201 */
    RT_FN_SAMPLE_COSINE = 6, /* in: n3 u1 u2 (5)                         out: l3 + sphere3 (6)  :54-63           */
    RT_FN_SAMPLE_VNDF = 7,   /* in: n3 v3 rough u1 u2 (9)                out: l3        :264-274 (same distribution, spherical-cap construction) */
    RT_FN_SAMPLE_LIGHT = 8,  /* in: point3 light_index u v (6)           out: l3        :111-125,151-158         */
    RT_FN_PHILOX = 9,        /* in: pixel sample call seed_lo (4, as exact integers) out: 4 (u32 as float bits)  */
    RT_FN_SAMPLE_LIGHT_GEN = 10, /* in: point3 light_index u1 u2 x01 sign (8)    out: l3   :84-125 incl. the Box arm :86-110 (general scenes) */
    RT_FN_DIELECTRIC = 11    /* in: n3 v3 ior outer u (9)                out: l3 + refracted flag (4)   own spec         */
};
int rt_eval(RtScene* scene_or_null, int32_t fn, const float* in, int64_t n, float* out);

/* ---- output (main.rs:88-95) ------------------------------------------------------------------------------ */
/* "P6\n{W} {H}\n255\n" + bytes.  append != 0 reproduces the reference's append-mode open (main.rs:62-66). */
int rt_write_ppm(const char* path, int32_t width, int32_t height, const uint8_t* rgb, int32_t append);
/* dump_rendered_to_png (src/main.rs:75-86): the same pixels as an 8-bit RGB PNG (self-contained writer, stored deflate). */
int rt_write_png(const char* path, int32_t width, int32_t height, const uint8_t* rgb);

/* ---- memory-safety instrument (compute-sanitizer is not available on every GPU pool) ---------------------------------- */
/* In a library built with -DRT_DEBUG_BOUNDS (`make -C csrc debug` -> _build_dbg/librt_b200.so, same ABI) every shared-memory pool /
 * queue / traversal-stack access, every scene-blob load, every primitive index and every framebuffer store of the kernels is
 * range-checked and violations are counted per kind: counts[0..6] = pool slot, queue, stack, blob offset, primitive index, layer
 * store, chunk index.  reset != 0 zeroes the counters.  A regular build returns RT_ERR_INVALID. */
int rt_debug_bounds_violations(int32_t device, uint64_t* counts, int32_t reset);

/* ---- roofline support ------------------------------------------------------------------------------------ */
/* FFMA issue micro-benchmark on `device`: measured FP32 TFLOP/s (2 flop per FMA lane) and SM clock in MHz. */
int rt_measure_fp32_peak(int32_t device, double* tflops, double* sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* RT_API_H */
