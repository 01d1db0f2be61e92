// rt_api.cu -- implementation of the C ABI declared in include/rt_api.h: scene ingest (flatten + BVH build +
// upload), the render entry points that replace render_scene (rendering.rs:21-69), the nearest-hit query and the
// test/roofline helpers.  No CPU fallback: every compute entry point needs a CUDA device.
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rt_api.h"
#include "host_scene.h"
#include "rt_kernels.h"

// ABI layout guards: rust/rt-sys/src/lib.rs asserts the same numbers (size_of / offset_of), tests/test_rust_binding.py compares the
// field lists -- the binding the reference's render_scene (src/rendering.rs:21, call site src/main.rs:55) links against must not drift.
#include <cstddef>
static_assert(sizeof(RtSceneDesc) == 192 && offsetof(RtSceneDesc, bg_color) == 16 && offsetof(RtSceneDesc, camera_fov_x) == 136 &&
              offsetof(RtSceneDesc, n_tris) == 152 && offsetof(RtSceneDesc, tri_v) == 160, "RtSceneDesc layout");
static_assert(sizeof(RtSceneDesc2) == 232 && offsetof(RtSceneDesc2, shape_kind) == 192, "RtSceneDesc2 layout");
static_assert(sizeof(RtSceneInfo) == 72 && offsetof(RtSceneInfo, device_bytes) == 40 && offsetof(RtSceneInfo, bvh_build_ms) == 56, "RtSceneInfo layout");
static_assert(sizeof(RtRenderParams) == 40 && offsetof(RtRenderParams, sample_begin) == 8, "RtRenderParams layout");
static_assert(sizeof(RtStats) == 152 && offsetof(RtStats, kernel_ms) == 80 && offsetof(RtStats, kernel) == 96 && offsetof(RtStats, render_ms) == 128, "RtStats layout");

namespace {

thread_local std::string g_last_error;
int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
#define CUDA_TRY(expr)                                                                                              \
    do {                                                                                                            \
        cudaError_t _e = (expr);                                                                                    \
        if (_e != cudaSuccess) return fail(RT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

const double kEps = 0.00001;  // geometry.rs:49

}  // namespace

struct RtScene {
    rtb::HostScene host;
    rtb::FlatBvh bvh, light_bvh;
    std::vector<int32_t> light_ids;        // load order ids of emissive triangles (gltf_to_scene.rs:240)
    std::vector<int32_t> light_order;      // device order of the lights -> original id
    std::vector<int32_t> plane_ids;        // load-order ids of the infinite primitives (scene.rs:37)
    rtd::SceneLayout L;
    std::vector<char> blob_host;
    std::vector<double> tri_d_host;        // BVH-ordered a, e1, e2 in f64 (precision-64 queries)
    int device = 0, sms = 0;
    int n_mats = 0, validate_failures = 0;
    int bvh_builder = 0;                   // 0 = host SAH sweep, 1 = GPU LBVH (RT_BVH_BUILDER=gpu)
    double bvh_build_ms = 0.0;
    uint32_t stack_entries = 16;
    bool use_smem = false;
    // device state
    char* blob_dev = nullptr;
    double* tri_d_dev = nullptr;
    float4* layers = nullptr; size_t layers_cap = 0;
    float4* accum = nullptr; size_t accum_cap = 0;
    uint8_t* rgb_dev = nullptr; size_t rgb_cap = 0;
    float* lin_dev = nullptr; size_t lin_cap = 0;
    unsigned int* work_counter = nullptr;
    int32_t* chunk_table = nullptr;        // chunk schedule of the current launch (RenderArgs::chunk_begin)
    unsigned long long* stats_dev = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

struct BlobWriter {
    std::vector<char>& buf;
    uint32_t add(const void* data, size_t bytes, size_t align = 16) {
        size_t off = (buf.size() + (align - 1)) & ~(align - 1);
        buf.resize(off + ((bytes + 15u) & ~(size_t)15u), 0);
        if (bytes) std::memcpy(buf.data() + off, data, bytes);
        return (uint32_t)off;
    }
};

inline float i2f(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }

// Ray-independent part of the triangle test (rt_device.cuh TriTest), evaluated in f64: unit normal N, vertex a, and the
// barycentric gradients Au = (e2 x Nu)/|Nu|^2, Av = (Nu x e1)/|Nu|^2 with Nu = e1 x e2.  out12 = N, a, Au, Av.
void tri_test_record(const double* v, float* out12) {
    double e1[3], e2[3];
    for (int a = 0; a < 3; ++a) { e1[a] = v[3 + a] - v[a]; e2[a] = v[6 + a] - v[a]; }
    const double nu[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const double n2 = nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2], len = std::sqrt(n2);
    const double au[3] = {e2[1] * nu[2] - e2[2] * nu[1], e2[2] * nu[0] - e2[0] * nu[2], e2[0] * nu[1] - e2[1] * nu[0]};
    const double av[3] = {nu[1] * e1[2] - nu[2] * e1[1], nu[2] * e1[0] - nu[0] * e1[2], nu[0] * e1[1] - nu[1] * e1[0]};
    for (int a = 0; a < 3; ++a) {
        out12[a] = (float)(nu[a] / len); out12[3 + a] = (float)v[a];
        out12[6 + a] = (float)(au[a] / n2); out12[9 + a] = (float)(av[a] / n2);
    }
}

// FlatBvh (box_a/box_b/box_c/child, host_scene.h) -> the device's octant-ordered pair nodes (rt_device.cuh, RT_NODE_BYTES):
// per axis (c0.min c1.min c0.max c1.max | c0.max c1.max c0.min c1.min), then the two child references.
// Outward rounding of a box plane whose low mantissa byte is replaced by `payload`: the result is <= v (lower = true) or
// >= v (lower = false) and differs from v by less than 2^-14 relative.
float plane_with_payload(float v, uint32_t payload, bool lower) {
    if (!lower) return -plane_with_payload(-v, payload, true);
    uint32_t bits; std::memcpy(&bits, &v, 4);
    uint32_t out;
    if (!(bits & 0x80000000u)) {                             // v >= +0: shrink the magnitude
        uint32_t c = (bits & ~0xffu) | payload;
        if (c > bits) c = bits >= 0x100u ? c - 0x100u : (0x80000000u | payload);   // tiny positive: any negative is <= v
        out = c;
    } else {                                                 // v < 0: grow the magnitude
        const uint32_t mag = bits & 0x7fffffffu;
        uint32_t c = (mag & ~0xffu) | payload;
        if (c < mag) c += 0x100u;
        out = 0x80000000u | c;
    }
    float r; std::memcpy(&r, &out, 4);
    return r;
}
bool refs_packable(const rtb::FlatBvh& b) { return b.n_nodes < 32768 && b.tri_order.size() < 4096; }

// stride_bytes: RT_NODE_BYTES (112).  Inner children are stored as BYTE offsets, so global-memory walks accept any stride; padding
// the nodes to one 128-byte line each was tried for the large meshes and lost 10 % (more L1 footprint, power-of-two set conflicts).
std::vector<float> octant_nodes(const rtb::FlatBvh& b, size_t stride_bytes) {
    const size_t wpn = stride_bytes / 4;
    std::vector<float> out((size_t)std::max(b.n_nodes, 1) * wpn, 0.f);
    for (int n = 0; n < b.n_nodes; ++n) {
        const float* A = &b.box_a[(size_t)n * 4]; const float* B = &b.box_b[(size_t)n * 4]; const float* C = &b.box_c[(size_t)n * 4];
        const float lo0[3] = {A[0], A[2], C[0]}, hi0[3] = {A[1], A[3], C[1]}, lo1[3] = {B[0], B[2], C[2]}, hi1[3] = {B[1], B[3], C[3]};
        float* o = &out[(size_t)n * wpn];
        for (int a = 0; a < 3; ++a) {
            float* q = o + a * 8;
            q[0] = lo0[a]; q[1] = lo1[a]; q[2] = hi0[a]; q[3] = hi1[a];
            q[4] = hi0[a]; q[5] = hi1[a]; q[6] = lo0[a]; q[7] = lo1[a];
        }
        if (refs_packable(b)) {            // 16-bit references (node index / leaf code) in the low bytes of the x planes
            const uint32_t r0 = (uint32_t)b.child[(size_t)n * 2] & 0xffffu, r1 = (uint32_t)b.child[(size_t)n * 2 + 1] & 0xffffu;
            const uint32_t pay[4] = {r0 & 0xffu, r0 >> 8, r1 & 0xffu, r1 >> 8};
            o[0] = plane_with_payload(o[0], pay[0], true);  o[1] = plane_with_payload(o[1], pay[1], true);     // c0.min c1.min
            o[2] = plane_with_payload(o[2], pay[2], false); o[3] = plane_with_payload(o[3], pay[3], false);    // c0.max c1.max
            o[4] = plane_with_payload(o[4], pay[0], false); o[5] = plane_with_payload(o[5], pay[1], false);    // c0.max c1.max
            o[6] = plane_with_payload(o[6], pay[2], true);  o[7] = plane_with_payload(o[7], pay[3], true);     // c0.min c1.min
        }
        for (int c = 0; c < 2; ++c) {      // inner children as byte offsets (index * RT_NODE_BYTES), leaf codes unchanged
            const int32_t r = b.child[(size_t)n * 2 + (size_t)c];
            o[24 + c] = i2f(r >= 0 ? r * (int32_t)stride_bytes : r);
        }
    }
    return out;
}

// FlatBvh -> quantised 32-byte pair nodes (rt_device.cuh pair_step_quant).  Grid: 16 bits per axis over the union of the finite
// child boxes, the scene box mapped to [2, 65533]; a minimum is moved down and a maximum up until the plane the DEVICE reconstructs
// (qorg + q * qcell, evaluated here with the same float constants) lies at least 1.5 cells outside the true plane -- the slack
// that absorbs the FP32 rounding of fma(q, qcell / d, -(o - qorg) / d) for ray origins within ~100 scene diameters.
std::vector<uint32_t> quant_nodes(const rtb::FlatBvh& b, float* qorg, float* qcell) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    auto slot_box = [&b](int node, int s, float* mn, float* mx) {
        const float* A = &b.box_a[(size_t)node * 4]; const float* B = &b.box_b[(size_t)node * 4]; const float* C = &b.box_c[(size_t)node * 4];
        const float* xy = s == 0 ? A : B;
        mn[0] = xy[0]; mx[0] = xy[1]; mn[1] = xy[2]; mx[1] = xy[3]; mn[2] = C[s * 2]; mx[2] = C[s * 2 + 1];
    };
    for (int n = 0; n < b.n_nodes; ++n)
        for (int s = 0; s < 2; ++s) {
            float mn[3], mx[3];
            slot_box(n, s, mn, mx);
            if (mn[0] > 1e29f) continue;                       // the far-away filler of a one-leaf tree
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], (double)mn[a]); hi[a] = std::max(hi[a], (double)mx[a]); }
        }
    for (int a = 0; a < 3; ++a) {
        if (lo[a] > hi[a]) { lo[a] = 0.0; hi[a] = 0.0; }
        qcell[a] = (float)std::max((hi[a] - lo[a]) / 65531.0, 1e-12);
        qorg[a] = (float)(lo[a] - 2.0 * (double)qcell[a]);
    }
    auto q_of = [&](int a, double p, bool lower) -> uint32_t {
        const double cell = (double)qcell[a], org = (double)qorg[a];
        double q = lower ? std::floor((p - org) / cell) - 2.0 : std::ceil((p - org) / cell) + 2.0;
        q = std::min(std::max(q, 0.0), 65535.0);
        if (lower) while (q > 0.0 && org + q * cell > p - 1.5 * cell) q -= 1.0;
        else while (q < 65535.0 && org + q * cell < p + 1.5 * cell) q += 1.0;
        return (uint32_t)q;
    };
    std::vector<uint32_t> out((size_t)std::max(b.n_nodes, 1) * 8, 0u);
    for (int n = 0; n < b.n_nodes; ++n) {
        uint32_t* o = &out[(size_t)n * 8];
        uint32_t qmn[2][3], qmx[2][3];
        for (int s = 0; s < 2; ++s) {
            float mn[3], mx[3];
            slot_box(n, s, mn, mx);
            for (int a = 0; a < 3; ++a) {
                if (mn[0] > 1e29f) { qmn[s][a] = 65535u; qmx[s][a] = 65535u; }      // filler: a point at the far corner
                else { qmn[s][a] = q_of(a, (double)mn[a], true); qmx[s][a] = q_of(a, (double)mx[a], false); }
            }
        }
        for (int a = 0; a < 3; ++a) { o[a * 2] = qmn[0][a] | (qmn[1][a] << 16); o[a * 2 + 1] = qmx[0][a] | (qmx[1][a] << 16); }
        for (int c = 0; c < 2; ++c) {
            const int32_t r = b.child[(size_t)n * 2 + (size_t)c];
            o[6 + c] = (uint32_t)(r >= 0 ? r * 32 : r);
        }
    }
    return out;
}

int env_int(const char* name, int dflt) { const char* v = std::getenv(name); return v && *v ? std::atoi(v) : dflt; }
double env_double(const char* name, double dflt) { const char* v = std::getenv(name); return v && *v ? std::atof(v) : dflt; }

// World-space vertices of triangle i: q * v + position for a general scene (the reference rotates the RAY into object space
// instead, geometry.rs:201-214 -- the same hit at the same t), the stored vertices otherwise.
void world_triangle(const rtb::HostScene& h, int i, double* v9) {
    const double* v = &h.tri_v[(size_t)i * 9];
    if (!h.general()) { for (int k = 0; k < 9; ++k) v9[k] = v[k]; return; }
    double m[9];
    rtb::quat_to_matrix(&h.rotation[(size_t)i * 4], m);
    const double* p = &h.position[(size_t)i * 3];
    for (int c = 0; c < 3; ++c)
        for (int a = 0; a < 3; ++a) v9[c * 3 + a] = m[a * 3] * v[c * 3] + m[a * 3 + 1] * v[c * 3 + 1] + m[a * 3 + 2] * v[c * 3 + 2] + p[a];
}

// calculate_aabb_for_object (aabb.rs:53-94): triangles: padded bounds of the (world-space) vertices; boxes / ellipsoids: bounds
// of the 8 rotated + translated corners of the EPS-padded object-space box +-(s + EPS).
rtb::BoxD object_box(const rtb::HostScene& h, int i) {
    const int kind = h.general() ? h.kind[(size_t)i] : RT_SHAPE_TRIANGLE;
    if (kind == RT_SHAPE_TRIANGLE) { double v[9]; world_triangle(h, i, v); return rtb::tri_box_d(v); }
    double m[9];
    rtb::quat_to_matrix(&h.rotation[(size_t)i * 4], m);
    const double* p = &h.position[(size_t)i * 3];
    const double* s = &h.tri_v[(size_t)i * 9];
    rtb::BoxD b;
    for (int a = 0; a < 3; ++a) { b.mn[a] = std::numeric_limits<double>::infinity(); b.mx[a] = -b.mn[a]; }
    for (int c = 0; c < 8; ++c) {
        const double x = ((c & 1) ? 1.0 : -1.0) * (s[0] + kEps), y = ((c & 2) ? 1.0 : -1.0) * (s[1] + kEps), z = ((c & 4) ? 1.0 : -1.0) * (s[2] + kEps);
        for (int a = 0; a < 3; ++a) {
            const double w = m[a * 3] * x + m[a * 3 + 1] * y + m[a * 3 + 2] * z + p[a];
            b.mn[a] = std::min(b.mn[a], w); b.mx[a] = std::max(b.mx[a], w);
        }
    }
    return b;
}

// 64-byte device record of primitive i of a general scene (rt_device.cuh "general primitives").
void prim_record(const rtb::HostScene& h, int i, float* out16) {
    const int kind = h.kind[(size_t)i];
    if (kind == RT_SHAPE_TRIANGLE) {
        double v[9]; world_triangle(h, i, v);
        float rec[12];
        tri_test_record(v, rec);
        double e1[3], e2[3];
        for (int a = 0; a < 3; ++a) { e1[a] = v[3 + a] - v[a]; e2[a] = v[6 + a] - v[a]; }
        const double nu[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double area = 0.5 * std::sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
        for (int q = 0; q < 4; ++q) for (int a = 0; a < 3; ++a) out16[q * 4 + a] = rec[q * 3 + a];
        out16[3] = i2f(RT_SHAPE_TRIANGLE); out16[7] = (float)(1.0 / area); out16[11] = 0.f; out16[15] = 0.f;   // 1/area: get_local_pdf distributions.rs:76-79
        return;
    }
    double m[9];
    rtb::quat_to_matrix(&h.rotation[(size_t)i * 4], m);
    const double* p = &h.position[(size_t)i * 3];
    const double* s = &h.tri_v[(size_t)i * 9];
    for (int r = 0; r < 3; ++r) for (int a = 0; a < 3; ++a) out16[r * 4 + a] = (float)m[a * 3 + r];   // rows of R^T = columns of R
    for (int a = 0; a < 3; ++a) out16[12 + a] = (float)p[a];
    out16[3] = i2f(kind); out16[7] = (float)s[0]; out16[11] = (float)s[1]; out16[15] = (float)s[2];
}

// Flatten the host scene into the device blob (rt_device.cuh SceneLayout).
int flatten_scene(RtScene* s) {
    const rtb::HostScene& h = s->host;
    const bool gen = h.general();
    const int n_all = h.n_tris();
    rtb::BvhBuildParams bp;
    // measured on B200 (tools/sweep.py): leaves of <= 2 triangles (tested side by side by the kernel) and a box test priced at
    // 2 triangle tests give the fastest trees; the reference's own limit is 4 (bvh.rs:89)
    bp.max_leaf_size = env_int("RT_BVH_MAX_LEAF", 2);
    bp.traversal_cost = env_double("RT_BVH_TRAV_COST", 2.0);
    bp.reinsertion = env_int("RT_BVH_REINSERT", 1) != 0;   // + subtree re-insertion passes (+0.5..2 %)
    bp.size_split = !gen && env_int("RT_BVH_SIZE_SPLIT", 1) != 0;   // triangle meshes: the sweep also tries "the k largest primitives | the rest" (walls vs the mesh inside)
    bp.agglomerative = env_int("RT_BVH_AGGLO", 1) != 0;   // sets of <= 512 primitives: bottom-up clustering (measured +4.7 % on practice7_4 over the sweep)
    if (bp.max_leaf_size < 1) bp.max_leaf_size = 1;
    if (bp.max_leaf_size > 8) bp.max_leaf_size = 8;
    // finite primitives -> BVH (scene.rs:33); planes -> infinite_primitives (scene.rs:37)
    std::vector<int32_t> all;
    s->plane_ids.clear();
    for (int i = 0; i < n_all; ++i) {
        if (gen && h.kind[(size_t)i] == RT_SHAPE_PLANE) s->plane_ids.push_back(i);
        else all.push_back(i);
    }
    const int n = (int)all.size(), n_planes = (int)s->plane_ids.size();
    std::vector<rtb::BoxD> boxes((size_t)n_all);
    for (int32_t id : all) boxes[(size_t)id] = object_box(h, id);
    // builder (RT_BVH_BUILDER = host | gpu | auto, default auto).  host: bottom-up clustering + re-insertion for sets of <= 512 primitives,
    // else the full-sweep SAH whose top (the 128 largest subtrees) is rebuilt bottom-up -- the best trees (practice7_2: 1 524 Msamples/s)
    // after 0.6 s of building.  gpu: Morton-code LBVH (leaves <= 2) whose top is rebuilt twice on the host (SAH sweep over ~8 192 subtrees,
    // then bottom-up over 128): 1 481 Msamples/s after 35 ms.  auto = host below RT_BVH_GPU_MIN_TRIS (1 000 000) triangles, where the
    // host's O(n log^2 n) stays within seconds.  Host-only scenes, general-primitive scenes and CUDA failures take the host builder.
    const char* which = std::getenv("RT_BVH_BUILDER");
    const bool want_gpu = which && *which && std::strcmp(which, "auto") != 0 ? std::strcmp(which, "gpu") == 0 : n >= env_int("RT_BVH_GPU_MIN_TRIS", 1000000);
    s->bvh_builder = 0; s->bvh_build_ms = 0.0;
    bool built = false;
    if (want_gpu && s->device >= 0 && !gen) {
        std::string gerr;
        built = rtb::build_bvh_gpu(h.tri_v.data(), n, all, bp, s->device, &s->bvh, &s->bvh_build_ms, &gerr);
        if (built) {
            s->bvh_builder = 1;
            const auto t0 = std::chrono::steady_clock::now();
            rtb::regraft_top_sah(&s->bvh, env_int("RT_BVH_TOP_SAH", 8192), bp);   // (0 = keep the Morton-code top)
            // ... and the top of THAT tree once more over <= 512 subtrees, which the bottom-up clustering of build_bvh handles (another +7..9 %)
            if (env_int("RT_BVH_TOP_AGGLO", 128) > 0) rtb::regraft_top_sah(&s->bvh, std::min(512, env_int("RT_BVH_TOP_AGGLO", 128)), bp);
            s->bvh_build_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
    }
    if (!built) {
        const auto t0 = std::chrono::steady_clock::now();
        rtb::build_bvh(boxes, all, bp, &s->bvh);
        // triangle meshes too big for the bottom-up builder get it for the TOP of their tree (the 128 largest subtrees): practice7_2 1 312 ->
        // 1 523 Msamples/s.  Not for general-primitive scenes: the heavily overlapping boxes / ellipsoids of working.txt lose 4 % with it.
        if (!gen && n > 512 && env_int("RT_BVH_TOP_AGGLO", 128) > 0) rtb::regraft_top_sah(&s->bvh, std::min(512, env_int("RT_BVH_TOP_AGGLO", 128)), bp);
        s->bvh_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (s->bvh.depth + 3 > 80) {                               // degenerate input (hundreds of coincident primitives make every split cost the same and
        rtb::BvhBuildParams mp = bp; mp.force_median = true;   // the trees chain-shaped): fall back to median splits, depth <= log2(n) + 2
        rtb::build_bvh(boxes, all, mp, &s->bvh);
        s->bvh_builder = 0;
    }
    s->validate_failures = rtb::validate_flat_bvh(s->bvh, boxes);

    // lights: emission.norm() > EPS (gltf_to_scene.rs:240), finite primitives only
    s->light_ids.clear();
    for (int32_t i : all) {
        const double* e = &h.tri_emission[(size_t)i * 3];
        if (std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) > kEps) s->light_ids.push_back(i);
    }
    const int n_lights = (int)s->light_ids.size();
    const bool use_light_bvh = n_lights > RT_BRUTE_LIGHTS;
    if (use_light_bvh) {
        rtb::build_bvh(boxes, s->light_ids, bp, &s->light_bvh);
        s->validate_failures += rtb::validate_flat_bvh(s->light_bvh, boxes);
        s->light_order = s->light_bvh.tri_order;
    } else {
        s->light_bvh = rtb::FlatBvh();
        s->light_order = s->light_ids;
    }

    // materials: per-primitive in the reference (scene.rs:13-20); deduplicated by value for the device table
    std::map<std::vector<double>, int> mat_index;
    std::vector<float> mat0, mat1, mat2;
    std::vector<int32_t> tri_mat((size_t)n_all);
    bool has_dielectric = false;
    for (int i = 0; i < n_all; ++i) {
        std::vector<double> key(10, 0.0);
        for (int k = 0; k < 5; ++k) key[(size_t)k] = h.tri_material[(size_t)i * 5 + (size_t)k];
        for (int k = 0; k < 3; ++k) key[(size_t)(5 + k)] = h.tri_emission[(size_t)i * 3 + (size_t)k];
        key[8] = gen ? h.ior[(size_t)i] : 1.0; key[9] = gen ? (double)h.mat_kind[(size_t)i] : 0.0;
        if (key[9] == (double)RT_MATERIAL_DIELECTRIC) has_dielectric = true;
        auto it = mat_index.find(key);
        if (it == mat_index.end()) {
            int id = (int)mat_index.size();
            mat_index[key] = id;
            mat0.insert(mat0.end(), {(float)key[0], (float)key[1], (float)key[2], (float)key[3]});
            mat1.insert(mat1.end(), {(float)key[5], (float)key[6], (float)key[7], (float)key[4]});
            mat2.insert(mat2.end(), {(float)key[8], i2f((int32_t)key[9]), 0.f, 0.f});
            tri_mat[(size_t)i] = id;
        } else tri_mat[(size_t)i] = it->second;
    }
    if (mat0.empty()) { mat0.assign(4, 0.f); mat1.assign(4, 0.f); mat2.assign(4, 0.f); }
    s->n_mats = (int)mat_index.size();

    // device order of the primitives: the finite ones in BVH order, then the planes
    std::vector<int32_t> dev_order = s->bvh.tri_order;
    dev_order.insert(dev_order.end(), s->plane_ids.begin(), s->plane_ids.end());
    const int n_dev = (int)dev_order.size();
    const int nt = n_dev > 0 ? n_dev : 1;   // an empty scene keeps one degenerate triangle (never hit: det == 0)
    std::vector<float> tri_a((size_t)nt * 4, 0.f), tri_e1((size_t)nt * 4, 0.f), tri_e2((size_t)nt * 4, 0.f);
    std::vector<float> tri_t((size_t)(gen ? 1 : nt) * 12, std::numeric_limits<float>::quiet_NaN());
    std::vector<float> prims((size_t)(gen ? nt : 1) * 16, std::numeric_limits<float>::quiet_NaN());
    std::vector<float> sh_n0((size_t)nt * 4, 0.f), sh_dn1((size_t)nt * 4, 0.f), sh_dn2((size_t)nt * 4, 0.f), sh_ng((size_t)nt * 4, 0.f);
    s->tri_d_host.assign((size_t)nt * 9, 0.0);
    sh_dn1[3] = i2f(-1);
    if (gen && n_dev == 0) prims[3] = i2f(RT_SHAPE_TRIANGLE);      // the degenerate filler: NaN triangle, never hit
    for (int k = 0; k < n_dev; ++k) {
        const int id = dev_order[(size_t)k];
        sh_n0[(size_t)k * 4 + 3] = i2f(tri_mat[(size_t)id]);
        sh_dn1[(size_t)k * 4 + 3] = i2f(id);
        if (gen) prim_record(h, id, &prims[(size_t)k * 16]);
        if (gen && h.kind[(size_t)id] != RT_SHAPE_TRIANGLE) continue;
        double v[9];
        world_triangle(h, id, v);
        const double* nn = &h.tri_n[(size_t)id * 9];
        double e1[3], e2[3];
        for (int a = 0; a < 3; ++a) { e1[a] = v[3 + a] - v[a]; e2[a] = v[6 + a] - v[a]; }
        double ng[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double len = std::sqrt(ng[0] * ng[0] + ng[1] * ng[1] + ng[2] * ng[2]);
        if (!gen) tri_test_record(v, &tri_t[(size_t)k * 12]);
        for (int a = 0; a < 3; ++a) {
            tri_a[(size_t)k * 4 + (size_t)a] = (float)v[a];
            tri_e1[(size_t)k * 4 + (size_t)a] = (float)e1[a];
            tri_e2[(size_t)k * 4 + (size_t)a] = (float)e2[a];
            sh_n0[(size_t)k * 4 + (size_t)a] = (float)nn[a];                       // general scenes: OBJECT-space normals (normal_shading is
            sh_dn1[(size_t)k * 4 + (size_t)a] = (float)(nn[3 + a] - nn[a]);        // not rotated back, geometry.rs:245-249)
            sh_dn2[(size_t)k * 4 + (size_t)a] = (float)(nn[6 + a] - nn[a]);
            sh_ng[(size_t)k * 4 + (size_t)a] = (float)(ng[a] / len);
            s->tri_d_host[(size_t)k * 9 + (size_t)a] = v[a];
            s->tri_d_host[(size_t)k * 9 + 3 + (size_t)a] = e1[a];
            s->tri_d_host[(size_t)k * 9 + 6 + (size_t)a] = e2[a];
        }
    }
    // light arrays (device light order)
    const int nl = n_lights > 0 ? n_lights : 1;
    std::vector<float> lt_a((size_t)nl * 4, 0.f), lt_e1((size_t)nl * 4, 0.f), lt_e2((size_t)nl * 4, 0.f);
    std::vector<float> lt_t((size_t)nl * 16, std::numeric_limits<float>::quiet_NaN());
    std::vector<float> lt_g((size_t)(gen ? nl : 1) * 16, std::numeric_limits<float>::quiet_NaN());
    if (gen) lt_g[3] = i2f(RT_SHAPE_TRIANGLE);
    for (int k = 0; k < n_lights; ++k) {
        const int id = s->light_order[(size_t)k];
        if (gen) prim_record(h, id, &lt_g[(size_t)k * 16]);
        if (gen && h.kind[(size_t)id] != RT_SHAPE_TRIANGLE) continue;
        double v[9];
        world_triangle(h, id, v);
        double e1[3], e2[3];
        for (int a = 0; a < 3; ++a) { e1[a] = v[3 + a] - v[a]; e2[a] = v[6 + a] - v[a]; }
        double ng[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double len = std::sqrt(ng[0] * ng[0] + ng[1] * ng[1] + ng[2] * ng[2]);
        for (int a = 0; a < 3; ++a) {
            lt_a[(size_t)k * 4 + (size_t)a] = (float)v[a];
            lt_e1[(size_t)k * 4 + (size_t)a] = (float)e1[a];
            lt_e2[(size_t)k * 4 + (size_t)a] = (float)e2[a];
        }
        float rec[12];
        tri_test_record(v, rec);
        for (int q = 0; q < 4; ++q) for (int a = 0; a < 3; ++a) lt_t[(size_t)k * 16 + (size_t)q * 4 + (size_t)a] = rec[q * 3 + a];
        lt_t[(size_t)k * 16 + 3] = (float)(1.0 / (len * 0.5));    // 1/area, get_local_pdf distributions.rs:76-79
        lt_t[(size_t)k * 16 + 7] = lt_t[(size_t)k * 16 + 11] = lt_t[(size_t)k * 16 + 15] = 0.f;
    }

    s->blob_host.clear();
    BlobWriter w{s->blob_host};
    rtd::SceneLayout& L = s->L;
    std::memset(&L, 0, sizeof(L));
    const size_t node_stride = RT_NODE_BYTES;      // (128-byte nodes for global-memory scenes were measured: -10 %)
    const std::vector<float> nodes = octant_nodes(s->bvh, node_stride);
    L.nodes = w.add(nodes.data(), nodes.size() * 4);
    L.tri_t = w.add(tri_t.data(), tri_t.size() * 4);
    if (gen) L.prims = w.add(prims.data(), prims.size() * 4);
    L.tri_a = w.add(tri_a.data(), tri_a.size() * 4);
    L.tri_e1 = w.add(tri_e1.data(), tri_e1.size() * 4);
    L.tri_e2 = w.add(tri_e2.data(), tri_e2.size() * 4);
    L.sh_n0 = w.add(sh_n0.data(), sh_n0.size() * 4);
    L.sh_dn1 = w.add(sh_dn1.data(), sh_dn1.size() * 4);
    L.sh_dn2 = w.add(sh_dn2.data(), sh_dn2.size() * 4);
    L.sh_ng = w.add(sh_ng.data(), sh_ng.size() * 4);
    L.mat0 = w.add(mat0.data(), mat0.size() * 4);
    L.mat1 = w.add(mat1.data(), mat1.size() * 4);
    if (gen) L.mat2 = w.add(mat2.data(), mat2.size() * 4);
    L.lt_a = w.add(lt_a.data(), lt_a.size() * 4);
    L.lt_e1 = w.add(lt_e1.data(), lt_e1.size() * 4);
    L.lt_e2 = w.add(lt_e2.data(), lt_e2.size() * 4);
    L.lt_t = w.add(lt_t.data(), lt_t.size() * 4);
    if (gen) L.lt_g = w.add(lt_g.data(), lt_g.size() * 4);
    if (use_light_bvh) {
        const std::vector<float> lnodes = octant_nodes(s->light_bvh, RT_NODE_BYTES);
        L.lnodes = w.add(lnodes.data(), lnodes.size() * 4);
    }
    // entry 0 = RT_CUR_DONE, one marker per walk; the light-pdf walk runs on top of a ray's live stack (wavefront kernel)
    const int light_extra = use_light_bvh ? s->light_bvh.depth + 2 : 0;
    const int need = s->bvh.depth + 3 + light_extra;
    if (need <= 96) s->stack_entries = (uint32_t)std::max(8, (need + 3) / 4 * 4);
    else return fail(RT_ERR_LIMIT, "BVH depth " + std::to_string(s->bvh.depth) + " exceeds the traversal stack");
    // global-memory triangle scenes: quantised 32-byte nodes for the nearest-hit walk of the render kernel (the octant-ordered pair
    // nodes stay for the f64 parity walk, the per-lane kernel and forced placements)
    const bool too_big_for_smem = !refs_packable(s->bvh) || s->blob_host.size() > (size_t)env_int("RT_SMEM_SCENE_MAX_BYTES", 48 * 1024);
    const int node_format = env_int("RT_NODE_FORMAT", 2);       // 0: octant-ordered pair nodes only, 2: + quantised 32-byte pair nodes (A/B knob)
    if (too_big_for_smem && !gen && node_format == 2 && (uint64_t)s->bvh.n_nodes * 32u <= 0x7fffffffull) {
        const std::vector<uint32_t> qn = quant_nodes(s->bvh, L.qorg, L.qcell);
        bool grid_ok = true;                                   // (a scene with non-finite or astronomically large coordinates keeps the octant nodes)
        for (int a = 0; a < 3; ++a) grid_ok = grid_ok && std::isfinite(L.qorg[a]) && std::isfinite(L.qcell[a]) && L.qcell[a] < 1e30f;
        if (grid_ok) { L.qnodes = w.add(qn.data(), qn.size() * 4, 32); L.quant = 1; }
    }
    if (s->blob_host.size() > 0xffffffffull || (uint64_t)std::max(s->bvh.n_nodes, s->light_bvh.n_nodes) * 128u > 0x7fffffffull)
        return fail(RT_ERR_LIMIT, "scene exceeds the 4 GiB device blob / 2 GiB node array addressed by 32-bit offsets");
    L.total_bytes = (uint32_t)s->blob_host.size();
    L.n_nodes = s->bvh.n_nodes; L.n_tris = n; L.n_mats = s->n_mats; L.n_lights = n_lights;
    L.n_lnodes = s->light_bvh.n_nodes; L.light_bvh = use_light_bvh ? 1 : 0;
    L.max_leaf = s->bvh.max_leaf;
    L.inv_n_lights = n_lights > 0 ? 1.0f / (float)n_lights : 0.0f;
    L.general = gen ? 1 : 0; L.n_planes = n_planes; L.has_dielectric = has_dielectric ? 1 : 0;
    L.flat_scan = gen && n + n_planes <= env_int("RT_FLAT_SCAN_MAX", RT_FLAT_SCAN_MAX) ? 1 : 0;

    const int smem_limit = env_int("RT_SMEM_SCENE_MAX_BYTES", 48 * 1024);
    L.packed_refs = refs_packable(s->bvh) ? 1 : 0;
    s->use_smem = (int)L.total_bytes <= smem_limit && L.packed_refs;   // the shared-memory kernel reads packed references
    return RT_OK;
}

void fill_camera(const rtb::HostScene& h, rtd::Camera* c) {
    for (int a = 0; a < 3; ++a) {
        c->pos[a] = (float)h.camera_position[a]; c->right[a] = (float)h.camera_right[a];
        c->up[a] = (float)h.camera_up[a]; c->fwd[a] = (float)h.camera_forward[a];
    }
    c->tan_x = (float)std::tan(h.camera_fov_x * 0.5);     // rendering.rs:76
    c->tan_y = (float)std::tan(h.camera_fov_y * 0.5);     // rendering.rs:77
    c->two_over_w = (float)(2.0 / (double)h.width); c->two_over_h = (float)(2.0 / (double)h.height);   // rendering.rs:74-75
}

// resident-thread configuration (rt_kernels.cu RT_WAVE): 384 x 2 for scenes staged in shared memory and for the quantised walk; 352 x 2
// for octant nodes read through L1 / L2 -- two warps fewer per SM bring the blocks under 97 KB of shared memory each, which lets the
// driver pick the 196 KB carve-out instead of 228 KB: 60 KB of L1 instead of 28 KB (measured: practice7_2 +4 %, practice7_3 +7 %);
// 320 x 2 for general-primitive scenes that walk the BVH (95 registers), 384 x 2 for the flat scan (80 registers).
void default_wave_cfgs(const RtScene* s, int* cfg_smem, int* cfg_gmem) {
    const int cfg_env = env_int("RT_WAVE_CFG", -1);
    *cfg_smem = cfg_env >= 0 ? cfg_env : (s->L.general && !s->L.flat_scan ? 6 : 2);
    *cfg_gmem = cfg_env >= 0 ? cfg_env : (s->L.general ? 6 : (s->L.quant ? 2 : 5));
}

int upload_scene(RtScene* s) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return fail(RT_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (s->device < 0 || s->device >= count) return fail(RT_ERR_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, s->device));
    CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 5; ++i) CUDA_TRY(cudaEventCreate(&s->ev[i]));
    CUDA_TRY(cudaMalloc((void**)&s->blob_dev, s->blob_host.size()));
    CUDA_TRY(cudaMemcpy(s->blob_dev, s->blob_host.data(), s->blob_host.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc((void**)&s->work_counter, sizeof(unsigned int)));
    CUDA_TRY(cudaMalloc((void**)&s->stats_dev, sizeof(unsigned long long) * rtd::RT_N_STATS));
    // Load the render kernel(s) this scene will run NOW (CUDA loads a kernel lazily at its first use: 10-25 ms for these ~35 KB
    // functions, which a one-frame CLI run would otherwise pay inside its "Rendering took" window -- the reference's window
    // (main.rs:54-58) does not contain program loading either).  Same queries as the launch plan makes; failures are not errors here.
    if (env_int("RT_PRELOAD_KERNELS", 1)) {
        int lanes = 0;
        const int variant = s->L.general ? 3 : env_int("RT_KERNEL", 3);
        const int fmt = s->L.quant ? 2 : (s->L.flat_scan ? 3 : 0);
        int cfg_smem = 0, cfg_gmem = 0;
        default_wave_cfgs(s, &cfg_smem, &cfg_gmem);
        if (variant >= 1 && variant <= 3) {
            if (s->use_smem && rtd::render_resident_lanes(variant, cfg_smem, true, false, s->L.general != 0, fmt, s->L.total_bytes, s->stack_entries, s->sms, &lanes) != cudaSuccess) cudaGetLastError();
            if (rtd::render_resident_lanes(variant, cfg_gmem, false, false, s->L.general != 0, fmt, s->L.total_bytes, s->stack_entries, s->sms, &lanes) != cudaSuccess) cudaGetLastError();
        }
    }
    return RT_OK;
}

int finish_create(RtScene* s, RtScene** out) {
    int rc = flatten_scene(s);
    if (rc == RT_OK && s->device >= 0) rc = upload_scene(s);   // device < 0: host-only scene (loader / BVH inspection)
    if (rc != RT_OK) { rt_scene_destroy(s); return rc; }
    *out = s;
    return RT_OK;
}

// Process-wide cache of the big frame buffers (per-chunk layers, accumulator, result bytes).  A host that renders one scene after
// another (create -> render -> destroy per frame, like bench.py's end-to-end leg or a render service) would otherwise pay a
// cudaMalloc and a cudaFree of ~0.7 GB per frame; cudaFree of such blocks was measured at 10 ms .. 0.9 s on B200.  rt_scene_destroy
// parks the buffers here, the next scene on the same device picks them up.  At most kCacheSlots blocks per process are kept (the
// smallest is released first); rt_release_device_cache() frees them all.
struct CachedBlock { int device; void* p; size_t bytes; };
const size_t kCacheSlots = 8, kCacheMinBytes = 1u << 20;
std::mutex& cache_mutex() { static std::mutex m; return m; }
std::vector<CachedBlock>& cache_blocks() { static std::vector<CachedBlock> v; return v; }

void cache_put(void* p, size_t bytes) {
    if (!p) return;
    int dev = 0;
    if (bytes < kCacheMinBytes || cudaGetDevice(&dev) != cudaSuccess) { cudaFree(p); return; }
    void* evict = nullptr;
    {
        std::lock_guard<std::mutex> lock(cache_mutex());
        std::vector<CachedBlock>& c = cache_blocks();
        c.push_back({dev, p, bytes});
        if (c.size() > kCacheSlots) {                        // over capacity: release the smallest block of THIS device (freeing needs it current;
            long k = -1;                                     // the block just pushed belongs to it, so one always exists)
            for (size_t i = 0; i < c.size(); ++i) if (c[i].device == dev && (k < 0 || c[i].bytes < c[(size_t)k].bytes)) k = (long)i;
            if (k >= 0) { evict = c[(size_t)k].p; c.erase(c.begin() + k); }
        }
    }
    if (evict) cudaFree(evict);
}
// Smallest cached block of the current device with need <= bytes <= 2 * need, or null.
void* cache_take(size_t need, size_t* got) {
    int dev = 0;
    if (need < kCacheMinBytes || cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(cache_mutex());
    std::vector<CachedBlock>& c = cache_blocks();
    long best = -1;
    for (size_t i = 0; i < c.size(); ++i)
        if (c[i].device == dev && c[i].bytes >= need && c[i].bytes <= 2 * need && (best < 0 || c[i].bytes < c[(size_t)best].bytes)) best = (long)i;
    if (best < 0) return nullptr;
    void* p = c[(size_t)best].p; *got = c[(size_t)best].bytes;
    c.erase(c.begin() + best);
    return p;
}

// Grows a device buffer of the scene (the scene's device is current).
template <class T>
int ensure(T** p, size_t* cap, size_t need_elems) {
    if (*cap >= need_elems && *p) return RT_OK;
    if (*p) {
        // the old block may still be read by work queued on ANY stream of this device (rt_render_accumulate_device is asynchronous and
        // runs on the caller's stream): wait like the implicit synchronisation of cudaFree did before the block can change owner
        cudaDeviceSynchronize();
        cache_put(*p, *cap * sizeof(T));
    }
    *p = nullptr; *cap = 0;
    size_t got = 0;
    if (void* c = cache_take(need_elems * sizeof(T), &got)) { *p = (T*)c; *cap = got / sizeof(T); return RT_OK; }
    cudaError_t e = cudaMalloc((void**)p, need_elems * sizeof(T));
    if (e != cudaSuccess) {                                  // out of memory: drop the cache and try once more
        cudaGetLastError();
        rt_release_device_cache();
        e = cudaMalloc((void**)p, need_elems * sizeof(T));
    }
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    *cap = need_elems;
    return RT_OK;
}

struct RenderPlan { rtd::RenderArgs args; bool use_smem; bool stats; int variant; int cfg; };

// Shared front half of the three render entry points: launches the path-tracing kernel into s->layers.
int render_to_layers(RtScene* s, const RtRenderParams* p, cudaStream_t stream, RenderPlan* plan, rtd::KernelInfo* ki) {
    const rtb::HostScene& h = s->host;
    if (!s->blob_dev) return fail(RT_ERR_CUDA, "scene is host-only (created with device < 0): rendering needs a CUDA device, there is no CPU fallback");
    if (h.width <= 0 || h.height <= 0) return fail(RT_ERR_INVALID, "width and height must be positive");
    if (h.samples <= 0) return fail(RT_ERR_INVALID, "samples must be positive (the reference panics on an empty reduce, rendering.rs:60)");
    RtRenderParams dflt;
    std::memset(&dflt, 0, sizeof(dflt));
    if (!p) p = &dflt;
    int s0 = p->sample_begin, s1 = p->sample_end;
    if (s0 == 0 && s1 == 0) s1 = h.samples;
    if (s0 < 0 || s1 > h.samples || s0 >= s1) return fail(RT_ERR_INVALID, "sample range must satisfy 0 <= begin < end <= samples");
    if ((uint64_t)h.width * (uint64_t)h.height > (1ull << 30)) return fail(RT_ERR_LIMIT, "frame larger than 2^30 pixels");
    if (h.width > 65535 || h.height > 65535) return fail(RT_ERR_LIMIT, "frame dimensions above 65535 (pixel coordinates are packed in 16 bits)");
    CUDA_TRY(cudaSetDevice(s->device));

    rtd::RenderArgs& a = plan->args;
    std::memset(&a, 0, sizeof(a));
    a.blob = s->blob_dev; a.L = s->L;
    fill_camera(h, &a.cam);
    for (int k = 0; k < 3; ++k) a.bg[k] = (float)h.bg_color[k];
    a.W = h.width; a.H = h.height; a.ray_depth = h.ray_depth;
    a.max_attempts = std::min(p->max_attempts > 0 ? p->max_attempts : 64, 127);   // every kernel: the wavefront kernel keeps the attempt count in 7 bits
    a.n_comp = s->L.n_lights > 0 ? 3 : 2;                     // rendering.rs:23-31
    a.inv_n_comp = 1.0f / (float)a.n_comp;
    a.s_begin = s0; a.s_end = s1;
    a.tiles_x = (uint32_t)((h.width + 7) / 8);
    const uint32_t tiles_y = (uint32_t)((h.height + 3) / 4);
    a.n_pix_items = a.tiles_x * tiles_y * 32u;
    a.stack_entries = s->stack_entries;
    a.shard_count = p->tile_shard_count > 1 ? (uint32_t)p->tile_shard_count : 1u;
    a.shard_index = a.shard_count > 1 ? (uint32_t)p->tile_shard_index : 0u;
    if (p->tile_shard_count > 1 && (p->tile_shard_index < 0 || p->tile_shard_index >= p->tile_shard_count)) return fail(RT_ERR_INVALID, "tile_shard_index must be in [0, tile_shard_count)");
    a.node_min = std::min(32, std::max(1, env_int("RT_NODE_MIN", 10)));
    a.burst_exit = std::min(32, std::max(1, env_int("RT_BURST_EXIT", 20)));
    a.debug_blob_limit = (uint32_t)std::max(0, env_int("RT_DEBUG_BLOB_LIMIT", 0));
    a.seed_lo = (uint32_t)(p->seed & 0xffffffffu); a.seed_hi = (uint32_t)(p->seed >> 32);
    plan->stats = p->collect_stats != 0;
    // kernel_variant = 10*kernel + scene placement: kernel 0 auto (= 3), 1 = v1 per-lane megakernel, 2 / 3 = v3 warp-local
    // wavefront with while-while / phased trace bursts; placement 0 auto, 1 global memory, 2 shared memory
    plan->use_smem = s->use_smem;
    const int placement = p->kernel_variant % 10, kern = p->kernel_variant / 10;
    if (placement == 1) plan->use_smem = false;
    if (placement == 2) {
        if (!s->L.packed_refs) return fail(RT_ERR_LIMIT, "kernel_variant: scene too large for the shared-memory placement");
        plan->use_smem = true;
    }
    plan->variant = kern == 0 ? (s->L.general ? 3 : env_int("RT_KERNEL", 3)) : kern;
    if (plan->variant < 1 || plan->variant > 3) return fail(RT_ERR_INVALID, "kernel_variant: unknown kernel");
    if (s->L.general && plan->variant != 3) return fail(RT_ERR_INVALID, "kernel_variant: general-primitive scenes (boxes, ellipsoids, planes, object transforms) run on the phased wavefront kernel (3x) only");
    if (h.ray_depth < 0) return fail(RT_ERR_INVALID, "ray_depth must be >= 0");
    if (plan->variant != 1 && (long long)h.ray_depth * std::min(a.max_attempts, 127) + 1 >= (1 << 14))
        return fail(RT_ERR_LIMIT, "ray_depth x max_attempts exceeds the 14-bit Philox call counter of the wavefront kernel (lower max_attempts, or kernel_variant 10)");
    if (plan->variant != 1 && h.ray_depth > 255) return fail(RT_ERR_LIMIT, "ray_depth above 255 (the wavefront kernel packs the remaining depth in 8 bits; kernel_variant 10 has no limit)");

    if (h.ray_depth == 0) {                                    // rendering.rs:93-95: recursion_depth <= 0 is black, the scene is never looked at
        a.n_chunks = 1; a.n_pix_items = 0; a.total_items = 0;
        const size_t n_pix0 = (size_t)h.width * (size_t)h.height;
        int rc0 = ensure(&s->layers, &s->layers_cap, n_pix0);
        if (rc0 != RT_OK) return rc0;
        a.layers = s->layers;
        if (plan->stats) CUDA_TRY(cudaMemsetAsync(s->stats_dev, 0, sizeof(unsigned long long) * rtd::RT_N_STATS, stream));
        CUDA_TRY(rtd::launch_black_layer(s->layers, h.width, h.height, a.tiles_x, a.shard_index, a.shard_count, (float)(s1 - s0), stream));
        std::memset(ki, 0, sizeof(*ki));
        return RT_OK;
    }
    int cfg_smem = 0, cfg_gmem = 0;
    default_wave_cfgs(s, &cfg_smem, &cfg_gmem);
    // sample chunks: enough (pixel, chunk) items to keep every resident path slot busy many times over
    int lanes = 0;
    if (plan->use_smem && placement == 0) {
        // automatic placement: staging the scene must not cost occupancy (a mid-size blob can push the block past half of the
        // SM's shared memory: one block per SM instead of two) -- fall back to the global-memory path in that case
        int with_smem = 0, without = 0;
        cudaError_t es = rtd::render_resident_lanes(plan->variant, cfg_smem, true, plan->stats, s->L.general != 0, s->L.quant ? 2 : (s->L.flat_scan ? 3 : 0), s->L.total_bytes, s->stack_entries, s->sms, &with_smem);
        if (es != cudaSuccess) { cudaGetLastError(); with_smem = 0; }
        CUDA_TRY(rtd::render_resident_lanes(plan->variant, cfg_gmem, false, plan->stats, s->L.general != 0, s->L.quant ? 2 : (s->L.flat_scan ? 3 : 0), s->L.total_bytes, s->stack_entries, s->sms, &without));
        if (with_smem < without * 9 / 10) { plan->use_smem = false; a.blob = s->blob_dev; }     // (352 x 2 keeps 8 % fewer lanes than 384 x 2: not a reason to leave shared memory)
    }
    plan->cfg = plan->use_smem ? cfg_smem : cfg_gmem;
    CUDA_TRY(rtd::render_resident_lanes(plan->variant, plan->cfg, plan->use_smem, plan->stats, s->L.general != 0, s->L.quant ? 2 : (s->L.flat_scan ? 3 : 0), s->L.total_bytes, s->stack_entries, s->sms, &lanes));
    const int n_samp = s1 - s0;
    // measured (B200): work items ~32x the resident path slots keep the end-of-frame tail short (512x512x1024 spp: +16 % over
    // 4x); frames with plenty of pixels still get up to 4 chunks of >= 128 samples (3840x2160x1024: +0.7 %)
    long long want = ((long long)env_int("RT_CHUNK_FACTOR", 32) * lanes + a.n_pix_items - 1) / a.n_pix_items;
    want = std::max<long long>(want, std::min<long long>(4, n_samp / 128));
    if (want < 1) want = 1;
    if (want > n_samp) want = n_samp;
    const size_t n_pix = (size_t)h.width * (size_t)h.height;
    const long long max_layers = std::max<long long>(1, (long long)((1ull << 30) / (n_pix * sizeof(float4))));
    if (want > max_layers) want = max_layers;
    // chunk schedule: `want` equal chunks, then a geometric tail (1/16, 1/32, 1/64, 1/64 of the samples) when the frame can afford the
    // extra layers -- measured: one 128-sample item is 5.5 ms of a lane's time, which is what an 8-GPU frame (403 ms) used to lose
    // at its end; with tail chunks of a few samples the last items take a fraction of a millisecond
    int n_tail = 0, tail[4] = {0, 0, 0, 0};
    if (env_int("RT_TAIL_CHUNKS", 1) && n_samp >= 64 && want + 4 <= std::min<long long>(max_layers, RT_MAX_CHUNKS)) {
        tail[0] = n_samp / 16; tail[1] = n_samp / 32; tail[2] = n_samp / 64; tail[3] = n_samp / 64;
        n_tail = 4;
    }
    if (want > RT_MAX_CHUNKS - n_tail) want = RT_MAX_CHUNKS - n_tail;
    const int main_samples = n_samp - (tail[0] + tail[1] + tail[2] + tail[3]);
    if (want > main_samples) want = std::max(1, main_samples);
    std::vector<int32_t> cb(1, 0);
    int pos = 0;
    for (long long k = 0; k < want; ++k) {                     // near-equal split of the main part
        pos = (int)((long long)main_samples * (k + 1) / want);
        if (pos > cb.back()) cb.push_back(pos);
    }
    for (int k = 0; k < n_tail; ++k) if (tail[k] > 0) { pos += tail[k]; cb.push_back(pos); }
    a.n_chunks = (int)cb.size() - 1;
    if (!s->chunk_table) CUDA_TRY(cudaMalloc((void**)&s->chunk_table, (RT_MAX_CHUNKS + 1) * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpyAsync(s->chunk_table, cb.data(), cb.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));   // pageable source: staged before the call returns
    a.chunk_begin = s->chunk_table;
    const uint64_t total = (uint64_t)a.n_pix_items * (uint64_t)a.n_chunks;
    if (total >= 0xffffffffull) return fail(RT_ERR_LIMIT, "too many work items");
    a.total_items = (uint32_t)total;

    int rc = ensure(&s->layers, &s->layers_cap, n_pix * (size_t)a.n_chunks);
    if (rc != RT_OK) return rc;
    a.layers = s->layers; a.work_counter = s->work_counter; a.stats = s->stats_dev;
    if (a.shard_count > 1) CUDA_TRY(cudaMemsetAsync(s->layers, 0, n_pix * (size_t)a.n_chunks * sizeof(float4), stream));   // tiles of other shards stay zero
    CUDA_TRY(cudaMemsetAsync(s->work_counter, 0, sizeof(unsigned int), stream));
    if (plan->stats) CUDA_TRY(cudaMemsetAsync(s->stats_dev, 0, sizeof(unsigned long long) * rtd::RT_N_STATS, stream));
    CUDA_TRY(rtd::launch_render(a, plan->variant, plan->cfg, plan->use_smem, plan->stats, s->sms, stream, ki));
    return RT_OK;
}

int collect_stats(RtScene* s, const RenderPlan& plan, const rtd::KernelInfo& ki, RtStats* st, int launches, cudaStream_t stream) {
    if (!st) return RT_OK;
    std::memset(st, 0, sizeof(*st));
    st->kernel_launches = (uint64_t)launches;
    st->kernel = plan.variant; st->block_threads = ki.block; st->blocks_per_sm = ki.blocks_per_sm; st->grid_blocks = ki.grid;
    st->regs_per_thread = ki.regs; st->smem_bytes_per_block = ki.smem_bytes; st->scene_in_shared_memory = plan.use_smem ? 1 : 0;
    st->n_chunks = plan.args.n_chunks;
    if (plan.stats) {
        unsigned long long v[rtd::RT_N_STATS];
        CUDA_TRY(cudaMemcpyAsync(v, s->stats_dev, sizeof(v), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        st->samples = v[rtd::RT_STAT_SAMPLES]; st->segments = v[rtd::RT_STAT_SEGMENTS]; st->vertices = v[rtd::RT_STAT_VERTICES];
        st->attempts = v[rtd::RT_STAT_ATTEMPTS]; st->node_tests = v[rtd::RT_STAT_NODE_TESTS]; st->tri_tests = v[rtd::RT_STAT_TRI_TESTS];
        st->light_tri_tests = v[rtd::RT_STAT_LIGHT_TRI_TESTS]; st->attempt_cap_hits = v[rtd::RT_STAT_CAP_HITS]; st->nonfinite_samples = v[rtd::RT_STAT_NONFINITE];
    } else {
        uint64_t pixels = (uint64_t)s->host.width * (uint64_t)s->host.height;
        if (plan.args.shard_count > 1) {                    // pixels of the 8x4 tiles t with t % count == index
            pixels = 0;
            const uint32_t tx_n = plan.args.tiles_x, ty_n = (uint32_t)((s->host.height + 3) / 4);
            for (uint32_t t = plan.args.shard_index; t < tx_n * ty_n; t += plan.args.shard_count) {
                const uint32_t ty = t / tx_n, tx = t - ty * tx_n;
                pixels += (uint64_t)std::min(8, s->host.width - (int)tx * 8) * (uint64_t)std::min(4, s->host.height - (int)ty * 4);
            }
        }
        st->samples = pixels * (uint64_t)(plan.args.s_end - plan.args.s_begin);
    }
    return RT_OK;
}

}  // namespace

extern "C" {

const char* rt_last_error(void) { return g_last_error.c_str(); }
int rt_api_version(void) { return RT_API_VERSION; }
int rt_device_count(int32_t* count) {
    if (!count) return fail(RT_ERR_INVALID, "count is null");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; return fail(RT_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *count = c;
    return RT_OK;
}

int rt_scene_load_gltf(const char* path, int32_t width, int32_t height, int32_t samples, int32_t device, RtScene** out) {
    if (!path || !out) return fail(RT_ERR_INVALID, "null argument");
    *out = nullptr;
    RtScene* s = new RtScene();
    s->device = device;
    rtb::LoadError err;
    if (!rtb::load_gltf_scene(path, width, height, samples, &s->host, &err)) { delete s; return fail(err.code, err.message); }
    return finish_create(s, out);
}

int rt_scene_load_text(const char* path, int32_t width, int32_t height, int32_t samples, int32_t device, RtScene** out) {
    if (!path || !out) return fail(RT_ERR_INVALID, "null argument");
    *out = nullptr;
    RtScene* s = new RtScene();
    s->device = device;
    rtb::LoadError err;
    if (!rtb::load_text_scene(path, width, height, samples, &s->host, &err)) { delete s; return fail(err.code, err.message); }
    return finish_create(s, out);
}

int rt_scene_load(const char* path, int32_t width, int32_t height, int32_t samples, int32_t device, RtScene** out) {
    if (!path || !out) return fail(RT_ERR_INVALID, "null argument");
    const size_t len = std::strlen(path);
    if (len >= 4 && std::strcmp(path + len - 4, ".txt") == 0) return rt_scene_load_text(path, width, height, samples, device, out);
    return rt_scene_load_gltf(path, width, height, samples, device, out);
}

int rt_scene_create2(const RtSceneDesc2* d2, int32_t device, RtScene** out) {
    if (!d2 || !out) return fail(RT_ERR_INVALID, "null argument");
    const bool ext = d2->shape_kind || d2->position || d2->rotation || d2->ior || d2->material_kind;
    if (!ext) return rt_scene_create(&d2->base, device, out);
    const RtSceneDesc* d = &d2->base;
    *out = nullptr;
    if (d->n_tris < 0) return fail(RT_ERR_INVALID, "n_tris < 0");
    if (d->n_tris > 0 && (!d->tri_v || !d->tri_n || !d->tri_material || !d->tri_emission)) return fail(RT_ERR_INVALID, "null primitive array");
    const size_t n = (size_t)d->n_tris;
    for (size_t i = 0; i < n; ++i) {
        if (d2->shape_kind && (d2->shape_kind[i] < RT_SHAPE_TRIANGLE || d2->shape_kind[i] > RT_SHAPE_PLANE)) return fail(RT_ERR_INVALID, "shape_kind out of range");
        if (d2->material_kind && (d2->material_kind[i] < RT_MATERIAL_PBR || d2->material_kind[i] > RT_MATERIAL_DIELECTRIC)) return fail(RT_ERR_INVALID, "material_kind out of range");
        if (d2->rotation) {
            const double* q = d2->rotation + 4 * i;
            const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
            if (!(std::fabs(n2 - 1.0) <= 1e-6)) return fail(RT_ERR_INVALID, "rotation must be a unit quaternion (i, j, k, w)");
        }
        if (d2->ior && d2->material_kind && d2->material_kind[i] == RT_MATERIAL_DIELECTRIC && !(d2->ior[i] > 0.0)) return fail(RT_ERR_INVALID, "ior must be positive");
    }
    RtScene* s = nullptr;
    RtScene* tmp = nullptr;
    {   // the base fields go through the same copy as rt_scene_create (host-only), then the extension arrays are attached
        const int rc = rt_scene_create(d, -1, &tmp);
        if (rc != RT_OK) return rc;
        s = new RtScene();
        s->host = tmp->host;
        rt_scene_destroy(tmp);
    }
    s->device = device;
    rtb::HostScene& h = s->host;
    h.kind.assign(n, RT_SHAPE_TRIANGLE); h.position.assign(n * 3, 0.0); h.rotation.assign(n * 4, 0.0); h.ior.assign(n, 1.0); h.mat_kind.assign(n, RT_MATERIAL_PBR);
    for (size_t i = 0; i < n; ++i) h.rotation[i * 4 + 3] = 1.0;
    if (d2->shape_kind) h.kind.assign(d2->shape_kind, d2->shape_kind + n);
    if (d2->position) h.position.assign(d2->position, d2->position + n * 3);
    if (d2->rotation) h.rotation.assign(d2->rotation, d2->rotation + n * 4);
    if (d2->ior) h.ior.assign(d2->ior, d2->ior + n);
    if (d2->material_kind) h.mat_kind.assign(d2->material_kind, d2->material_kind + n);
    return finish_create(s, out);
}

int rt_scene_get_desc2(const RtScene* s, RtSceneDesc2* d) {
    if (!s || !d) return fail(RT_ERR_INVALID, "null argument");
    std::memset(d, 0, sizeof(*d));
    const int rc = rt_scene_get_desc(s, &d->base);
    if (rc != RT_OK) return rc;
    const rtb::HostScene& h = s->host;
    if (h.general()) {
        d->shape_kind = h.kind.data(); d->position = h.position.data(); d->rotation = h.rotation.data(); d->ior = h.ior.data(); d->material_kind = h.mat_kind.data();
    }
    return RT_OK;
}

int rt_scene_create(const RtSceneDesc* d, int32_t device, RtScene** out) {
    if (!d || !out) return fail(RT_ERR_INVALID, "null argument");
    *out = nullptr;
    if (d->n_tris < 0) return fail(RT_ERR_INVALID, "n_tris < 0");
    if (d->n_tris > 0 && (!d->tri_v || !d->tri_n || !d->tri_material || !d->tri_emission)) return fail(RT_ERR_INVALID, "null triangle array");
    RtScene* s = new RtScene();
    s->device = device;
    rtb::HostScene& h = s->host;
    h.width = d->width; h.height = d->height; h.samples = d->samples; h.ray_depth = d->ray_depth;
    for (int a = 0; a < 3; ++a) {
        h.bg_color[a] = d->bg_color[a]; h.camera_position[a] = d->camera_position[a]; h.camera_forward[a] = d->camera_forward[a];
        h.camera_right[a] = d->camera_right[a]; h.camera_up[a] = d->camera_up[a];
    }
    h.camera_fov_x = d->camera_fov_x; h.camera_fov_y = d->camera_fov_y;
    const size_t n = (size_t)d->n_tris;
    h.tri_v.assign(d->tri_v, d->tri_v + n * 9); h.tri_n.assign(d->tri_n, d->tri_n + n * 9);
    h.tri_material.assign(d->tri_material, d->tri_material + n * 5); h.tri_emission.assign(d->tri_emission, d->tri_emission + n * 3);
    return finish_create(s, out);
}

int rt_release_device_cache(void) {
    std::vector<CachedBlock> all;
    {
        std::lock_guard<std::mutex> lock(cache_mutex());
        all.swap(cache_blocks());
    }
    int prev = 0;
    const bool have_prev = cudaGetDevice(&prev) == cudaSuccess;
    for (const CachedBlock& b : all) { cudaSetDevice(b.device); cudaFree(b.p); }
    if (have_prev) cudaSetDevice(prev);
    return RT_OK;
}

void rt_scene_destroy(RtScene* s) {
    if (!s) return;
    if (s->stream || s->blob_dev) cudaSetDevice(s->device);
    if (s->stream || s->blob_dev) cudaDeviceSynchronize();    // the frame buffers go back to the cache: nothing (on any stream) may still use them
    cache_put(s->layers, s->layers_cap * sizeof(float4)); cache_put(s->accum, s->accum_cap * sizeof(float4));
    cache_put(s->rgb_dev, s->rgb_cap); cache_put(s->lin_dev, s->lin_cap * sizeof(float));
    cudaFree(s->blob_dev); cudaFree(s->tri_d_dev);
    cudaFree(s->work_counter); cudaFree(s->stats_dev); cudaFree(s->chunk_table);
    for (int i = 0; i < 5; ++i) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int rt_scene_get_desc(const RtScene* s, RtSceneDesc* d) {
    if (!s || !d) return fail(RT_ERR_INVALID, "null argument");
    const rtb::HostScene& h = s->host;
    std::memset(d, 0, sizeof(*d));
    d->width = h.width; d->height = h.height; d->samples = h.samples; d->ray_depth = h.ray_depth;
    for (int a = 0; a < 3; ++a) {
        d->bg_color[a] = h.bg_color[a]; d->camera_position[a] = h.camera_position[a]; d->camera_forward[a] = h.camera_forward[a];
        d->camera_right[a] = h.camera_right[a]; d->camera_up[a] = h.camera_up[a];
    }
    d->camera_fov_x = h.camera_fov_x; d->camera_fov_y = h.camera_fov_y;
    d->n_tris = h.n_tris();
    d->tri_v = h.tri_v.data(); d->tri_n = h.tri_n.data(); d->tri_material = h.tri_material.data(); d->tri_emission = h.tri_emission.data();
    return RT_OK;
}

int rt_scene_info(const RtScene* s, RtSceneInfo* o) {
    if (!s || !o) return fail(RT_ERR_INVALID, "null argument");
    std::memset(o, 0, sizeof(*o));
    o->n_tris = s->L.n_tris; o->n_lights = (int32_t)s->light_ids.size(); o->n_materials = s->n_mats;
    o->n_infinite = s->L.n_planes; o->general_primitives = s->L.general;
    o->n_nodes = s->bvh.n_nodes; o->n_leaves = s->bvh.n_leaves; o->bvh_depth = s->bvh.depth; o->max_leaf_size = s->bvh.max_leaf;
    o->bvh_validate_failures = s->validate_failures; o->scene_in_shared_memory = s->use_smem ? 1 : 0; o->device = s->device;
    o->device_bytes = (int64_t)s->blob_host.size();
    o->bvh_builder = s->bvh_builder; o->bvh_build_ms = s->bvh_build_ms;
    return RT_OK;
}

int rt_scene_set_frame(RtScene* s, int32_t width, int32_t height, int32_t samples) {
    if (!s) return fail(RT_ERR_INVALID, "null scene");
    if (width <= 0 || height <= 0 || samples < 0) return fail(RT_ERR_INVALID, "rt_scene_set_frame: width and height must be positive, samples >= 0");
    if (width > 65535 || height > 65535) return fail(RT_ERR_LIMIT, "frame dimensions above 65535 (pixel coordinates are packed in 16 bits)");
    s->host.width = width; s->host.height = height; s->host.samples = samples;
    return RT_OK;
}

int rt_scene_get_bvh(const RtScene* s, float* nodes, int32_t* tri_order) {
    if (!s) return fail(RT_ERR_INVALID, "null scene");
    if (nodes)
        for (int n = 0; n < s->bvh.n_nodes; ++n) {
            const float* A = &s->bvh.box_a[(size_t)n * 4]; const float* B = &s->bvh.box_b[(size_t)n * 4]; const float* C = &s->bvh.box_c[(size_t)n * 4];
            float* o = nodes + (size_t)n * 14;
            o[0] = A[0]; o[1] = A[2]; o[2] = C[0]; o[3] = A[1]; o[4] = A[3]; o[5] = C[1];
            o[6] = B[0]; o[7] = B[2]; o[8] = C[2]; o[9] = B[1]; o[10] = B[3]; o[11] = C[3];
            o[12] = (float)s->bvh.child[(size_t)n * 2]; o[13] = (float)s->bvh.child[(size_t)n * 2 + 1];
        }
    if (tri_order) for (size_t k = 0; k < s->bvh.tri_order.size(); ++k) tri_order[k] = s->bvh.tri_order[k];
    return RT_OK;
}

int rt_scene_get_quantised_bvh(const RtScene* s, uint32_t* words, float* grid6, int32_t* present) {
    if (!s || !present) return fail(RT_ERR_INVALID, "null argument");
    *present = s->L.quant ? 1 : 0;
    if (!s->L.quant) return RT_OK;
    if (words) std::memcpy(words, s->blob_host.data() + s->L.qnodes, (size_t)s->bvh.n_nodes * 32u);
    if (grid6) for (int a = 0; a < 3; ++a) { grid6[a] = s->L.qorg[a]; grid6[3 + a] = s->L.qcell[a]; }
    return RT_OK;
}

// render_scene (rendering.rs:21-69)
int rt_render(RtScene* s, const RtRenderParams* p, uint8_t* rgb_out, RtStats* st) {
    if (!s || !rgb_out) return fail(RT_ERR_INVALID, "null argument");
    if (!s->blob_dev) return fail(RT_ERR_CUDA, "scene is host-only (created with device < 0): rendering needs a CUDA device, there is no CPU fallback");
    RenderPlan plan; rtd::KernelInfo ki;
    const size_t n_pix = (size_t)s->host.width * (size_t)s->host.height;
    cudaStream_t stream = s->stream;
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaEventRecord(s->ev[0], stream));
    CUDA_TRY(cudaEventRecord(s->ev[1], stream));
    int rc = render_to_layers(s, p, stream, &plan, &ki);
    if (rc != RT_OK) return rc;
    if ((rc = ensure(&s->accum, &s->accum_cap, n_pix)) != RT_OK) return rc;
    if ((rc = ensure(&s->rgb_dev, &s->rgb_cap, n_pix * 3)) != RT_OK) return rc;
    CUDA_TRY(rtd::launch_sum_layers(s->layers, plan.args.n_chunks, n_pix, s->accum, false, stream));
    CUDA_TRY(cudaEventRecord(s->ev[4], stream));
    CUDA_TRY(rtd::launch_resolve_u8(s->accum, n_pix, s->rgb_dev, stream));
    CUDA_TRY(cudaEventRecord(s->ev[2], stream));
    CUDA_TRY(cudaMemcpyAsync(rgb_out, s->rgb_dev, n_pix * 3, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaEventRecord(s->ev[3], stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if ((rc = collect_stats(s, plan, ki, st, 3, stream)) != RT_OK) return rc;
    if (st) {
        float k = 0, t = 0, r = 0, v = 0;
        cudaEventElapsedTime(&k, s->ev[1], s->ev[2]); cudaEventElapsedTime(&t, s->ev[0], s->ev[3]);
        cudaEventElapsedTime(&r, s->ev[0], s->ev[4]); cudaEventElapsedTime(&v, s->ev[4], s->ev[2]);
        st->kernel_ms = k; st->total_ms = t; st->render_ms = r; st->reduce_ms = 0.0; st->resolve_ms = v;
    }
    return RT_OK;
}

int rt_render_linear(RtScene* s, const RtRenderParams* p, float* out, RtStats* st) {
    if (!s || !out) return fail(RT_ERR_INVALID, "null argument");
    if (!s->blob_dev) return fail(RT_ERR_CUDA, "scene is host-only (created with device < 0): rendering needs a CUDA device, there is no CPU fallback");
    RenderPlan plan; rtd::KernelInfo ki;
    const size_t n_pix = (size_t)s->host.width * (size_t)s->host.height;
    cudaStream_t stream = s->stream;
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaEventRecord(s->ev[0], stream));
    int rc = render_to_layers(s, p, stream, &plan, &ki);
    if (rc != RT_OK) return rc;
    if ((rc = ensure(&s->accum, &s->accum_cap, n_pix)) != RT_OK) return rc;
    if ((rc = ensure(&s->lin_dev, &s->lin_cap, n_pix * 3)) != RT_OK) return rc;
    CUDA_TRY(rtd::launch_sum_layers(s->layers, plan.args.n_chunks, n_pix, s->accum, false, stream));
    CUDA_TRY(rtd::launch_resolve_linear(s->accum, n_pix, s->lin_dev, stream));
    CUDA_TRY(cudaEventRecord(s->ev[2], stream));
    CUDA_TRY(cudaMemcpyAsync(out, s->lin_dev, n_pix * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaEventRecord(s->ev[3], stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if ((rc = collect_stats(s, plan, ki, st, 3, stream)) != RT_OK) return rc;
    if (st) {
        float k = 0, t = 0;
        cudaEventElapsedTime(&k, s->ev[0], s->ev[2]); cudaEventElapsedTime(&t, s->ev[0], s->ev[3]);
        st->kernel_ms = k; st->total_ms = t;
    }
    return RT_OK;
}

int rt_render_accumulate_device(RtScene* s, const RtRenderParams* p, float* accum_dev, void* stream_v, RtStats* st) {
    if (!s || !accum_dev) return fail(RT_ERR_INVALID, "null argument");
    if (!s->blob_dev) return fail(RT_ERR_CUDA, "scene is host-only (created with device < 0): rendering needs a CUDA device, there is no CPU fallback");
    RenderPlan plan; rtd::KernelInfo ki;
    const size_t n_pix = (size_t)s->host.width * (size_t)s->host.height;
    cudaStream_t stream = stream_v ? (cudaStream_t)stream_v : s->stream;
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaEventRecord(s->ev[0], stream));
    int rc = render_to_layers(s, p, stream, &plan, &ki);
    if (rc != RT_OK) return rc;
    CUDA_TRY(rtd::launch_sum_layers(s->layers, plan.args.n_chunks, n_pix, reinterpret_cast<float4*>(accum_dev), true, stream));
    CUDA_TRY(cudaEventRecord(s->ev[2], stream));
    if (st) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        if ((rc = collect_stats(s, plan, ki, st, 2, stream)) != RT_OK) return rc;
        float k = 0;
        cudaEventElapsedTime(&k, s->ev[0], s->ev[2]);
        st->kernel_ms = k; st->total_ms = k;
    }
    return RT_OK;
}

// ---- single-process multi-GPU orchestration (north_star: main.rs device and multi-GPU orchestration) -------------------
// NCCL is reached through dlopen so that the library has no link-time dependency on a particular libnccl (a Python host
// usually carries its own); when it cannot be loaded the framebuffers are summed with peer copies instead.
namespace {
struct NcclApi {
    void* handle = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
// One lock serialises libnccl loading, communicator (re)creation and every multi-GPU frame of the process: the communicator
// cache is process-wide state and NCCL group calls on one communicator set must not interleave between host threads.  Distinct
// scenes used through the single-GPU entry points from distinct threads never touch it.
std::mutex& multi_gpu_mutex() { static std::mutex m; return m; }
NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* override_path = std::getenv("RT_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (api.handle) break;
    }
    if (!api.handle) return api;
    api.CommInitAll = (int (*)(void**, int, const int*))dlsym(api.handle, "ncclCommInitAll");
    api.CommDestroy = (int (*)(void*))dlsym(api.handle, "ncclCommDestroy");
    api.GroupStart = (int (*)())dlsym(api.handle, "ncclGroupStart");
    api.GroupEnd = (int (*)())dlsym(api.handle, "ncclGroupEnd");
    api.Reduce = (int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t))dlsym(api.handle, "ncclReduce");
    api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
    api.ok = api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Reduce;
    return api;
}
struct CommCache { std::vector<int> devs; std::vector<void*> comms; };
CommCache& comm_cache() { static CommCache c; return c; }
const int kNcclFloat32 = 7, kNcclSum = 0;     // nccl.h: ncclFloat32 = 7, ncclSum = 0 (stable since NCCL 2.0)
}  // namespace

namespace {
// Creates (or reuses) the NCCL communicators of this device list; false when NCCL is unavailable or switched off.
bool ensure_comms(const std::vector<int>& devs) {
    NcclApi& nc = nccl_api();
    if (!nc.ok || env_int("RT_NO_NCCL", 0)) return false;
    CommCache& cc = comm_cache();
    if (cc.devs == devs) return true;
    for (void* c : cc.comms) nc.CommDestroy(c);
    cc.comms.assign(devs.size(), nullptr); cc.devs.clear();
    if (nc.CommInitAll(cc.comms.data(), (int)devs.size(), devs.data()) != 0) { cc.comms.clear(); return false; }
    cc.devs = devs;
    return true;
}
}  // namespace

int rt_multi_init(RtScene* const* scenes, int32_t n) {
    if (!scenes || n < 1) return fail(RT_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lock(multi_gpu_mutex());
    std::vector<int> devs;
    for (int g = 0; g < n; ++g) { if (!scenes[g] || !scenes[g]->blob_dev) return fail(RT_ERR_CUDA, "rt_multi_init: every scene must live on a CUDA device"); devs.push_back(scenes[g]->device); }
    // frame buffers of the scenes' current frame size: the first frame must not pay ~1 GB of cudaMalloc per device inside the timing
    // window of main.rs:54-58 (it did: the one-frame CLI ran 5 % behind the warmed-up bench loop on 8 GPUs)
    for (int g = 0; g < n; ++g) {
        RtScene* s = scenes[g];
        if (s->host.width <= 0 || s->host.height <= 0) continue;
        CUDA_TRY(cudaSetDevice(s->device));
        const size_t n_pix = (size_t)s->host.width * (size_t)s->host.height;
        const size_t layers = std::min<size_t>(8, std::max<size_t>(1, (size_t)((1ull << 30) / (n_pix * sizeof(float4)))));
        int rc = ensure(&s->layers, &s->layers_cap, n_pix * layers);
        if (rc == RT_OK) rc = ensure(&s->accum, &s->accum_cap, n_pix);
        if (rc == RT_OK && g == 0) rc = ensure(&s->rgb_dev, &s->rgb_cap, n_pix * 3);
        if (rc != RT_OK) return rc;
    }
    if (n > 1 && ensure_comms(devs)) {                      // first collective = connection set-up: do it now, on a few floats
        NcclApi& nc = nccl_api();
        CommCache& cc = comm_cache();
        for (int g = 0; g < n; ++g) {
            CUDA_TRY(cudaSetDevice(devs[(size_t)g]));
            const int rc = ensure(&scenes[g]->accum, &scenes[g]->accum_cap, 256);
            if (rc != RT_OK) return rc;
            CUDA_TRY(cudaMemsetAsync(scenes[g]->accum, 0, 256 * sizeof(float4), scenes[g]->stream));
        }
        int rc = nc.GroupStart();
        for (int g = 0; g < n && rc == 0; ++g) rc = nc.Reduce(scenes[g]->accum, scenes[0]->accum, 1024, kNcclFloat32, kNcclSum, 0, cc.comms[(size_t)g], scenes[g]->stream);
        const int rc2 = nc.GroupEnd();
        if (rc != 0 || rc2 != 0) return fail(RT_ERR_CUDA, "rt_multi_init: ncclReduce warm-up failed");
        for (int g = 0; g < n; ++g) { CUDA_TRY(cudaSetDevice(devs[(size_t)g])); CUDA_TRY(cudaStreamSynchronize(scenes[g]->stream)); }
    } else if (n > 1) {                                     // no NCCL: make the peer copies direct
        for (int g = 1; g < n; ++g) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[0], devs[(size_t)g]);
            if (can) { cudaSetDevice(devs[0]); cudaError_t e = cudaDeviceEnablePeerAccess(devs[(size_t)g], 0); if (e != cudaSuccess) cudaGetLastError(); }
        }
    }
    return RT_OK;
}

namespace {
int render_multi_locked(RtScene* const* scenes, int32_t n, const RtRenderParams* p, uint8_t* rgb_out, RtStats* st);
}
int rt_render_multi(RtScene* const* scenes, int32_t n, const RtRenderParams* p, uint8_t* rgb_out, RtStats* st) {
    if (!scenes || n < 1 || !rgb_out) return fail(RT_ERR_INVALID, "bad argument");
    if (n == 1) return rt_render(scenes[0], p, rgb_out, st);
    std::lock_guard<std::mutex> lock(multi_gpu_mutex());
    const int rc = render_multi_locked(scenes, n, p, rgb_out, st);
    if (rc != RT_OK) {
        // a failure part-way leaves kernels / copies in flight on the devices launched so far: drain every stream before the
        // caller may free or reuse its buffers (rgb_out is the target of an asynchronous copy), keeping the first error message
        const std::string first = g_last_error;
        for (int g = 0; g < n; ++g)
            if (scenes[g] && scenes[g]->stream) { cudaSetDevice(scenes[g]->device); cudaStreamSynchronize(scenes[g]->stream); }
        cudaGetLastError();
        g_last_error = first;
    }
    return rc;
}
namespace {
int render_multi_locked(RtScene* const* scenes, int32_t n, const RtRenderParams* p, uint8_t* rgb_out, RtStats* st) {
    RtScene* s0 = scenes[0];
    if (!s0) return fail(RT_ERR_INVALID, "null scene");
    const int W = s0->host.width, H = s0->host.height, S = s0->host.samples;
    const size_t n_pix = (size_t)W * (size_t)H;
    std::vector<int> devs((size_t)n);
    for (int g = 0; g < n; ++g) {
        RtScene* s = scenes[g];
        if (!s || !s->blob_dev) return fail(RT_ERR_CUDA, "rt_render_multi: every scene must live on a CUDA device");
        if (s->host.width != W || s->host.height != H || s->host.samples != S || s->host.n_tris() != s0->host.n_tris())
            return fail(RT_ERR_INVALID, "rt_render_multi: the scenes must describe the same frame");
        devs[(size_t)g] = s->device;
        for (int k = 0; k < g; ++k) if (devs[(size_t)k] == s->device) return fail(RT_ERR_INVALID, "rt_render_multi: one scene per device");
    }
    RtRenderParams base;
    std::memset(&base, 0, sizeof(base));
    if (p) base = *p;
    int s_lo = base.sample_begin, s_hi = base.sample_end;
    if (s_lo == 0 && s_hi == 0) s_hi = S;
    if (s_lo < 0 || s_hi > S || s_lo >= s_hi) return fail(RT_ERR_INVALID, "sample range must satisfy 0 <= begin < end <= samples");
    // 1. every device renders its sample shard into its own accumulator (launches are asynchronous: the GPUs run together)
    std::vector<RenderPlan> plans((size_t)n);
    std::vector<rtd::KernelInfo> kis((size_t)n);
    std::vector<char> active((size_t)n, 0);
    CUDA_TRY(cudaSetDevice(s0->device));
    CUDA_TRY(cudaEventRecord(s0->ev[0], s0->stream));
    // One host thread per device for the launch sequence (set device, occupancy queries, chunk table upload, memsets, kernel launch:
    // ~0.5 ms of host time each): launched from a single thread the 8th GPU started 4 ms after the first (measured with the CLI).
    auto launch_one = [&](int g, std::string* err) -> int {
        RtScene* s = scenes[g];
        auto done = [&](int rc) { if (rc != RT_OK) *err = g_last_error; return rc; };      // g_last_error is thread-local: hand it back
        cudaError_t ce = cudaSetDevice(s->device);
        if (ce != cudaSuccess) return done(fail(RT_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce)));
        int rc = ensure(&s->accum, &s->accum_cap, n_pix);
        if (rc != RT_OK) return done(rc);
        const long long span = s_hi - s_lo;
        RtRenderParams q = base;
        if (span >= n) {
            q.sample_begin = s_lo + (int)(span * g / n); q.sample_end = s_lo + (int)(span * (g + 1) / n);   // sample sharding (multigpu.shard_range)
        } else {
            q.sample_begin = s_lo; q.sample_end = s_hi;                                                    // fewer samples than GPUs: interleaved
            q.tile_shard_index = g; q.tile_shard_count = n;                                                // tiles, every sample of each
        }
        if (q.sample_end > q.sample_begin) {
            rc = render_to_layers(s, &q, s->stream, &plans[(size_t)g], &kis[(size_t)g]);
            if (rc != RT_OK) return done(rc);
            ce = rtd::launch_sum_layers(s->layers, plans[(size_t)g].args.n_chunks, n_pix, s->accum, false, s->stream);
            if (ce != cudaSuccess) return done(fail(RT_ERR_CUDA, std::string("sum_layers: ") + cudaGetErrorString(ce)));
            active[(size_t)g] = 1;
        } else {
            ce = cudaMemsetAsync(s->accum, 0, n_pix * sizeof(float4), s->stream);
            if (ce != cudaSuccess) return done(fail(RT_ERR_CUDA, std::string("cudaMemsetAsync: ") + cudaGetErrorString(ce)));
        }
        if (g == 0) { ce = cudaEventRecord(s->ev[1], s->stream); if (ce != cudaSuccess) return done(fail(RT_ERR_CUDA, std::string("cudaEventRecord: ") + cudaGetErrorString(ce))); }
        return RT_OK;
    };
    {
        std::vector<int> rcs((size_t)n, RT_OK);
        std::vector<std::string> errs((size_t)n);
        std::vector<std::thread> workers;
        for (int g = 1; g < n; ++g) workers.emplace_back([&, g]() { rcs[(size_t)g] = launch_one(g, &errs[(size_t)g]); });
        rcs[0] = launch_one(0, &errs[0]);
        for (std::thread& w : workers) w.join();
        for (int g = 0; g < n; ++g) if (rcs[(size_t)g] != RT_OK) return fail(rcs[(size_t)g], errs[(size_t)g]);
    }
    // 2. ONE reduce(sum) of the W*H*4 accumulators to device 0 (NVLink), the only communication of the frame
    NcclApi& nc = nccl_api();
    bool reduced = false;
    if (ensure_comms(devs)) {
        CommCache& cc = comm_cache();
        int rc = nc.GroupStart();
        for (int g = 0; g < n && rc == 0; ++g)
            rc = nc.Reduce(scenes[g]->accum, s0->accum, n_pix * 4, kNcclFloat32, kNcclSum, 0, cc.comms[(size_t)g], scenes[g]->stream);
        const int rc2 = nc.GroupEnd();
        if (rc != 0 || rc2 != 0) return fail(RT_ERR_CUDA, std::string("ncclReduce: ") + (nc.GetErrorString ? nc.GetErrorString(rc ? rc : rc2) : "error"));
        reduced = true;
    }
    if (!reduced) {                                   // peer copies into device 0's (now free) layer buffer + sum
        for (int g = 1; g < n; ++g) {
            RtScene* s = scenes[g];
            CUDA_TRY(cudaSetDevice(s->device));
            CUDA_TRY(cudaEventRecord(s->ev[3], s->stream));
            CUDA_TRY(cudaSetDevice(s0->device));
            CUDA_TRY(cudaStreamWaitEvent(s0->stream, s->ev[3], 0));
            int rc = ensure(&s0->layers, &s0->layers_cap, n_pix);
            if (rc != RT_OK) return rc;
            CUDA_TRY(cudaMemcpyPeerAsync(s0->layers, s0->device, s->accum, s->device, n_pix * sizeof(float4), s0->stream));
            CUDA_TRY(rtd::launch_sum_layers(s0->layers, 1, n_pix, s0->accum, true, s0->stream));
        }
    }
    // 3. device 0 resolves (color_to_pixel) and returns the bytes
    CUDA_TRY(cudaSetDevice(s0->device));
    int rc = ensure(&s0->rgb_dev, &s0->rgb_cap, n_pix * 3);
    if (rc != RT_OK) return rc;
    CUDA_TRY(cudaEventRecord(s0->ev[4], s0->stream));
    CUDA_TRY(rtd::launch_resolve_u8(s0->accum, n_pix, s0->rgb_dev, s0->stream));
    CUDA_TRY(cudaEventRecord(s0->ev[2], s0->stream));
    CUDA_TRY(cudaMemcpyAsync(rgb_out, s0->rgb_dev, n_pix * 3, cudaMemcpyDeviceToHost, s0->stream));
    CUDA_TRY(cudaEventRecord(s0->ev[3], s0->stream));
    for (int g = n - 1; g >= 0; --g) { CUDA_TRY(cudaSetDevice(scenes[g]->device)); CUDA_TRY(cudaStreamSynchronize(scenes[g]->stream)); }
    if (st) {
        RtStats total; std::memset(&total, 0, sizeof(total));
        for (int g = 0; g < n; ++g) {
            if (!active[(size_t)g]) continue;
            RtStats one;
            CUDA_TRY(cudaSetDevice(scenes[g]->device));
            if ((rc = collect_stats(scenes[g], plans[(size_t)g], kis[(size_t)g], &one, 2, scenes[g]->stream)) != RT_OK) return rc;
            const uint64_t* a = &one.samples; uint64_t* b = &total.samples;
            for (int k = 0; k < 10; ++k) b[k] += a[k];                       // the ten uint64 counters
            total.kernel = one.kernel; total.block_threads = one.block_threads; total.blocks_per_sm = one.blocks_per_sm; total.grid_blocks = one.grid_blocks;
            total.regs_per_thread = one.regs_per_thread; total.smem_bytes_per_block = one.smem_bytes_per_block; total.scene_in_shared_memory = one.scene_in_shared_memory;
            total.n_chunks = one.n_chunks;
        }
        total.kernel_launches += 2;                                           // resolve + (reduce or sum)
        float k = 0, t = 0;
        CUDA_TRY(cudaSetDevice(s0->device));
        cudaEventElapsedTime(&k, s0->ev[0], s0->ev[2]); cudaEventElapsedTime(&t, s0->ev[0], s0->ev[3]);
        total.kernel_ms = k; total.total_ms = t;
        float r0 = 0, r1 = 0, r2 = 0;                                         // phases on device 0's clock: its render, the reduce, the resolve
        cudaEventElapsedTime(&r0, s0->ev[0], s0->ev[1]); cudaEventElapsedTime(&r1, s0->ev[1], s0->ev[4]); cudaEventElapsedTime(&r2, s0->ev[4], s0->ev[2]);
        total.render_ms = r0; total.reduce_ms = r1; total.resolve_ms = r2;
        *st = total;
    }
    return RT_OK;
}
}  // namespace

int rt_resolve_device(const float* accum_dev, int32_t width, int32_t height, uint8_t* rgb_dev, void* stream_v) {
    if (!accum_dev || !rgb_dev || width <= 0 || height <= 0) return fail(RT_ERR_INVALID, "bad argument");
    CUDA_TRY(rtd::launch_resolve_u8(reinterpret_cast<const float4*>(accum_dev), (size_t)width * (size_t)height, rgb_dev, (cudaStream_t)stream_v));
    return RT_OK;
}

namespace {
int trace_rays_impl(RtScene* s, const double* rays, int64_t n, int32_t precision, int32_t* tri_id, double* t, double* hits);
}
int rt_trace_hits(RtScene* s, const double* rays, int64_t n, double* out) {
    if (!s || !rays || !out || n < 0) return fail(RT_ERR_INVALID, "bad argument");
    return trace_rays_impl(s, rays, n, 32, nullptr, nullptr, out);
}
int rt_trace_primary(RtScene* s, const double* rays, int64_t n, int32_t precision, int32_t* tri_id, double* t) {
    if (!s || !rays || !tri_id || !t || n < 0) return fail(RT_ERR_INVALID, "bad argument");
    if (precision != 32 && precision != 64) return fail(RT_ERR_INVALID, "precision must be 32 or 64");
    return trace_rays_impl(s, rays, n, precision, tri_id, t, nullptr);
}
namespace {
int trace_rays_impl(RtScene* s, const double* rays, int64_t n, int32_t precision, int32_t* tri_id, double* t, double* hits) {
    if (n == 0) return RT_OK;
    if (!s->blob_dev) return fail(RT_ERR_CUDA, "scene is host-only: no CUDA device bound, there is no CPU fallback");
    if (precision == 64 && s->L.general) return fail(RT_ERR_INVALID, "precision 64 exists for triangle scenes only (f64 triangle test); general-primitive scenes trace in FP32");
    CUDA_TRY(cudaSetDevice(s->device));
    if (precision == 64 && !s->tri_d_dev) {
        CUDA_TRY(cudaMalloc((void**)&s->tri_d_dev, s->tri_d_host.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpy(s->tri_d_dev, s->tri_d_host.data(), s->tri_d_host.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    double *rays_d = nullptr, *t_d = nullptr, *hits_d = nullptr; int32_t* id_d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&rays_d, (size_t)n * 6 * sizeof(double)));
    cudaError_t e = cudaSuccess;
    if (t) e = cudaMalloc((void**)&t_d, (size_t)n * sizeof(double));
    if (e == cudaSuccess && tri_id) e = cudaMalloc((void**)&id_d, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess && hits) e = cudaMalloc((void**)&hits_d, (size_t)n * 9 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(rays_d, rays, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) e = rtd::launch_trace_rays(s->blob_dev, s->L, s->tri_d_dev, s->stack_entries, rays_d, n, precision == 64, id_d, t_d, hits_d, s->stream);
    if (e == cudaSuccess && tri_id) e = cudaMemcpyAsync(tri_id, id_d, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess && t) e = cudaMemcpyAsync(t, t_d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess && hits) e = cudaMemcpyAsync(hits, hits_d, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(rays_d); cudaFree(t_d); cudaFree(id_d); cudaFree(hits_d);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("rt_trace_primary: ") + cudaGetErrorString(e));
    return RT_OK;
}
}  // namespace

int rt_primary_rays(RtScene* s, const int32_t* xy, const double* xi, int64_t n, double* rays_out) {
    if (!s || !xy || !xi || !rays_out || n < 0) return fail(RT_ERR_INVALID, "bad argument");
    if (n == 0) return RT_OK;
    if (!s->blob_dev) return fail(RT_ERR_CUDA, "scene is host-only: no CUDA device bound, there is no CPU fallback");
    CUDA_TRY(cudaSetDevice(s->device));
    rtd::Camera cam; fill_camera(s->host, &cam);
    int32_t* xy_d = nullptr; double *xi_d = nullptr, *r_d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&xy_d, (size_t)n * 2 * sizeof(int32_t)));
    cudaError_t e = cudaMalloc((void**)&xi_d, (size_t)n * 2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&r_d, (size_t)n * 6 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(xy_d, xy, (size_t)n * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(xi_d, xi, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) e = rtd::launch_primary_rays(cam, s->host.width, s->host.height, xy_d, xi_d, n, r_d, s->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rays_out, r_d, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(xy_d); cudaFree(xi_d); cudaFree(r_d);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("rt_primary_rays: ") + cudaGetErrorString(e));
    return RT_OK;
}

int rt_eval(RtScene* s, int32_t fn, const float* in, int64_t n, float* out) {
    const int win = rtd::eval_in_width(fn), wout = rtd::eval_out_width(fn);
    if (win == 0 || !in || !out || n < 0) return fail(RT_ERR_INVALID, "bad argument");
    const bool needs_scene = fn == RT_FN_PDF_LIGHT || fn == RT_FN_PDF_MIX || fn == RT_FN_SAMPLE_LIGHT;
    if (needs_scene && !s) return fail(RT_ERR_INVALID, "this function needs a scene");
    if (s && !s->blob_dev) return fail(RT_ERR_CUDA, "scene is host-only: no CUDA device bound, there is no CPU fallback");
    if (n == 0) return RT_OK;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return fail(RT_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (s) CUDA_TRY(cudaSetDevice(s->device));
    rtd::SceneLayout L;
    std::memset(&L, 0, sizeof(L));
    if (s) L = s->L;
    float *in_d = nullptr, *out_d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&in_d, (size_t)n * (size_t)win * sizeof(float)));
    cudaError_t e = cudaMalloc((void**)&out_d, (size_t)n * (size_t)wout * sizeof(float));
    cudaStream_t stream = s ? s->stream : (cudaStream_t)0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(in_d, in, (size_t)n * (size_t)win * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = rtd::launch_eval(s ? s->blob_dev : nullptr, L, s ? s->stack_entries : 16u, fn, in_d, n, out_d, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, out_d, (size_t)n * (size_t)wout * sizeof(float), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(in_d); cudaFree(out_d);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("rt_eval: ") + cudaGetErrorString(e));
    return RT_OK;
}

// dump_rendered_to_ppm (main.rs:88-95)
int rt_write_ppm(const char* path, int32_t width, int32_t height, const uint8_t* rgb, int32_t append) {
    if (!path || !rgb || width <= 0 || height <= 0) return fail(RT_ERR_INVALID, "bad argument");
    FILE* f = std::fopen(path, append ? "ab" : "wb");
    if (!f) return fail(RT_ERR_IO, std::string("cannot open ") + path);
    std::fprintf(f, "P6\n%d %d\n255\n", width, height);
    const size_t n = (size_t)width * (size_t)height * 3;
    const bool ok = std::fwrite(rgb, 1, n, f) == n;
    std::fclose(f);
    return ok ? RT_OK : fail(RT_ERR_IO, std::string("short write to ") + path);
}

// dump_rendered_to_png (main.rs:75-86): 8-bit RGB PNG with the same pixels as the PPM.  The reference goes through the
// `image` crate; this writer is self-contained: one IDAT chunk holding a zlib stream of STORED deflate blocks (filter
// type 0 on every row), CRC-32 per chunk, Adler-32 over the raw scanlines.  Any PNG reader decodes it to the same bytes.
namespace {
uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}
void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x); }
bool write_chunk(FILE* f, const char type[4], const std::vector<uint8_t>& data) {
    std::vector<uint8_t> head; put_be32(head, (uint32_t)data.size());
    uint32_t crc = crc32_update(0xffffffffu, (const uint8_t*)type, 4);
    if (!data.empty()) crc = crc32_update(crc, data.data(), data.size());
    std::vector<uint8_t> tail; put_be32(tail, crc ^ 0xffffffffu);
    return std::fwrite(head.data(), 1, 4, f) == 4 && std::fwrite(type, 1, 4, f) == 4 &&
           (data.empty() || std::fwrite(data.data(), 1, data.size(), f) == data.size()) && std::fwrite(tail.data(), 1, 4, f) == 4;
}
}  // namespace

int rt_write_png(const char* path, int32_t width, int32_t height, const uint8_t* rgb) {
    if (!path || !rgb || width <= 0 || height <= 0) return fail(RT_ERR_INVALID, "bad argument");
    const size_t row = (size_t)width * 3, raw_n = (row + 1) * (size_t)height;
    if (raw_n > 0x7fffffffull) return fail(RT_ERR_LIMIT, "image too large for a single-IDAT PNG");
    std::vector<uint8_t> raw(raw_n);
    for (int32_t y = 0; y < height; ++y) { raw[(size_t)y * (row + 1)] = 0; std::memcpy(&raw[(size_t)y * (row + 1) + 1], rgb + (size_t)y * row, row); }
    std::vector<uint8_t> z; z.reserve(raw_n + raw_n / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);                                   // zlib header, no compression
    uint32_t a1 = 1, a2 = 0;
    for (size_t off = 0; off < raw_n;) {
        const size_t n = std::min<size_t>(65535, raw_n - off);
        z.push_back(off + n == raw_n ? 1 : 0);                               // BFINAL, BTYPE = 00 (stored)
        z.push_back((uint8_t)(n & 0xff)); z.push_back((uint8_t)(n >> 8)); z.push_back((uint8_t)(~n & 0xff)); z.push_back((uint8_t)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + (long)off, raw.begin() + (long)(off + n));
        for (size_t i = off; i < off + n; ++i) { a1 = (a1 + raw[i]) % 65521u; a2 = (a2 + a1) % 65521u; }
        off += n;
    }
    put_be32(z, (a2 << 16) | a1);
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(RT_ERR_IO, std::string("cannot open ") + path);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr; put_be32(ihdr, (uint32_t)width); put_be32(ihdr, (uint32_t)height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8 bit, RGB, deflate, no filter method, no interlace
    const bool ok = std::fwrite(sig, 1, 8, f) == 8 && write_chunk(f, "IHDR", ihdr) && write_chunk(f, "IDAT", z) && write_chunk(f, "IEND", {});
    std::fclose(f);
    return ok ? RT_OK : fail(RT_ERR_IO, std::string("short write to ") + path);
}

int rt_debug_bounds_violations(int32_t device, uint64_t* counts, int32_t reset) {
    if (!counts) return fail(RT_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long v[rtd::RT_BOUNDS_KINDS];
    const cudaError_t e = rtd::read_bounds_violations(v, reset);
    if (e == cudaErrorNotSupported) { cudaGetLastError(); return fail(RT_ERR_INVALID, "not a -DRT_DEBUG_BOUNDS build (make -C csrc debug -> _build_dbg/librt_b200.so)"); }
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("rt_debug_bounds_violations: ") + cudaGetErrorString(e));
    for (int k = 0; k < rtd::RT_BOUNDS_KINDS; ++k) counts[k] = v[k];
    return RT_OK;
}

int rt_measure_fp32_peak(int32_t device, double* tflops, double* sm_mhz) {
    if (!tflops) return fail(RT_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(device));
    int sms = 0, khz = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    float* sink = nullptr;
    CUDA_TRY(cudaMalloc((void**)&sink, 64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256, iters = 1 << 15;
    double best = 0.0;
    cudaError_t e = rtd::launch_ffma(blocks, threads, 1 << 10, sink, 0);
    for (int rep = 0; rep < 5 && e == cudaSuccess; ++rep) {
        cudaEventRecord(e0, 0);
        e = rtd::launch_ffma(blocks, threads, iters, sink, 0);
        cudaEventRecord(e1, 0);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 16.0 * (double)iters * (double)threads * (double)blocks;
        if (ms > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    if (e != cudaSuccess) return fail(RT_ERR_CUDA, std::string("ffma benchmark: ") + cudaGetErrorString(e));
    *tflops = best;
    if (sm_mhz) *sm_mhz = (double)khz / 1000.0;
    return RT_OK;
}

}  // extern "C"
