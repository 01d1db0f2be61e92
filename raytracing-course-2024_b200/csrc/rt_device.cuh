// rt_device.cuh -- device-side building blocks of the sm_100a path tracer: vector math, Philox4x32-10,
// shared/global scene accessors, ray/box and ray/triangle tests, BVH traversal, the glTF metallic-roughness BRDF
// and the cosine / GGX-VNDF / light samplers with their pdfs.  Every function cites the reference lines whose
// SEMANTICS it reproduces (FP32 instead of the reference's f64 -- see DESIGN.md "numerics").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtd {

#define RT_DEV __device__ __forceinline__
#define RT_PI_F 3.14159265358979323846f
#define RT_INV_PI_F 0.31830988618379067154f
#define RT_EPS_F 0.00001f /* geometry.rs:49 */
#define RT_INF_F __int_as_float(0x7f800000)

// ------------------------------------------------------------------------------------------------ debug bounds build
// compute-sanitizer is closed on the B200 pool this repository is developed against, so memory safety gets its own instrument:
// `make debug` compiles the same sources with -DRT_DEBUG_BOUNDS into _build_dbg/.  Every shared-memory pool / queue / stack access,
// every scene-blob load, every primitive index and every layer store is then range-checked; a violation is COUNTED per kind (and the
// access is still performed -- the point is to find it, on small cases, without killing the context) and read back through
// rt_debug_bounds_violations().  tests/test_gpu_debug_bounds.py runs the parity workloads under this build and asserts all-zero.
enum { RT_BOUNDS_POOL = 0, RT_BOUNDS_QUEUE, RT_BOUNDS_STACK, RT_BOUNDS_BLOB, RT_BOUNDS_PRIM, RT_BOUNDS_LAYER, RT_BOUNDS_CHUNK, RT_BOUNDS_KINDS };
#ifdef RT_DEBUG_BOUNDS
static __device__ unsigned long long g_bounds_violations[RT_BOUNDS_KINDS];
#define RT_BOUNDS(cond, kind) do { if (!(cond)) atomicAdd(&rtd::g_bounds_violations[kind], 1ull); } while (0)
#else
#define RT_BOUNDS(cond, kind) ((void)0)
#endif

// ------------------------------------------------------------------------------------------------ float3 math
RT_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
RT_DEV float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
RT_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
RT_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
RT_DEV float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
RT_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_DEV float dot(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
RT_DEV float3 cross(float3 a, float3 b) { return f3(fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x)); }
RT_DEV float3 fma3(float3 a, float s, float3 b) { return f3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }  // a*s + b
RT_DEV float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // one MUFU.RSQ, no denormal rescaling
RT_DEV float3 normalize(float3 a) { return a * fast_rsqrt(dot(a, a)); }
RT_DEV float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
RT_DEV float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // MUFU.SQRT, ~1 ulp
RT_DEV bool finite3(float3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }

// ------------------------------------------------------------------------------------------------ Philox4x32-10
// Counter-based RNG (Salmon et al. 2011) replacing the reference's per-row xoshiro256** stream
// (rendering.rs:50-51): counter = (pixel, sample, call, tag), key = 64-bit user seed, so the image does not
// depend on how pixels and samples are sharded over lanes, CTAs or GPUs.
RT_DEV uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}
#define RT_PHILOX_TAG 0x52544232u /* "RTB2" */
RT_DEV float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0, 1)

// ------------------------------------------------------------------------------------------------ scene blob
// The device scene is ONE 16-byte-aligned blob of SoA arrays; SceneLayout holds byte offsets into it.  Render
// kernels either read it through the read-only global path (ld.global.nc) or copy it to shared memory once per
// CTA when it fits (practice7_4: ~12 KB) and read it with ld.shared -- same code, different Space.
struct SceneLayout {
    uint32_t nodes;                              // finite-primitive BVH: octant-ordered child-pair nodes, RT_NODE_BYTES each (below)
    uint32_t tri_t;                              // BVH-ordered triangle TEST records, 48 bytes each (TriTest below)
    uint32_t tri_a, tri_e1, tri_e2;              // BVH-ordered triangles: a, b-a, c-a (w unused) -- hit-point evaluation
    uint32_t sh_n0, sh_dn1, sh_dn2, sh_ng;       // a_norm|material id, b_norm-a_norm|orig id, c_norm-a_norm, unit face normal
    uint32_t mat0, mat1;                         // base rgb|metallic, emission rgb|roughness
    uint32_t lt_a, lt_e1, lt_e2;                 // light triangles (light-BVH order): a, e1, e2 -- point sampling
    uint32_t lt_t;                               // light TEST records, 64 bytes each: N|1/area, a, Au, Av
    uint32_t lnodes;                             // light BVH, same node format (used when n_lights > RT_BRUTE_LIGHTS)
    uint32_t total_bytes;
    int32_t n_nodes, n_tris, n_mats, n_lights, n_lnodes;
    int32_t light_bvh;                           // 1: pdf walks the light BVH, 0: loops over all lights
    int32_t max_leaf;                            // largest leaf of the finite-primitive BVH (the kernel unrolls leaves of <= 2 triangles)
    float inv_n_lights;                          // 1 / n_lights (MultipleLightSamplingDistribution::pdf, distributions.rs:183)
    int32_t packed_refs;                         // 1: the x planes of `nodes` carry 16-bit child references (pair_step PACKED)
    // ---- general-primitive scenes (Shape3D::Box + Object3D transforms, geometry.rs:27-46; text-scene shapes): `general` = 1
    int32_t general;                             // 1: leaves index `prims` (RT_PRIM_BYTES records) instead of `tri_t`
    uint32_t prims;                              // n_tris finite records in BVH order, then n_planes infinite ones (scene.rs:37)
    uint32_t lt_g;                               // light records (same format, light order) for sampling and the pdf
    uint32_t mat2;                               // per material: ior, material kind (as int bits), 0, 0
    int32_t n_planes;
    int32_t has_dielectric;                      // any material of kind RT_MAT_DIELECTRIC
    // ---- quantised 32-byte pair nodes for the global-memory walks (pair_step_quant): `quant` = 1
    uint32_t qnodes;                             // 32-byte aligned
    int32_t quant;
    float qorg[3], qcell[3];                     // plane = qorg + q * qcell
    int32_t flat_scan;                           // general scene of <= RT_FLAT_SCAN_MAX primitives: the render kernel scans `prims` instead of walking `nodes`
};
#define RT_BRUTE_LIGHTS 8
#define RT_FLAT_SCAN_MAX 12

struct SmemSpace {
    uint32_t base;  // shared-window address of the blob
#ifdef RT_DEBUG_BOUNDS
    uint32_t limit;  // blob bytes
#endif
    RT_DEV float4 ld4(uint32_t off) const {
        RT_BOUNDS((off & 15u) == 0u && off + 16u <= limit, RT_BOUNDS_BLOB);
        float4 v;
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + off));
        return v;
    }
    RT_DEV int2 ld2i(uint32_t off) const {
        RT_BOUNDS((off & 7u) == 0u && off + 8u <= limit, RT_BOUNDS_BLOB);
        int2 v;
        asm("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(base + off));
        return v;
    }
};
struct GmemSpace {
    const char* base;
#ifdef RT_DEBUG_BOUNDS
    uint32_t limit;
#endif
    RT_DEV float4 ld4(uint32_t off) const {
        RT_BOUNDS((off & 15u) == 0u && off + 16u <= limit, RT_BOUNDS_BLOB);
        return __ldg(reinterpret_cast<const float4*>(base + off));
    }
    RT_DEV int2 ld2i(uint32_t off) const {
        RT_BOUNDS((off & 7u) == 0u && off + 8u <= limit, RT_BOUNDS_BLOB);
        return __ldg(reinterpret_cast<const int2*>(base + off));
    }
    // One 32-byte sector with ONE instruction (LDG.E.256, sm_100).  The L1 data pipe spends a cycle per (lane, sector) on the
    // scattered loads of a traversal whatever their width: 32 bytes per cycle instead of the 16 of LDG.128.
    RT_DEV void ld8(uint32_t off, float4& a, float4& b) const {
        RT_BOUNDS((off & 31u) == 0u && off + 32u <= limit, RT_BOUNDS_BLOB);
        asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(base + off));
    }
};

// Per-thread traversal stack in shared memory, laid out [entry][thread]: bank = thread % 32 for every entry, so pushes
// and pops are conflict-free however far the lanes' stack depths have diverged.  Entry 0 holds RT_CUR_DONE for good:
// a pop never tests for an empty stack, the traversal is over when the popped reference is RT_CUR_DONE.
#define RT_CUR_DONE ((int)0x80000000) /* never a node reference (>= 0) nor a leaf code (~code > INT_MIN) */
struct SmemStack {
    uint32_t top;     // shared-window address of the next free entry of this thread
    uint32_t stride;  // bytes between entries = 4 * blockDim.x
#ifdef RT_DEBUG_BOUNDS
    uint32_t lo, hi;  // addresses of this thread's entry 0 and of the entry past the last one
#endif
    RT_DEV void init(uint32_t thread_entry0, uint32_t stride_bytes, uint32_t entries = 0u) {
        stride = stride_bytes; top = thread_entry0 + stride_bytes;
#ifdef RT_DEBUG_BOUNDS
        lo = thread_entry0; hi = thread_entry0 + entries * stride_bytes;
#else
        (void)entries;
#endif
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(thread_entry0), "r"(RT_CUR_DONE));
    }
    RT_DEV void reset(uint32_t thread_entry0) { top = thread_entry0 + stride; }
    // `top` must stay in (lo, hi]: entry 0 (the sentinel) is never popped, the last entry may be pushed but not passed
    RT_DEV void check() const { RT_BOUNDS(top > lo && top <= hi, RT_BOUNDS_STACK); }
    RT_DEV void push(int v) { RT_BOUNDS(top < hi, RT_BOUNDS_STACK); asm volatile("st.shared.s32 [%0], %1;" ::"r"(top), "r"(v)); top += stride; }
    RT_DEV int pop() { int v; top -= stride; RT_BOUNDS(top >= lo, RT_BOUNDS_STACK); asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(top)); return v; }
};

struct Counters { unsigned long long node_tests, tri_tests, light_tri_tests; };

// ------------------------------------------------------------------------------------------------ intersection
// Reciprocal direction for the slab tests; zero components are replaced by +-1e-20 before inversion (the reference
// adds 1e-8 to the direction instead, geometry.rs:144-155).
RT_DEV float3 safe_inv_dir(float3 d) {
    const float tiny = 1e-20f;
    float x = fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x;
    float y = fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y;
    float z = fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z;
    return f3(fast_rcp(x), fast_rcp(y), fast_rcp(z));
}
// Octant-ordered child-pair node (RT_NODE_BYTES = 112, 16-byte aligned).  A node keeps the boxes of BOTH children;
// per axis it stores the four planes twice, once in each order, so that a ray picks its (near c0, near c1, far c0,
// far c1) planes with ONE 16-byte load at an offset chosen by the sign of its direction -- no per-axis min/max:
//   [ 0, 32) x: c0.min c1.min c0.max c1.max | c0.max c1.max c0.min c1.min
//   [32, 64) y: same                          [64, 96) z: same
//   [96,104) child references;  [104,112) padding
// A child reference is >= 0 for an inner node (its BYTE offset inside the node array = index * 112, so a visit needs no
// multiply) and < 0 for a leaf: ~((first_tri << 3) | (count - 1)).  The root is reference 0.  Scenes with < 32768 nodes
// and < 4096 triangles ALSO carry 16-bit references (node index / leaf code) in the low bytes of the x planes: see
// pair_step<.., PACKED> and SceneLayout::packed_refs.
#define RT_NODE_BYTES 112u
struct RaySetup {
    float3 inv, od;          // 1/d (zero components replaced, safe_inv_dir) and o/d
    uint32_t ox, oy, oz;     // blob offsets of this ray's x / y / z plane quadruples of node 0
};
RT_DEV void ray_octant(RaySetup& r, uint32_t nodes) {
    r.ox = nodes + ((__float_as_uint(r.inv.x) >> 27) & 16u);
    r.oy = nodes + 32u + ((__float_as_uint(r.inv.y) >> 27) & 16u);
    r.oz = nodes + 64u + ((__float_as_uint(r.inv.z) >> 27) & 16u);
}
RT_DEV RaySetup ray_setup(float3 o, float3 d, uint32_t nodes) {
    RaySetup r;
    r.inv = safe_inv_dir(d);
    r.od = o * r.inv;
    ray_octant(r, nodes);
    return r;
}
RT_DEV float max3f(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
RT_DEV float min3f(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// Child selection of a box-pair step, written in PTX so that it stays in predicate registers (nvcc turns the bool algebra into
// integer SEL / LOP3 / PRMT chains otherwise): enter the nearer hit child (pushing the other one), or pop.
template <bool ORDERED>
RT_DEV void pair_descend(float t0, float t1, float e0, float e1, int2 ch, int& cur, SmemStack& st) {
    asm volatile(
        "{\n\t"
        ".reg .pred h0, h1, sf, both, any;\n\t"
        ".reg .b32 first, second;\n\t"
        "setp.le.f32 h0, %2, %4;\n\t"
        "setp.le.f32 h1, %3, %5;\n\t"
        "setp.lt.f32 sf, %3, %9;\n\t"                // second child first iff it is strictly nearer (ordered walks only) ...
        "not.pred both, h0;\n\t"
        "or.pred sf, sf, both;\n\t"                  // ... or the first child is not hit at all
        "and.pred sf, sf, h1;\n\t"
        "and.pred both, h0, h1;\n\t"
        "or.pred any, h0, h1;\n\t"
        "selp.b32 first, %7, %6, sf;\n\t"
        "selp.b32 second, %6, %7, sf;\n\t"
        "@both st.shared.b32 [%1], second;\n\t"
        "@both add.u32 %1, %1, %8;\n\t"
        "@any mov.b32 %0, first;\n\t"
        "@!any sub.u32 %1, %1, %8;\n\t"
        "@!any ld.shared.b32 %0, [%1];\n\t"
        "}"
        : "+r"(cur), "+r"(st.top)
        : "f"(t0), "f"(t1), "f"(e0), "f"(e1), "r"(ch.x), "r"(ch.y), "r"(st.stride), "f"(ORDERED ? t0 : t1));
}

// One box-pair step of the near-child-first traversal (interesect_with_bvh_nearest_point, bvh.rs:231-297, iterative):
// conservative slab test of the two child boxes of node `cur`, then descend into the nearer hit child (pushing the
// other one), or pop.  Accept/prune rule of get_aabb_intersection + bvh.rs:258-263: a box is entered iff the ray's
// [max(t_entry, 0), min(t_exit, best)] interval is non-empty.  Boxes are EPS-padded (1e-5, aabb.rs:53-65) and rounded
// outward: ~20 ulps of slack for the FP32 arithmetic.  ORDERED = false: all-hits walk (children in storage order).
// The child selection is written in PTX so that it stays in predicate registers (nvcc turns the bool algebra into
// integer SEL / LOP3 / PRMT chains otherwise).
// PACKED (small scenes, the shared-memory render kernel): node references are node INDICES / leaf codes of 16 bits, and
// the two references of a node ride in the low mantissa bytes of the four x planes the ray loads anyway (the host rounds
// those planes outward by up to 2^-15 relative, so the boxes stay conservative) -- the 8-byte child load disappears.
template <class Space, bool ORDERED = true, bool PACKED = false>
RT_DEV void pair_step(const Space& sp, uint32_t nodes, const RaySetup& r, float t_best, int& cur, SmemStack& st) {
    const uint32_t c = PACKED ? (uint32_t)cur * RT_NODE_BYTES : (uint32_t)cur;
    const float4 X = sp.ld4(c + r.ox), Y = sp.ld4(c + r.oy), Z = sp.ld4(c + r.oz);
    int2 ch;
    if (PACKED) {
        asm("prmt.b32 %0, %1, %2, 0xCC40;" : "=r"(ch.x) : "r"(__float_as_int(X.x)), "r"(__float_as_int(X.y)));   // sign-extended 16 bits
        asm("prmt.b32 %0, %1, %2, 0xCC40;" : "=r"(ch.y) : "r"(__float_as_int(X.z)), "r"(__float_as_int(X.w)));
    } else {
        ch = sp.ld2i(c + (nodes + 96u));
    }
    const float nx0 = fmaf(X.x, r.inv.x, -r.od.x), nx1 = fmaf(X.y, r.inv.x, -r.od.x), fx0 = fmaf(X.z, r.inv.x, -r.od.x), fx1 = fmaf(X.w, r.inv.x, -r.od.x);
    const float ny0 = fmaf(Y.x, r.inv.y, -r.od.y), ny1 = fmaf(Y.y, r.inv.y, -r.od.y), fy0 = fmaf(Y.z, r.inv.y, -r.od.y), fy1 = fmaf(Y.w, r.inv.y, -r.od.y);
    const float nz0 = fmaf(Z.x, r.inv.z, -r.od.z), nz1 = fmaf(Z.y, r.inv.z, -r.od.z), fz0 = fmaf(Z.z, r.inv.z, -r.od.z), fz1 = fmaf(Z.w, r.inv.z, -r.od.z);
    const float t0 = fmaxf(max3f(nx0, ny0, nz0), 0.0f), e0 = fminf(min3f(fx0, fy0, fz0), t_best);
    const float t1 = fmaxf(max3f(nx1, ny1, nz1), 0.0f), e1 = fminf(min3f(fx1, fy1, fz1), t_best);
    pair_descend<ORDERED>(t0, t1, e0, e1, ch, cur, st);
#ifdef RT_DEBUG_BOUNDS
    RT_BOUNDS(st.top >= st.lo && st.top <= st.hi, RT_BOUNDS_STACK);      // (== lo: the sentinel was popped, the walk is over)
#endif
}

// QUANTISED child-pair node for the global-memory walks: 32 bytes = ONE sector, read with one LDG.E.256.  The twelve box planes are
// 16-bit coordinates on a global grid over the scene box (SceneLayout::qorg / qcell; minima rounded down, maxima up, two cells of
// slack for the FP32 evaluation), the two child references stay 32-bit (inner: byte offset = index * 32):
//   word 0 / 1: x minima / maxima (child 0 in the low half, child 1 in the high half), 2 / 3: y, 4 / 5: z, 6 / 7: references.
// Why: the large meshes are bound by the L1 data pipe (ncu: l1tex__data_pipe_lsu_wavefronts 90-93 % of peak) and a scattered load
// costs about one wavefront per lane and INSTRUCTION (microbenchmark tests/tools/microbench/l1_gather.cu): the octant-ordered
// 112-byte node needs four loads per step, this one needs one, and the node array shrinks 3.5x (L1 / L2 hit rates).
// The ray carries s = qcell / d and b = (o - qorg) / d per axis (RaySetup::inv / od), so a plane distance is fma(q, s, -b).
#define RT_QNODE_BYTES 32u
template <bool ORDERED = true>
RT_DEV void pair_step_quant(const GmemSpace& sp, uint32_t qnodes, const RaySetup& r, float t_best, int& cur, SmemStack& st) {
    float4 A, B;
    sp.ld8(qnodes + (uint32_t)cur, A, B);
    const bool sx = r.inv.x < 0.0f, sy = r.inv.y < 0.0f, sz = r.inv.z < 0.0f;
    const uint32_t xn = __float_as_uint(sx ? A.y : A.x), xf = __float_as_uint(sx ? A.x : A.y);
    const uint32_t yn = __float_as_uint(sy ? A.w : A.z), yf = __float_as_uint(sy ? A.z : A.w);
    const uint32_t zn = __float_as_uint(sz ? B.y : B.x), zf = __float_as_uint(sz ? B.x : B.y);
    const float nx0 = fmaf((float)(xn & 0xffffu), r.inv.x, -r.od.x), nx1 = fmaf((float)(xn >> 16), r.inv.x, -r.od.x);
    const float fx0 = fmaf((float)(xf & 0xffffu), r.inv.x, -r.od.x), fx1 = fmaf((float)(xf >> 16), r.inv.x, -r.od.x);
    const float ny0 = fmaf((float)(yn & 0xffffu), r.inv.y, -r.od.y), ny1 = fmaf((float)(yn >> 16), r.inv.y, -r.od.y);
    const float fy0 = fmaf((float)(yf & 0xffffu), r.inv.y, -r.od.y), fy1 = fmaf((float)(yf >> 16), r.inv.y, -r.od.y);
    const float nz0 = fmaf((float)(zn & 0xffffu), r.inv.z, -r.od.z), nz1 = fmaf((float)(zn >> 16), r.inv.z, -r.od.z);
    const float fz0 = fmaf((float)(zf & 0xffffu), r.inv.z, -r.od.z), fz1 = fmaf((float)(zf >> 16), r.inv.z, -r.od.z);
    const float t0 = fmaxf(max3f(nx0, ny0, nz0), 0.0f), e0 = fminf(min3f(fx0, fy0, fz0), t_best);
    const float t1 = fmaxf(max3f(nx1, ny1, nz1), 0.0f), e1 = fminf(min3f(fx1, fy1, fz1), t_best);
    pair_descend<ORDERED>(t0, t1, e0, e1, make_int2(__float_as_int(B.z), __float_as_int(B.w)), cur, st);
#ifdef RT_DEBUG_BOUNDS
    RT_BOUNDS(st.top >= st.lo && st.top <= st.hi, RT_BOUNDS_STACK);
#endif
}
// Ray set-up for the quantised walk, and the way back to (origin, direction) at a leaf (the wavefront kernel keeps only RaySetup).
RT_DEV RaySetup ray_setup_quant(float3 o, float3 inv_d, const float* qorg, const float* qcell) {
    RaySetup r;
    r.inv = f3(qcell[0] * inv_d.x, qcell[1] * inv_d.y, qcell[2] * inv_d.z);
    r.od = f3((o.x - qorg[0]) * inv_d.x, (o.y - qorg[1]) * inv_d.y, (o.z - qorg[2]) * inv_d.z);
    r.ox = r.oy = r.oz = 0u;
    return r;
}
RT_DEV void ray_from_quant(const RaySetup& r, const float* qorg, const float* qcell, float3& o, float3& d) {
    d = f3(qcell[0] * fast_rcp(r.inv.x), qcell[1] * fast_rcp(r.inv.y), qcell[2] * fast_rcp(r.inv.z));
    o = f3(fmaf(r.od.x, d.x, qorg[0]), fmaf(r.od.y, d.y, qorg[1]), fmaf(r.od.z, d.z, qorg[2]));
}

// Two-sided ray/triangle test with inclusive edges: u >= 0, v >= 0, u + v <= 1, t > 0 (geometry.rs:109-113).  The
// reference solves [b-a, c-a, -d] (u,v,t)^T = o-a with Matrix3::try_inverse (geometry.rs:103-111).  Here the
// ray-independent part of that solution is precomputed per triangle on the host in f64 (TriTest):
//   N  = unit normal of (e1 x e2),  Au = (e2 x Nu)/|Nu|^2,  Av = (Nu x e1)/|Nu|^2   with Nu = e1 x e2,
// so that   t = N.(a-o) / N.d,   Q = o + t d - a,   u = Au.Q,   v = Av.Q   (19 flops + 1 rcp instead of 30 + 1).
// N.d == 0 (the reference's det == 0) -> inf/NaN -> the strict `t < best` / the comparisons reject it; a degenerate
// triangle has non-finite Au/Av and is never hit, like try_inverse failing.
struct TriTest { float3 N, a, Au, Av; };
template <class Space>
RT_DEV TriTest load_tri_test(const Space& sp, uint32_t tri_t, int i) {     // packed 48-byte record
    const uint32_t o = tri_t + (uint32_t)i * 48u;
    const float4 r0 = sp.ld4(o), r1 = sp.ld4(o + 16u), r2 = sp.ld4(o + 32u);
    TriTest T;
    T.N = f3(r0.x, r0.y, r0.z); T.a = f3(r0.w, r1.x, r1.y); T.Au = f3(r1.z, r1.w, r2.x); T.Av = f3(r2.y, r2.z, r2.w);
    return T;
}
RT_DEV bool tri_test(float3 o, float3 d, const TriTest& T, float& t, float& u, float& v) {
    const float3 oa = T.a - o;
    t = dot(T.N, oa) * fast_rcp(dot(T.N, d));
    const float3 Q = fma3(d, t, -oa);
    u = dot(T.Au, Q);
    v = dot(T.Av, Q);
    // Edges are inclusive in the reference (f64, effectively watertight).  In FP32 the two triangles sharing an
    // edge can both round a ray just outside; a 2e-6 barycentric overlap closes those cracks for well-conditioned
    // triangles (overlap is harmless: the nearest hit still wins).
    const float tol = 2e-6f;
    return (u >= -tol) & (v >= -tol) & (u + v <= 1.0f + tol) & (t > 0.0f);
}

struct Hit { float t, u, v; int tri; };

// Nearest hit over the finite-primitive BVH: interesect_with_bvh_nearest_point (bvh.rs:231-297) with an
// iterative, near-child-first traversal.  Strict `<` keeps the first-found hit on exact ties (bvh.rs:269); the
// visiting order differs from the reference's unordered DFS, so only exact ties can resolve differently.
// skip_tri: triangle (BVH order) to ignore, or -1 -- used to make the reference's `t - EPS` origin back-off
// (rendering.rs:98) robust in FP32: a ray that LEAVES the side of the surface it started on cannot hit the
// triangle it started from in exact arithmetic, so that triangle is skipped instead of relying on a 1e-5 gap.
template <class Space, bool STATS, int NODES = 0>
RT_DEV void trace_nearest(const Space& sp, const SceneLayout& L, SmemStack& st, float3 o, float3 d, int skip_tri, Hit& hit, Counters& cnt) {
    const RaySetup r = NODES == 2 ? ray_setup_quant(o, safe_inv_dir(d), L.qorg, L.qcell) : ray_setup(o, d, L.nodes);
    hit.t = RT_INF_F; hit.u = 0.f; hit.v = 0.f; hit.tri = -1;
    st.push(RT_CUR_DONE);                        // marker: the walk may start on top of a live stack
    int cur = 0;
    for (;;) {
        while (cur >= 0) {
            if constexpr (NODES == 2) pair_step_quant(sp, L.qnodes, r, hit.t, cur, st);
            else pair_step(sp, L.nodes, r, hit.t, cur, st);
            if (STATS) cnt.node_tests += 2;
        }
        if (cur == RT_CUR_DONE) break;
        const uint32_t code = (uint32_t)~cur;
        const int first = (int)(code >> 3), n = (int)(code & 7u) + 1;
        for (int i = first; i < first + n; ++i) {
            float t, u, v;
            const bool ok = tri_test(o, d, load_tri_test(sp, L.tri_t, i), t, u, v);
            if (STATS) cnt.tri_tests += 1;
            if (ok && t < hit.t && i != skip_tri) { hit.t = t; hit.u = u; hit.v = v; hit.tri = i; }
        }
        cur = st.pop();
    }
}

// ------------------------------------------------------------------------------------------------ general primitives
// Scenes that hold more than world-space triangles (Shape3D::Box, Object3D position / rotation: geometry.rs:27-46,196-251;
// the text-scene shapes ELLIPSOID and PLANE: own spec) keep ONE 64-byte record per primitive:
//   triangle : r0 = N | kind,  r1 = a | 1/area,  r2 = Au,  r3 = Av          (world space: the object transform is applied on the
//              host -- the reference rotates the ray instead, geometry.rs:201-214; same hit, same t)
//   others   : r0.xyz, r1.xyz, r2.xyz = rows of R^T (world -> object rotation = the conjugate quaternion of geometry.rs:206-213),
//              r3.xyz = position, (r1.w, r2.w, r3.w) = s: box half sizes | ellipsoid radii | plane normal;  r0.w = kind
#define RT_PRIM_BYTES 64u
enum { RT_KIND_TRIANGLE = 0, RT_KIND_BOX = 1, RT_KIND_ELLIPSOID = 2, RT_KIND_PLANE = 3 };
enum { RT_MAT_PBR = 0, RT_MAT_DIELECTRIC = 1 };
#define RT_AUX_EXIT 4 /* aux bit: the hit leaves the solid (is_outer_to_inner == false, geometry.rs:180-188) */
struct PrimRec { float4 r0, r1, r2, r3; };
template <class Space>
RT_DEV PrimRec load_prim(const Space& sp, uint32_t table, int i) {
    const uint32_t o = table + (uint32_t)i * RT_PRIM_BYTES;
    PrimRec R; R.r0 = sp.ld4(o); R.r1 = sp.ld4(o + 16u); R.r2 = sp.ld4(o + 32u); R.r3 = sp.ld4(o + 48u);
    return R;
}
RT_DEV int prim_kind(const PrimRec& R) { return __float_as_int(R.r0.w); }
RT_DEV float3 prim_s(const PrimRec& R) { return f3(R.r1.w, R.r2.w, R.r3.w); }
RT_DEV float3 prim_to_local(const PrimRec& R, float3 v) { return f3(dot(f3(R.r0), v), dot(f3(R.r1), v), dot(f3(R.r2), v)); }   // q.conjugate() * v
RT_DEV float3 prim_to_world(const PrimRec& R, float3 v) { return f3(R.r0) * v.x + f3(R.r1) * v.y + f3(R.r2) * v.z; }          // q * v

// Face of a box hit: the slab that produced it, overridden by the reference's `s - |p| < EPS` cascade (geometry.rs:161-169: x
// before y before z) when the hit point lies within EPS of a higher-priority face -- the same answer as the reference at edges
// and corners, and a robust one (the slab) in the interior of a face, where FP32 could not resolve `s - |p| < 1e-5` reliably.
RT_DEV int box_face(float3 s, float3 ol, float3 dl, float t, int slab_axis) {
    const float3 p = fma3(dl, t, ol);
    if (s.x - fabsf(p.x) < RT_EPS_F) return 0;
    if (slab_axis == 2 && s.y - fabsf(p.y) < RT_EPS_F) return 1;
    return slab_axis;
}

// FIRST hit of a primitive = what competes in the nearest-hit query (points.0[0], bvh.rs:269) and what
// intersect_ray_with_object3d returns (geometry.rs:51-58,196-223).  Out: t; triangles: (u, v) barycentrics; others: u = t,
// v = aux bits (box: axis 0..2 of the face; RT_AUX_EXIT when the ray leaves the solid / hits the back of a plane).
//   box       geometry.rs:140-194: object-space slabs, entry hit if t_min > 0, else the exit hit if t_max > 0.  Face: box_face.
//   ellipsoid own spec: |(o + t d)/r|^2 = 1 around the closest approach (no cancellation in FP32); outside -> entry root if the
//             ray approaches, inside -> exit root.  Same roots as the oracle's half-b form.
//   plane     own spec: t = n.(pos - o) / n.d with the world normal n = R s.
RT_DEV bool prim_first_hit(const PrimRec& R, float3 o, float3 d, float& t, float& u, float& v) {
    const int kind = prim_kind(R);
    if (kind == RT_KIND_TRIANGLE) {
        TriTest T; T.N = f3(R.r0); T.a = f3(R.r1); T.Au = f3(R.r2); T.Av = f3(R.r3);
        return tri_test(o, d, T, t, u, v);
    }
    const float3 s = prim_s(R);
    const float3 rel = o - f3(R.r3);
    if (kind == RT_KIND_PLANE) {
        const float3 nw = prim_to_world(R, s);
        const float dn = dot(nw, d);
        t = -dot(nw, rel) * fast_rcp(dn);
        u = t; v = __int_as_float(dn < 0.0f ? 0 : RT_AUX_EXIT);
        return t > 0.0f;                                   // dn == 0: +-inf / NaN, rejected here or by the caller's `t < best`
    }
    const float3 ol = prim_to_local(R, rel), dl = prim_to_local(R, d);
    if (kind == RT_KIND_BOX) {
        const float3 inv = safe_inv_dir(dl);
        const float ax = (-s.x - ol.x) * inv.x, bx = (s.x - ol.x) * inv.x;
        const float ay = (-s.y - ol.y) * inv.y, by = (s.y - ol.y) * inv.y;
        const float az = (-s.z - ol.z) * inv.z, bz = (s.z - ol.z) * inv.z;
        const float nx = fminf(ax, bx), fx = fmaxf(ax, bx), ny = fminf(ay, by), fy = fmaxf(ay, by), nz = fminf(az, bz), fz = fmaxf(az, bz);
        const float tmin = max3f(nx, ny, nz), tmax = min3f(fx, fy, fz);
        if (!(tmin <= tmax)) return false;
        const bool entry = tmin > 0.0f;
        t = entry ? tmin : tmax;
        const int slab = entry ? ((nx >= ny && nx >= nz) ? 0 : (ny >= nz ? 1 : 2)) : ((fx <= fy && fx <= fz) ? 0 : (fy <= fz ? 1 : 2));
        const int axis = box_face(s, ol, dl, t, slab);
        u = t; v = __int_as_float(axis | (entry ? 0 : RT_AUX_EXIT));
        return t > 0.0f;
    }
    // ellipsoid
    const float3 ir = f3(fast_rcp(s.x), fast_rcp(s.y), fast_rcp(s.z));
    const float3 od = ol * ir, dd = dl * ir;
    const float a = dot(dd, dd), hb = dot(od, dd), c = dot(od, od) - 1.0f;
    const float ia = fast_rcp(a);
    const float tc = -hb * ia;
    const float3 pc = fma3(dd, tc, od);                    // closest approach to the centre, unit-sphere space
    const float h2 = (1.0f - dot(pc, pc)) * ia;
    if (!(h2 >= 0.0f)) return false;
    const float h = fast_sqrt(h2);
    const bool inside = c < 0.0f;
    t = inside ? tc + h : tc - h;
    u = t; v = __int_as_float(inside ? RT_AUX_EXIT : 0);
    return inside ? t > 0.0f : (hb < 0.0f && t > 0.0f);
}

// intersect_ray_with_scene (rendering.rs:201-226) for general-primitive scenes: the nearest-hit walk of trace_nearest with
// prim_first_hit at the leaves, then the linear scan of the infinite primitives (records n_tris .. n_tris + n_planes) under
// the running bound with strict `<` (:215-224).  hit.tri indexes `prims`; (hit.u, hit.v) = barycentrics or (t, aux).
// skip_prim: see trace_nearest -- set by the caller only when a re-hit is impossible in exact arithmetic.
template <class Space, bool STATS, int NODES = 0>
RT_DEV void trace_nearest_gen(const Space& sp, const SceneLayout& L, SmemStack& st, float3 o, float3 d, int skip_prim, Hit& hit, Counters& cnt) {
    const RaySetup r = NODES == 2 ? ray_setup_quant(o, safe_inv_dir(d), L.qorg, L.qcell) : ray_setup(o, d, L.nodes);
    hit.t = RT_INF_F; hit.u = 0.f; hit.v = 0.f; hit.tri = -1;
    st.push(RT_CUR_DONE);
    int cur = 0;
    for (;;) {
        while (cur >= 0) {
            if constexpr (NODES == 2) pair_step_quant(sp, L.qnodes, r, hit.t, cur, st);
            else pair_step(sp, L.nodes, r, hit.t, cur, st);
            if (STATS) cnt.node_tests += 2;
        }
        if (cur == RT_CUR_DONE) break;
        const uint32_t code = (uint32_t)~cur;
        const int first = (int)(code >> 3), n = (int)(code & 7u) + 1;
        for (int i = first; i < first + n; ++i) {
            float t, u, v;
            const bool ok = prim_first_hit(load_prim(sp, L.prims, i), o, d, t, u, v);
            if (STATS) cnt.tri_tests += 1;
            if (ok && t < hit.t && i != skip_prim) { hit.t = t; hit.u = u; hit.v = v; hit.tri = i; }
        }
        cur = st.pop();
    }
    for (int i = L.n_tris; i < L.n_tris + L.n_planes; ++i) {
        float t, u, v;
        const bool ok = prim_first_hit(load_prim(sp, L.prims, i), o, d, t, u, v);
        if (STATS) cnt.tri_tests += 1;
        if (ok && t < hit.t && i != skip_prim) { hit.t = t; hit.u = u; hit.v = v; hit.tri = i; }
    }
}

// What get_ray_color (rendering.rs:96-101,107) reads from the hit of a non-triangle primitive: normal_geometry in WORLD space
// (rotated back, geometry.rs:245-249), normal_shading LEFT IN OBJECT SPACE (the reference does not rotate it back; it only
// gates acceptance, rendering.rs:107), both facing the incoming ray; `outer` = is_outer_to_inner.
struct HitFrame { float3 n, ns; bool outer; };
RT_DEV HitFrame prim_frame(const PrimRec& R, float3 o, float3 d, float t, int aux) {
    const int kind = prim_kind(R);
    HitFrame F;
    F.outer = !(aux & RT_AUX_EXIT);
    float3 nl;
    if (kind == RT_KIND_PLANE) {
        nl = prim_s(R);
        if (!F.outer) nl = -nl;
    } else if (kind == RT_KIND_BOX) {
        const float3 dl = prim_to_local(R, d);
        const int axis = aux & 3;
        const float da = axis == 0 ? dl.x : (axis == 1 ? dl.y : dl.z);
        const float sg = da > 0.0f ? -1.0f : 1.0f;        // entry: the face whose outward normal opposes d; exit: -(outward normal)
        nl = f3(axis == 0 ? sg : 0.0f, axis == 1 ? sg : 0.0f, axis == 2 ? sg : 0.0f);
    } else {
        const float3 s = prim_s(R);
        const float3 pl = prim_to_local(R, fma3(d, t, o - f3(R.r3)));
        const float3 ir2 = f3(fast_rcp(s.x * s.x), fast_rcp(s.y * s.y), fast_rcp(s.z * s.z));
        nl = normalize(pl * ir2);
        if (!F.outer) nl = -nl;
    }
    F.ns = nl;
    F.n = prim_to_world(R, nl);
    return F;
}

// ------------------------------------------------------------------------------------------------ BRDF pieces
// GGX terms shared by specular_brdf (rendering.rs:157-184) and the VNDF pdf (distributions.rs:236-260,
// 276-297).  In the local frame of distributions.rs the half vector has x^2+y^2 = 1-(n.h)^2, so Dn(:245-252)
// equals the D of rendering.rs:162-167 without chi+, and g1(:236-243) equals the G1 of rendering.rs:169-179.
//   D = alpha^2 / (pi * ((alpha^2-1)(n.h)^2 + 1)^2); the denominator is evaluated as alpha^2 (n.h)^2 + |n x h|^2
//   to avoid the catastrophic FP32 cancellation at alpha^2 = 8.1e-7 (roughness 0.03).
RT_DEV float ggx_d_nochi(float3 n, float3 h, float alpha2) {
    const float nh = dot(n, h);
    const float3 c = cross(n, h);
    const float den = fmaf(alpha2 * nh, nh, dot(c, c));
    return alpha2 * fast_rcp(RT_PI_F * den * den);
}
// G1 = 1/(1+Lambda), Lambda = (sqrt(1 + alpha^2 (1-nx^2)/nx^2) - 1)/2; 0 for nx <= 0 (chi+ -> a = 0 -> Lambda = inf)
RT_DEV float ggx_g1(float nx, float alpha2) {
    if (!(nx > 0.0f)) return 0.0f;
    const float nx2 = nx * nx;
    const float tan2 = fmaxf(1.0f - nx2, 0.0f) * fast_rcp(nx2);
    return 2.0f * fast_rcp(1.0f + fast_sqrt(fmaf(alpha2, tan2, 1.0f)));
}
RT_DEV float pow5(float x) { const float x2 = x * x; return x2 * x2 * x; }

struct Material { float3 base; float metallic; float3 emission; float roughness; };

// brdf (rendering.rs:133-155) given the precomputed shared terms.  hl = |h.l|, nl = l.n, nv = v.n.
RT_DEV float3 brdf_eval(const Material& m, float d_chi, float g1l, float g1v, float nl, float nv, float hl) {
    const float g = g1l * g1v;
    const float spec = g > 0.0f ? d_chi * g * fast_rcp(4.0f * nl * nv) : 0.0f;   // rendering.rs:182
    const float w = pow5(1.0f - hl);                                             // fresnel_term rendering.rs:129-131
    const float3 f_metal = m.base + (f3(1.f, 1.f, 1.f) - m.base) * w;
    const float f_diel = fmaf(0.96f, w, 0.04f);
    const float3 metal = f_metal * spec;
    const float3 diel = f3(spec * f_diel, spec * f_diel, spec * f_diel) + m.base * (RT_INV_PI_F * (1.0f - f_diel));
    return metal * m.metallic + diel * (1.0f - m.metallic);
}

// ------------------------------------------------------------------------------------------------ samplers
// CosineWeightedDistribution::sample_unit_vector (distributions.rs:54-63): normalize(uniform_on_sphere + n).
// The reference draws the sphere point from three normals; (u1,u2) -> (z, phi) is the same distribution.
RT_DEV float3 sphere_uniform(float u1, float u2) {
    const float z = fmaf(-2.0f, u1, 1.0f);
    const float r = fast_sqrt(fmaxf(0.0f, fmaf(-z, z, 1.0f)));
    float s, c;
    __sincosf(2.0f * RT_PI_F * u2, &s, &c);
    return f3(r * c, r * s, z);
}
RT_DEV float3 sample_cosine(float3 n, float u1, float u2) { return normalize(sphere_uniform(u1, u2) + n); }

// VndfDistribution::sample_unit_vector (distributions.rs:264-274): a GGX visible normal m for the view direction v,
// then l = reflect(v, m) (geometry.rs:65-69).  The reference draws m with Heitz's 2018 construction (:209-234) in the
// frame built from the fixed helper axis of :265; here the SAME distribution is drawn with the spherical-cap
// construction (Dupuy & Benyoub 2023: a uniform point on the cap z > -Vh.z of the unit sphere plus Vh is a visible
// normal of the stretched configuration) in a branchless tangent frame (Duff et al. 2017).  With the reference's
// isotropic alpha the distribution of m does not depend on the tangent frame, so only the random stream differs --
// which it does anyway (Philox instead of xoshiro).  ~45 instructions instead of ~110.
RT_DEV float3 sample_vndf(float3 n, float3 v, float alpha, float u1, float u2) {
    const float sign = copysignf(1.0f, n.z);
    const float a = -fast_rcp(sign + n.z), b = n.x * n.y * a;
    const float3 t1 = f3(fmaf(sign * n.x, n.x * a, 1.0f), sign * b, -sign * n.x), t2 = f3(b, fmaf(n.y, n.y * a, sign), -n.y);
    const float3 vl = f3(dot(t1, v), dot(t2, v), dot(n, v));
    const float3 Vh = normalize(f3(alpha * vl.x, alpha * vl.y, vl.z));
    float sn, cs;
    __sincosf(2.0f * RT_PI_F * u1, &sn, &cs);
    const float z = fmaf(1.0f - u2, 1.0f + Vh.z, -Vh.z);
    const float r = fast_sqrt(fminf(fmaxf(fmaf(-z, z, 1.0f), 0.0f), 1.0f));
    const float3 Nh = f3(fmaf(r, cs, Vh.x), fmaf(r, sn, Vh.y), z + Vh.z);
    const float3 Ne = normalize(f3(alpha * Nh.x, alpha * Nh.y, fmaxf(0.0f, Nh.z)));
    const float3 m = t1 * Ne.x + t2 * Ne.y + n * Ne.z;
    const float3 l = m * (2.0f * dot(v, m)) - v;                        // reflect_vec geometry.rs:65-69
    return normalize(l);
}

// DirectLightSamplingDistribution::sample_unit_vector, triangle arm (distributions.rs:111-125) for light `idx`
// chosen uniformly by count (MultipleLightSamplingDistribution, :151-158).
template <class Space>
RT_DEV float3 sample_light(const Space& sp, const SceneLayout& L, float3 point, int idx, float u, float v) {
    if (!(u + v < 1.0f)) { u = 1.0f - u; v = 1.0f - v; }
    const uint32_t o16 = (uint32_t)idx * 16u;
    const float3 a = f3(sp.ld4(L.lt_a + o16)), e1 = f3(sp.ld4(L.lt_e1 + o16)), e2 = f3(sp.ld4(L.lt_e2 + o16));
    const float3 p = fma3(e2, v, fma3(e1, u, a));
    return normalize(p - point);
}

// One light triangle's term of MultipleLightSamplingDistribution::pdf (distributions.rs:166-182):
// (1/area) * |p - point|^2 / |ng . omega| for a ray that pierces it at t > 0 (both faces, no occlusion).
template <class Space>
RT_DEV float light_tri_pdf(const Space& sp, const SceneLayout& L, int i, float3 point, float3 l) {
    const uint32_t o = L.lt_t + (uint32_t)i * 64u;
    const float4 r0 = sp.ld4(o);
    TriTest T;
    T.N = f3(r0); T.a = f3(sp.ld4(o + 16u)); T.Au = f3(sp.ld4(o + 32u)); T.Av = f3(sp.ld4(o + 48u));
    float t, u, v;
    if (!tri_test(point, l, T, t, u, v)) return 0.0f;
    return r0.w * t * t * fast_rcp(fabsf(dot(T.N, l)));   // l is unit: |p - point|^2 = t^2, omega = l, N = unit normal
}

// ---- lights that are general primitives (Shape3D::Box arm of distributions.rs:70-148; ellipsoid: own spec) ----------------
// DirectLightSamplingDistribution::sample_unit_vector (distributions.rs:84-125) for light `idx` (chosen uniformly by count,
// :151-158).  Box arm :86-110: a face pair picked by area from x in [0, wx+wy+wz), a random sign, two uniform in-face
// coordinates.  `bits` supplies x (bits 0..30) and the sign (bit 31); the reference spends separate draws (gen_range,
// gen_bool) -- the streams differ anyway.  Ellipsoid (own spec): radii (.) a uniform point of the unit sphere.
template <class Space>
RT_DEV float3 sample_light_gen(const Space& sp, const SceneLayout& L, float3 point, int idx, float u1, float u2, uint32_t bits) {
    const PrimRec R = load_prim(sp, L.lt_g, idx);
    const int kind = prim_kind(R);
    if (kind == RT_KIND_TRIANGLE) return sample_light(sp, L, point, idx, u1, u2);
    const float3 s = prim_s(R);
    float3 pl;
    if (kind == RT_KIND_BOX) {
        const float wx = s.y * s.z, wy = s.x * s.z, wz = s.x * s.y;                  // the common factor 4 of :87 cancels
        const float x = u01(bits << 1) * (wx + wy + wz);
        const float sg = (bits >> 31) ? 1.0f : -1.0f;
        const float c1 = fmaf(2.0f, u1, -1.0f), c2 = fmaf(2.0f, u2, -1.0f);         // gen_range(-s..s) = (2u - 1) s
        if (x < wx) pl = f3(s.x * sg, c1 * s.y, c2 * s.z);
        else if (x < wx + wy) pl = f3(c1 * s.x, s.y * sg, c2 * s.z);
        else pl = f3(c1 * s.x, c2 * s.y, s.z * sg);
    } else {
        pl = sphere_uniform(u1, u2) * s;
    }
    return normalize(prim_to_world(R, pl) + f3(R.r3) - point);                       // :121-124
}

// One light's term of MultipleLightSamplingDistribution::pdf (distributions.rs:166-182) for a general primitive: EVERY hit of
// the ray with the light (entry and exit of a box: intersect_ray_with_object3d_all_points, geometry.rs:226-251) adds
// local_pdf * |p - point|^2 / |ng . omega|; l is unit, so |p - point|^2 = t^2 and, R being orthonormal, ng . omega is
// evaluated in object space.  local_pdf: get_local_pdf (:70-81): box 1 / (8 (sx sy + sy sz + sz sx)); ellipsoid (own spec)
// 1 / (4 pi sqrt((ux ry rz)^2 + (rx uy rz)^2 + (rx ry uz)^2)) at the unit-sphere point u = p / r.
template <class Space>
RT_DEV float light_prim_pdf(const Space& sp, const SceneLayout& L, int i, float3 point, float3 l) {
    const PrimRec R = load_prim(sp, L.lt_g, i);
    const int kind = prim_kind(R);
    if (kind == RT_KIND_TRIANGLE) {
        TriTest T; T.N = f3(R.r0); T.a = f3(R.r1); T.Au = f3(R.r2); T.Av = f3(R.r3);
        float t, u, v;
        if (!tri_test(point, l, T, t, u, v)) return 0.0f;
        return R.r1.w * t * t * fast_rcp(fabsf(dot(T.N, l)));
    }
    const float3 s = prim_s(R);
    const float3 ol = prim_to_local(R, point - f3(R.r3)), dl = prim_to_local(R, l);
    if (kind == RT_KIND_BOX) {
        const float3 inv = safe_inv_dir(dl);
        const float ax = (-s.x - ol.x) * inv.x, bx = (s.x - ol.x) * inv.x;
        const float ay = (-s.y - ol.y) * inv.y, by = (s.y - ol.y) * inv.y;
        const float az = (-s.z - ol.z) * inv.z, bz = (s.z - ol.z) * inv.z;
        const float nx = fminf(ax, bx), fx = fmaxf(ax, bx), ny = fminf(ay, by), fy = fmaxf(ay, by), nz = fminf(az, bz), fz = fmaxf(az, bz);
        const float tmin = max3f(nx, ny, nz), tmax = min3f(fx, fy, fz);
        if (!(tmin <= tmax) || !(tmax > 0.0f)) return 0.0f;
        const float local_pdf = fast_rcp(8.0f * (s.x * s.y + s.y * s.z + s.z * s.x));
        const int an = box_face(s, ol, dl, tmin, (nx >= ny && nx >= nz) ? 0 : (ny >= nz ? 1 : 2));
        const int af = box_face(s, ol, dl, tmax, (fx <= fy && fx <= fz) ? 0 : (fy <= fz ? 1 : 2));
        const float dn = fabsf(an == 0 ? dl.x : (an == 1 ? dl.y : dl.z));
        const float df = fabsf(af == 0 ? dl.x : (af == 1 ? dl.y : dl.z));
        float sum = tmax * tmax * fast_rcp(df);
        if (tmin > 0.0f) sum += tmin * tmin * fast_rcp(dn);
        return local_pdf * sum;
    }
    if (kind != RT_KIND_ELLIPSOID) return 0.0f;                                       // planes are never lights
    const float3 ir = f3(fast_rcp(s.x), fast_rcp(s.y), fast_rcp(s.z));
    const float3 od = ol * ir, dd = dl * ir;
    const float a = dot(dd, dd), hb = dot(od, dd), c = dot(od, od) - 1.0f;
    const float ia = fast_rcp(a);
    const float tc = -hb * ia;
    const float3 pc = fma3(dd, tc, od);
    const float h2 = (1.0f - dot(pc, pc)) * ia;
    if (!(h2 >= 0.0f)) return 0.0f;
    const float h = fast_sqrt(h2);
    const bool inside = c < 0.0f;
    if (!inside && !(hb < 0.0f)) return 0.0f;
    const float3 rr = f3(s.y * s.z, s.x * s.z, s.x * s.y);
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float t = k == 0 ? tc - h : tc + h;
        if (k == 0 && inside) continue;
        if (!(t > 0.0f)) continue;
        const float3 us = fma3(dd, t, od);                                            // unit-sphere point
        const float3 j = us * rr;
        const float local_pdf = fast_rcp(4.0f * RT_PI_F) * fast_rsqrt(dot(j, j));
        const float3 nl = normalize(us * ir);                                         // object-space normal p / r^2
        sum += local_pdf * t * t * fast_rcp(fabsf(dot(nl, dl)));
    }
    return sum;
}

// MultipleLightSamplingDistribution::pdf (distributions.rs:160-184): sum over ALL lights the ray
// pierces, divided by the light count.  Few lights: plain loop (the reference's light BVH is a single leaf for
// every shipped scene).  Many lights: all-hits walk of the light BVH = intersect_with_bvh_all_points
// (bvh.rs:174-229): no pruning by distance, every intersected leaf primitive contributes.  The walk pushes its own
// RT_CUR_DONE marker first, so it can run on top of the traversal stack of a ray that is still in flight.
// GEN: the lights are general primitives (light_prim_pdf) instead of triangle TEST records (light_tri_pdf).
template <class Space, bool STATS, bool GEN = false>
RT_DEV float light_pdf(const Space& sp, const SceneLayout& L, SmemStack& st, float3 point, float3 l, Counters& cnt) {
    float sum = 0.0f;
    if (!L.light_bvh) {
        for (int i = 0; i < L.n_lights; ++i) sum += GEN ? light_prim_pdf(sp, L, i, point, l) : light_tri_pdf(sp, L, i, point, l);
        if (STATS) cnt.light_tri_tests += (unsigned long long)L.n_lights;
    } else {
        const RaySetup r = ray_setup(point, l, L.lnodes);
        st.push(RT_CUR_DONE);                    // marker: in the wavefront kernel the lane's own ray keeps its entries below
        int cur = 0;
        for (;;) {
            while (cur >= 0) pair_step<Space, false>(sp, L.lnodes, r, RT_INF_F, cur, st);
            if (cur == RT_CUR_DONE) break;
            const uint32_t code = (uint32_t)~cur;
            const int first = (int)(code >> 3), n = (int)(code & 7u) + 1;
            for (int i = first; i < first + n; ++i) sum += GEN ? light_prim_pdf(sp, L, i, point, l) : light_tri_pdf(sp, L, i, point, l);
            if (STATS) cnt.light_tri_tests += (unsigned long long)n;
            cur = st.pop();
        }
    }
    return sum * L.inv_n_lights;
}

// pdf of the cosine component (distributions.rs:65-67) and of the VNDF component (:276-297) for unit l.
RT_DEV float pdf_cosine(float nl) { return fmaxf(0.0f, nl) * RT_INV_PI_F; }
// Dv(ni,v)/(4 v.ni) = G1(v) max(0,v.ni) Dn(ni) / (v.z 4 v.ni) = G1(v) Dn(ni) / (4 n.v)   (v.ni >= 0 always)
RT_DEV float pdf_vndf(float d_nochi, float g1v, float nv) { return g1v * d_nochi * fast_rcp(4.0f * nv); }

}  // namespace rtd
