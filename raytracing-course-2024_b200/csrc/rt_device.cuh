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
RT_DEV float3 normalize(float3 a) { return a * rsqrtf(dot(a, a)); }
RT_DEV float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
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
    uint32_t box_a, box_b, box_c, child;         // finite-primitive BVH, child-pair nodes (host_scene.h)
    uint32_t tri_a, tri_e1, tri_e2;              // BVH-ordered triangles: a, b-a, c-a (w unused)
    uint32_t sh_n0, sh_dn1, sh_dn2, sh_ng;       // a_norm|material id, b_norm-a_norm|orig id, c_norm-a_norm, unit face normal
    uint32_t mat0, mat1;                         // base rgb|metallic, emission rgb|roughness
    uint32_t lt_a, lt_e1, lt_e2, lt_ng;          // light triangles (light-BVH order): a|1/area, e1, e2, unit normal
    uint32_t lbox_a, lbox_b, lbox_c, lchild;     // light BVH (used when n_lights > RT_BRUTE_LIGHTS)
    uint32_t total_bytes;
    int32_t n_nodes, n_tris, n_mats, n_lights, n_lnodes;
    int32_t light_bvh;                           // 1: pdf walks the light BVH, 0: loops over all lights
};
#define RT_BRUTE_LIGHTS 8

struct SmemSpace {
    uint32_t base;  // shared-window address of the blob
    RT_DEV float4 ld4(uint32_t off) const {
        float4 v;
        asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + off));
        return v;
    }
    RT_DEV int2 ld2i(uint32_t off) const {
        int2 v;
        asm("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(base + off));
        return v;
    }
};
struct GmemSpace {
    const char* base;
    RT_DEV float4 ld4(uint32_t off) const { return __ldg(reinterpret_cast<const float4*>(base + off)); }
    RT_DEV int2 ld2i(uint32_t off) const { return __ldg(reinterpret_cast<const int2*>(base + off)); }
};

// Per-thread traversal stack in shared memory, laid out [entry][thread]: bank = thread % 32 for every entry, so
// pushes and pops are conflict-free however far the lanes' stack depths have diverged.
struct SmemStack {
    uint32_t addr;    // shared-window address of entry 0 of this thread
    uint32_t stride;  // bytes between entries = 4 * blockDim.x
    RT_DEV void store(int i, int v) const { asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr + (uint32_t)i * stride), "r"(v)); }
    RT_DEV int load(int i) const { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr + (uint32_t)i * stride)); return v; }
};

struct Counters { unsigned long long node_tests, tri_tests, light_tri_tests; };

// ------------------------------------------------------------------------------------------------ intersection
// Conservative slab test of the two child boxes of a pair node.  Same accept/prune rule as
// get_aabb_intersection + bvh.rs:258-263: a box is entered iff the ray's [max(t_entry,0), t_exit] interval is
// non-empty and t_entry <= best (prune iff best < t_entry while outside).  The reference perturbs the
// direction by 1e-8 (geometry.rs:144-155); here zero components are replaced by +-1e-20 before inversion.
RT_DEV float3 safe_inv_dir(float3 d) {
    const float tiny = 1e-20f;
    float x = fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x;
    float y = fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y;
    float z = fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z;
    return f3(fast_rcp(x), fast_rcp(y), fast_rcp(z));
}
RT_DEV bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz, float3 inv, float3 od, float t_best, float& t_entry) {
    float x0 = fmaf(lox, inv.x, -od.x), x1 = fmaf(hix, inv.x, -od.x);
    float y0 = fmaf(loy, inv.y, -od.y), y1 = fmaf(hiy, inv.y, -od.y);
    float z0 = fmaf(loz, inv.z, -od.z), z1 = fmaf(hiz, inv.z, -od.z);
    float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    float tmax = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_best));
    t_entry = tmin;
    // boxes are EPS-padded (1e-5, aabb.rs:53-65) and rounded outward: ~20 ulps of slack for the FP32 slab arithmetic
    return tmin <= tmax;
}

// Two-sided ray/triangle test with inclusive edges: u >= 0, v >= 0, u + v <= 1, t > 0 (geometry.rs:109-113).
// The reference solves the 3x3 system with Matrix3::try_inverse; this is the same solution by Cramer's rule on
// precomputed edges (Moller-Trumbore form).  det == 0 -> inf/NaN -> every comparison fails, like try_inverse.
RT_DEV bool tri_test(float3 o, float3 d, float3 a, float3 e1, float3 e2, float& t, float& u, float& v) {
    float3 pvec = cross(d, e2);
    float det = dot(e1, pvec);
    float inv_det = fast_rcp(det);
    float3 tvec = o - a;
    float3 qvec = cross(tvec, e1);
    u = dot(tvec, pvec) * inv_det;
    v = dot(d, qvec) * inv_det;
    t = dot(e2, qvec) * inv_det;
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
template <class Space, bool STATS>
RT_DEV void trace_nearest(const Space& sp, const SceneLayout& L, const SmemStack& st, float3 o, float3 d, int skip_tri, Hit& hit, Counters& cnt) {
    const float3 inv = safe_inv_dir(d);
    const float3 od = o * inv;
    hit.t = RT_INF_F; hit.u = 0.f; hit.v = 0.f; hit.tri = -1;
    int cur = 0, sptr = 0;
    for (;;) {
        bool done = false;
        while (cur >= 0) {
            const uint32_t o16 = (uint32_t)cur * 16u;
            const float4 A = sp.ld4(L.box_a + o16), B = sp.ld4(L.box_b + o16), C = sp.ld4(L.box_c + o16);
            const int2 ch = sp.ld2i(L.child + (uint32_t)cur * 8u);
            float t0, t1;
            const bool h0 = slab(A.x, A.y, A.z, A.w, C.x, C.y, inv, od, hit.t, t0);
            const bool h1 = slab(B.x, B.y, B.z, B.w, C.z, C.w, inv, od, hit.t, t1);
            if (STATS) cnt.node_tests += 2;
            if (h0 & h1) {
                const bool swap = t1 < t0;
                st.store(sptr++, swap ? ch.x : ch.y);
                cur = swap ? ch.y : ch.x;
            } else if (h0 | h1) {
                cur = h0 ? ch.x : ch.y;
            } else {
                if (sptr == 0) { done = true; break; }
                cur = st.load(--sptr);
            }
        }
        if (done) break;
        const uint32_t code = (uint32_t)~cur;
        const int first = (int)(code >> 3), n = (int)(code & 7u) + 1;
        for (int i = first; i < first + n; ++i) {
            const uint32_t o16 = (uint32_t)i * 16u;
            const float4 a = sp.ld4(L.tri_a + o16), e1 = sp.ld4(L.tri_e1 + o16), e2 = sp.ld4(L.tri_e2 + o16);
            float t, u, v;
            const bool ok = tri_test(o, d, f3(a), f3(e1), f3(e2), t, u, v);
            if (STATS) cnt.tri_tests += 1;
            if (ok && t < hit.t && i != skip_tri) { hit.t = t; hit.u = u; hit.v = v; hit.tri = i; }
        }
        if (sptr == 0) break;
        cur = st.load(--sptr);
    }
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
    return 2.0f * fast_rcp(1.0f + sqrtf(fmaf(alpha2, tan2, 1.0f)));
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
    const float r = sqrtf(fmaxf(0.0f, fmaf(-z, z, 1.0f)));
    float s, c;
    __sincosf(2.0f * RT_PI_F * u2, &s, &c);
    return f3(r * c, r * s, z);
}
RT_DEV float3 sample_cosine(float3 n, float u1, float u2) { return normalize(sphere_uniform(u1, u2) + n); }

// VndfDistribution::sample_unit_vector (distributions.rs:264-274) incl. the fixed helper axis of :265 and
// sample_ggx_vndf (:209-234, Heitz 2018).
RT_DEV float3 sample_vndf(float3 n, float3 v, float alpha, float u1, float u2) {
    const float3 helper = f3(0.23120537f, 0.12192488f, 0.96517408f);   // normalize(0.234, 0.1234, 0.97686)
    const float3 t1 = normalize(cross(n, helper));
    const float3 t2 = normalize(cross(n, t1));
    const float3 vl = f3(dot(t1, v), dot(t2, v), dot(n, v));
    const float3 Vh = normalize(f3(alpha * vl.x, alpha * vl.y, vl.z));
    const float lensq = fmaf(Vh.x, Vh.x, Vh.y * Vh.y);
    const float3 T1 = lensq > 0.0f ? f3(-Vh.y, Vh.x, 0.0f) * rsqrtf(lensq) : f3(1.0f, 0.0f, 0.0f);
    const float3 T2 = cross(Vh, T1);
    const float r = sqrtf(u1);
    float sn, cs;
    __sincosf(2.0f * RT_PI_F * u2, &sn, &cs);
    const float a1 = r * cs;
    float a2 = r * sn;
    const float s = 0.5f * (1.0f + Vh.z);
    a2 = fmaf(1.0f - s, sqrtf(fmaxf(0.0f, fmaf(-a1, a1, 1.0f))), s * a2);
    const float3 Nh = T1 * a1 + T2 * a2 + Vh * sqrtf(fmaxf(0.0f, 1.0f - a1 * a1 - a2 * a2));
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
    const uint32_t o16 = (uint32_t)i * 16u;
    const float4 a = sp.ld4(L.lt_a + o16);
    const float3 e1 = f3(sp.ld4(L.lt_e1 + o16)), e2 = f3(sp.ld4(L.lt_e2 + o16));
    float t, u, v;
    if (!tri_test(point, l, f3(a), e1, e2, t, u, v)) return 0.0f;
    const float3 ng = f3(sp.ld4(L.lt_ng + o16));
    return a.w * t * t * fast_rcp(fabsf(dot(ng, l)));   // l is unit: |p - point|^2 = t^2, omega = l
}

// MultipleLightSamplingDistribution::pdf (distributions.rs:160-184): sum over ALL light triangles the ray
// pierces, divided by the light count.  Few lights: plain loop (the reference's light BVH is a single leaf for
// every shipped scene).  Many lights: all-hits walk of the light BVH = intersect_with_bvh_all_points
// (bvh.rs:174-229): no pruning by distance, every intersected leaf triangle contributes.
template <class Space, bool STATS>
RT_DEV float light_pdf(const Space& sp, const SceneLayout& L, const SmemStack& st, float3 point, float3 l, Counters& cnt) {
    float sum = 0.0f;
    if (!L.light_bvh) {
        for (int i = 0; i < L.n_lights; ++i) sum += light_tri_pdf(sp, L, i, point, l);
        if (STATS) cnt.light_tri_tests += (unsigned long long)L.n_lights;
    } else {
        const float3 inv = safe_inv_dir(l);
        const float3 od = point * inv;
        int cur = 0, sptr = 0;
        for (;;) {
            bool done = false;
            while (cur >= 0) {
                const uint32_t o16 = (uint32_t)cur * 16u;
                const float4 A = sp.ld4(L.lbox_a + o16), B = sp.ld4(L.lbox_b + o16), C = sp.ld4(L.lbox_c + o16);
                const int2 ch = sp.ld2i(L.lchild + (uint32_t)cur * 8u);
                float t0, t1;
                const bool h0 = slab(A.x, A.y, A.z, A.w, C.x, C.y, inv, od, RT_INF_F, t0);
                const bool h1 = slab(B.x, B.y, B.z, B.w, C.z, C.w, inv, od, RT_INF_F, t1);
                if (h0 & h1) { st.store(sptr++, ch.y); cur = ch.x; }
                else if (h0 | h1) cur = h0 ? ch.x : ch.y;
                else { if (sptr == 0) { done = true; break; } cur = st.load(--sptr); }
            }
            if (done) break;
            const uint32_t code = (uint32_t)~cur;
            const int first = (int)(code >> 3), n = (int)(code & 7u) + 1;
            for (int i = first; i < first + n; ++i) sum += light_tri_pdf(sp, L, i, point, l);
            if (STATS) cnt.light_tri_tests += (unsigned long long)n;
            if (sptr == 0) break;
            cur = st.load(--sptr);
        }
    }
    return sum / (float)L.n_lights;
}

// pdf of the cosine component (distributions.rs:65-67) and of the VNDF component (:276-297) for unit l.
RT_DEV float pdf_cosine(float nl) { return fmaxf(0.0f, nl) * RT_INV_PI_F; }
// Dv(ni,v)/(4 v.ni) = G1(v) max(0,v.ni) Dn(ni) / (v.z 4 v.ni) = G1(v) Dn(ni) / (4 n.v)   (v.ni >= 0 always)
RT_DEV float pdf_vndf(float d_nochi, float g1v, float nv) { return g1v * d_nochi * fast_rcp(4.0f * nv); }

}  // namespace rtd
