// rt_kernels.cu -- the sm_100a kernels that replace the pixel loop of render_scene (rendering.rs:21-69): the shading
// helpers shared by both render kernels, v1 (`render_kernel`, the per-lane megakernel kept as the A/B baseline), v3
// (`render_wave_kernel`, rt_render_wave.cuh: the product path), plus the nearest-hit query kernel, the resolve
// (color_to_pixel) kernels, the unit-function kernel used by the parity tests and the FFMA issue micro-benchmark.
//
// v1 design (DESIGN.md has the full story, rt_render_wave.cuh the v3 one):
//   * one persistent grid, CTAs = SMs x occupancy; every LANE owns one (pixel, sample-chunk) work item at a time
//     and pulls the next one from a global counter (warp-aggregated atomic) the moment it runs out of samples, so
//     there is no per-warp tail; inside an item the lane regenerates a camera path as soon as its path ends;
//   * one loop iteration = one path segment: nearest-hit BVH traversal, emission, the reference's rejection loop
//     over the 3-component mixture (cosine / GGX-VNDF / lights), BRDF, throughput update;
//   * the scene blob lives in shared memory when it fits, traversal stacks always do ([entry][thread]);
//   * RNG: Philox4x32-10, counter (pixel, sample, call#) -> the image is independent of the sharding.
#include "rt_kernels.h"

#include <cstring>

#include "../../include/rt_api.h"

namespace rtd {

#ifndef RT_BLOCK
#define RT_BLOCK 256
#endif
#ifndef RT_MIN_BLOCKS
#define RT_MIN_BLOCKS 2
#endif

// ------------------------------------------------------------------------------------------------ shading helpers
struct DirTerms { float nl, hl, d_nochi, nh; };

// Per-direction quantities shared by the three pdfs and the BRDF: h = normalize(l+v) (rendering.rs:136,
// distributions.rs:290), n.l, |h.l|, n.h and the GGX D without chi+.  l+v == 0 gives NaN in the reference
// (sample rejected by `pdf > 0`); here d_nochi becomes NaN too and the same test rejects it.
RT_DEV DirTerms dir_terms(float3 n, float3 v, float3 l, float alpha2) {
    DirTerms t;
    const float3 h = normalize(l + v);
    t.nl = dot(n, l);
    t.hl = fabsf(dot(h, l));
    t.nh = dot(n, h);
    t.d_nochi = ggx_d_nochi(n, h, alpha2);
    return t;
}

template <class Space>
RT_DEV Material load_material(const Space& sp, const SceneLayout& L, int id) {
    const float4 m0 = sp.ld4(L.mat0 + (uint32_t)id * 16u), m1 = sp.ld4(L.mat1 + (uint32_t)id * 16u);
    Material m;
    m.base = f3(m0); m.metallic = m0.w; m.emission = f3(m1); m.roughness = m1.w;
    return m;
}

// What get_ray_color (rendering.rs:96-101,107) takes from a hit: corrected_point = o + d (t - EPS) (:98), normal_geometry
// (sampling, pdfs, BRDF, cosine) and normal_shading (acceptance gate only) both facing the incoming ray, is_outer_to_inner.
// Triangles: the point comes from the barycentrics (on the triangle's plane to ~1 ulp), the shading normal is interpolated and
// left un-normalised (only its sign test is used).  GEN scenes dispatch on the primitive record (prim_frame); `flat` says that a
// ray leaving the primitive can never meet it again (triangles, planes), which the self-hit skip relies on.
struct Vertex { float3 P, surf, n, ns; bool outer, flat; };
template <class Space, bool GEN>
RT_DEV Vertex hit_vertex(const Space& sp, const SceneLayout& L, int tri, float4 n0, float3 o, float3 din, float hu, float hv) {
    Vertex V;
    const uint32_t o16 = (uint32_t)tri * 16u;
    if (GEN) {
        const PrimRec R = load_prim(sp, L.prims, tri);
        const int kind = prim_kind(R);
        if (kind != RT_KIND_TRIANGLE) {
            const HitFrame F = prim_frame(R, o, din, hu, __float_as_int(hv));
            V.n = F.n; V.ns = F.ns; V.outer = F.outer; V.flat = kind == RT_KIND_PLANE;
            V.surf = fma3(din, hu, o);
            V.P = fma3(din, -RT_EPS_F, V.surf);
            return V;
        }
    }
    const float3 ng = f3(sp.ld4(L.sh_ng + o16));
    V.outer = dot(ng, din) < 0.0f;
    const float sgn = V.outer ? 1.0f : -1.0f;                    // geometry.rs:115-126
    V.n = ng * sgn;
    const float3 dn1 = f3(sp.ld4(L.sh_dn1 + o16)), dn2 = f3(sp.ld4(L.sh_dn2 + o16));
    V.ns = (f3(n0) + dn1 * hu + dn2 * hv) * sgn;
    const float3 ta = f3(sp.ld4(L.tri_a + o16)), te1 = f3(sp.ld4(L.tri_e1 + o16)), te2 = f3(sp.ld4(L.tri_e2 + o16));
    V.surf = fma3(te2, hv, fma3(te1, hu, ta));
    V.P = fma3(din, -RT_EPS_F, V.surf);                          // :98
    V.flat = true;
    return V;
}

// [OWN SPEC, DESIGN.md section 12 -- reference HEAD has no transmissive material, only the dead fields `ior` (scene.rs:18) and
// `is_outer_to_inner` (geometry.rs:23)]  Smooth dielectric, delta BSDF: eta = n_from / n_to with the outside index 1, Schlick
// reflectance, total internal reflection reflects, ONE uniform picks reflection (u < R) or refraction; the refracted ray
// starts EPS BEHIND the surface and is tinted by the base colour when it ENTERS the solid.  Returns true for refraction.
RT_DEV bool dielectric_sample(float3 n, float3 v, float nv, float ior, bool outer, float u, float3& l) {
    const float eta = outer ? fast_rcp(ior) : ior;
    const float sin2sq = eta * eta * fmaxf(0.0f, fmaf(-nv, nv, 1.0f));
    bool refract = false;
    float cos2 = 0.0f;
    if (sin2sq < 1.0f) {
        cos2 = fast_sqrt(1.0f - sin2sq);
        const float q = (eta - 1.0f) * fast_rcp(eta + 1.0f), r0 = q * q;
        const float refl = fmaf(1.0f - r0, pow5(1.0f - nv), r0);
        refract = !(u < refl);
    }
    l = refract ? normalize(n * fmaf(eta, nv, -cos2) - v * eta) : normalize(n * (2.0f * nv) - v);   // geometry.rs:65-69 for the mirror arm
    return refract;
}

// get_ray_to_pixel (rendering.rs:71-84) with explicit jitter.
RT_DEV void camera_ray(const Camera& c, int W, int H, int x, int y, float xi1, float xi2, float3& o, float3& d) {
    const float rx = (float)x + xi1, ry = (float)y + xi2;
    const float px = fmaf(rx, c.two_over_w, -1.0f) * c.tan_x;
    const float py = -fmaf(ry, c.two_over_h, -1.0f) * c.tan_y;
    const float3 dir = f3(c.right[0], c.right[1], c.right[2]) * px + f3(c.up[0], c.up[1], c.up[2]) * py + f3(c.fwd[0], c.fwd[1], c.fwd[2]);
    o = f3(c.pos[0], c.pos[1], c.pos[2]);
    d = normalize(dir);
}

// One iteration of the rejection loop rendering.rs:102-110: MixDistribution::sample_unit_vector
// (distributions.rs:188-192) followed by MixDistribution::pdf (:194-201).  Returns the mixture pdf; `terms` are
// the per-direction quantities for the BRDF of the accepted direction.
template <class Space, bool STATS, bool GEN = false>
RT_DEV float mix_sample_and_pdf(const Space& sp, const SceneLayout& L, SmemStack& st, int n_comp, float inv_n_comp, float3 P, float3 n, float3 v,
                                float nv, float alpha, float alpha2, float g1v, uint4 rnd, float3& l, DirTerms& terms, Counters& cnt) {
    const uint32_t comp = __umulhi(rnd.x, (uint32_t)n_comp);           // gen_range(0..len)
    const float u1 = u01(rnd.y), u2 = u01(rnd.z);
    if (comp == 0u) l = sample_cosine(n, u1, u2);
    else if (comp == 1u) l = sample_vndf(n, v, alpha, u1, u2);
    else if (!GEN) l = sample_light(sp, L, P, (int)__umulhi(rnd.w, (uint32_t)L.n_lights), u1, u2);
    else l = sample_light_gen(sp, L, P, (int)__umulhi(rnd.w, (uint32_t)L.n_lights), u1, u2, rnd.w * (uint32_t)L.n_lights);   // low product word: uniform given the index
    terms = dir_terms(n, v, l, alpha2);
    float pdf = pdf_cosine(terms.nl) + pdf_vndf(terms.d_nochi, g1v, nv);
    if (n_comp == 3) pdf += light_pdf<Space, STATS, GEN>(sp, L, st, P, l, cnt);
    return pdf * inv_n_comp;
}

template <class Space> struct SpaceMaker;
template <> struct SpaceMaker<SmemSpace> { static RT_DEV SmemSpace make(const char*, uint32_t smem_addr) { SmemSpace s; s.base = smem_addr; return s; } };
template <> struct SpaceMaker<GmemSpace> { static RT_DEV GmemSpace make(const char* g, uint32_t) { GmemSpace s; s.base = g; return s; } };
template <class Space> struct IsSmem { static const bool value = false; };
template <> struct IsSmem<SmemSpace> { static const bool value = true; };

// ------------------------------------------------------------------------------------------------ render kernel
template <class Space, bool STATS>
__global__ void __launch_bounds__(RT_BLOCK, RT_MIN_BLOCKS) render_kernel(const RenderArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t stack_bytes = a.stack_entries * blockDim.x * 4u;
    if (IsSmem<Space>::value) {
        const uint4* src = reinterpret_cast<const uint4*>(a.blob);
        uint4* dst = reinterpret_cast<uint4*>(smem_raw + stack_bytes);
        for (uint32_t i = threadIdx.x; i < a.L.total_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    Space sp = SpaceMaker<Space>::make(a.blob, smem_base + stack_bytes);
#ifdef RT_DEBUG_BOUNDS
    sp.limit = a.debug_blob_limit ? a.debug_blob_limit : a.L.total_bytes;
#endif
    const SceneLayout& L = a.L;
    SmemStack st; st.init(smem_base + threadIdx.x * 4u, blockDim.x * 4u, a.stack_entries);
    const unsigned lane = threadIdx.x & 31u;
    const uint2 key = make_uint2(a.seed_lo, a.seed_hi);
    const float3 bg = f3(a.bg[0], a.bg[1], a.bg[2]);
    const size_t n_pix = (size_t)a.W * (size_t)a.H;

    // work-item state
    bool have_item = false, exhausted = false, alive = false;
    uint32_t pix = 0, chunk = 0;
    int px = 0, py = 0, s_cur = 0, s_stop = 0;
    float3 acc = f3(0.f, 0.f, 0.f);
    float acc_n = 0.f;
    // path state
    float3 o = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 1.f), T = f3(1.f, 1.f, 1.f), Ls = f3(0.f, 0.f, 0.f);
    int depth_left = 0, skip_tri = -1, s_this = 0;
    uint32_t call = 0;

    Counters cnt; cnt.node_tests = 0; cnt.tri_tests = 0; cnt.light_tri_tests = 0;
    unsigned long long c_samples = 0, c_segments = 0, c_vertices = 0, c_attempts = 0, c_cap = 0, c_nonfinite = 0;

    for (;;) {
        if (!alive && have_item && s_cur >= s_stop) {           // item done: one store per (pixel, chunk)
            RT_BOUNDS(chunk < (uint32_t)a.n_chunks, RT_BOUNDS_CHUNK);
            RT_BOUNDS((size_t)pix < n_pix, RT_BOUNDS_LAYER);
            a.layers[(size_t)chunk * n_pix + pix] = make_float4(acc.x, acc.y, acc.z, acc_n);
            have_item = false;
        }
        for (;;) {                                               // warp-aggregated work fetch
            const bool need = !alive && !have_item && !exhausted;
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (m == 0u) break;
            const int leader = __ffs(m) - 1;
            unsigned int base = 0;
            if ((int)lane == leader) base = atomicAdd(a.work_counter, (unsigned int)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (need) {
                const uint32_t idx = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
                if (idx >= a.total_items) exhausted = true;
                else {
                    chunk = idx / a.n_pix_items;
                    const uint32_t r = idx - chunk * a.n_pix_items;
                    const uint32_t tile = r >> 5, w = r & 31u;
                    const uint32_t ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
                    px = (int)(tx * 8u + (w & 7u)); py = (int)(ty * 4u + (w >> 3));
                    if (px < a.W && py < a.H && tile % a.shard_count == a.shard_index) {
                        have_item = true;
                        pix = (uint32_t)py * (uint32_t)a.W + (uint32_t)px;
                        s_cur = a.s_begin + __ldg(a.chunk_begin + chunk);
                        s_stop = a.s_begin + __ldg(a.chunk_begin + chunk + 1);
                        acc = f3(0.f, 0.f, 0.f); acc_n = 0.f;
                    }
                }
            }
        }
        if (!alive && have_item) {                               // regenerate: next camera path of this item
            s_this = s_cur++;
            const uint4 r = philox4x32_10(make_uint4(pix, (uint32_t)s_this, 0u, RT_PHILOX_TAG), key);
            camera_ray(a.cam, a.W, a.H, px, py, u01(r.x), u01(r.y), o, d);
            T = f3(1.f, 1.f, 1.f); Ls = f3(0.f, 0.f, 0.f);
            depth_left = a.ray_depth; skip_tri = -1; call = 1u; alive = true;
            if (STATS) ++c_samples;
        }
        if (!__any_sync(0xffffffffu, alive)) break;
        if (alive) {

        // ---- one path segment: get_ray_color (rendering.rs:86-127), iteratively
        Hit hit;
        trace_nearest<Space, STATS>(sp, L, st, o, d, skip_tri, hit, cnt);           // :96 intersect_ray_with_scene
        if (STATS) ++c_segments;
        bool end_path = false;
        if (hit.tri < 0) {
            Ls = Ls + T * bg;                                                        // :125
            end_path = true;
        } else {
            const uint32_t o16 = (uint32_t)hit.tri * 16u;
            const float4 n0 = sp.ld4(L.sh_n0 + o16);
            const Material mat = load_material(sp, L, __float_as_int(n0.w));
            Ls = Ls + T * mat.emission;                                              // :99
            if (--depth_left <= 0) {
                end_path = true;                                                     // :93-95 (child returns 0)
            } else {
                if (STATS) ++c_vertices;
                const float3 ng = f3(sp.ld4(L.sh_ng + o16));
                const float sgn = dot(ng, d) < 0.0f ? 1.0f : -1.0f;                  // geometry.rs:115-126
                const float3 n = ng * sgn;
                const float3 dn1 = f3(sp.ld4(L.sh_dn1 + o16)), dn2 = f3(sp.ld4(L.sh_dn2 + o16));
                const float3 ns = (f3(n0) + dn1 * hit.u + dn2 * hit.v) * sgn;        // un-normalised: only its sign test is used (:107)
                const float3 ta = f3(sp.ld4(L.tri_a + o16)), te1 = f3(sp.ld4(L.tri_e1 + o16)), te2 = f3(sp.ld4(L.tri_e2 + o16));
                // corrected_point = origin + direction * (t - EPS) (:98), evaluated from the barycentrics so that
                // the FP32 point lies on the triangle's plane to ~1 ulp before the back-off
                const float3 P = fma3(d, -RT_EPS_F, fma3(te2, hit.v, fma3(te1, hit.u, ta)));
                const float3 v = -d;                                                 // :101
                const float nv = dot(n, v);
                const float alpha = mat.roughness * mat.roughness, alpha2 = alpha * alpha;   // :158
                const float g1v = ggx_g1(nv, alpha2);
                float3 l = f3(0.f, 0.f, 1.f);
                DirTerms terms; terms.nl = 0.f; terms.hl = 0.f; terms.d_nochi = 0.f; terms.nh = 0.f;
                float pdf = 0.0f;
                bool accepted = false;
                for (int attempt = 0; attempt < a.max_attempts; ++attempt) {        // :102-110
                    const uint4 rnd = philox4x32_10(make_uint4(pix, (uint32_t)s_this, call++, RT_PHILOX_TAG), key);
                    pdf = mix_sample_and_pdf<Space, STATS>(sp, L, st, a.n_comp, a.inv_n_comp, P, n, v, nv, alpha, alpha2, g1v, rnd, l, terms, cnt);
                    if (STATS) ++c_attempts;
                    if (pdf > 0.0f && dot(l, ns) > 0.0f) { accepted = true; break; }
                }
                if (!accepted) {
                    if (STATS) ++c_cap;
                    end_path = true;
                } else {
                    const float d_chi = terms.nh > 0.0f ? terms.d_nochi : 0.0f;      // chi_plus(hn) :164
                    const float3 f = brdf_eval(mat, d_chi, ggx_g1(terms.nl, alpha2), g1v, terms.nl, nv, terms.hl);   // :121
                    T = T * f * (terms.nl * fast_rcp(pdf));                          // :122
                    if (!finite3(T)) { if (STATS) ++c_nonfinite; end_path = true; }  // same policy as the wavefront kernel: cut here, keep what was gathered
                    o = P; d = l;
                    skip_tri = terms.nl > 0.0f ? hit.tri : -1;
                }
            }
        }
        if (end_path) {
            acc = acc + Ls;
            acc_n += 1.0f;
            alive = false;
        }
        }  // if (alive)
    }
    if (STATS) {
        atomicAdd(a.stats + RT_STAT_SAMPLES, c_samples); atomicAdd(a.stats + RT_STAT_SEGMENTS, c_segments);
        atomicAdd(a.stats + RT_STAT_VERTICES, c_vertices); atomicAdd(a.stats + RT_STAT_ATTEMPTS, c_attempts);
        atomicAdd(a.stats + RT_STAT_NODE_TESTS, cnt.node_tests); atomicAdd(a.stats + RT_STAT_TRI_TESTS, cnt.tri_tests);
        atomicAdd(a.stats + RT_STAT_LIGHT_TRI_TESTS, cnt.light_tri_tests); atomicAdd(a.stats + RT_STAT_CAP_HITS, c_cap);
        atomicAdd(a.stats + RT_STAT_NONFINITE, c_nonfinite);
    }
}

template <class Space, bool STATS>
static cudaError_t launch_render_t(const RenderArgs& a, int device_sms, cudaStream_t stream, KernelInfo* info, bool launch, int* lanes) {
    const uint32_t smem = a.stack_entries * RT_BLOCK * 4u + (IsSmem<Space>::value ? a.L.total_bytes : 0u);
    cudaError_t e = cudaFuncSetAttribute(render_kernel<Space, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_kernel<Space, STATS>, RT_BLOCK, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const int grid = per_sm * device_sms;
    if (lanes) *lanes = grid * RT_BLOCK;
    if (info) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, render_kernel<Space, STATS>);
        info->block = RT_BLOCK; info->blocks_per_sm = per_sm; info->regs = fa.numRegs; info->smem_bytes = (int)smem; info->grid = grid;
    }
    if (!launch) return cudaSuccess;
    render_kernel<Space, STATS><<<grid, RT_BLOCK, smem, stream>>>(a);
    return cudaGetLastError();
}

#include "rt_render_wave.cuh"

// variant: 1 = v1 per-lane megakernel, 2 = v3 warp-local wavefront with while-while bursts, 3 = v3 with phased bursts.
// cfg (v3 only): resident-thread configuration, see wave_cfg_name().  General-primitive scenes (a.L.general) run on the phased
// wavefront kernel only, 320 threads x 2 blocks per SM (the wider vertex code wants more than the 80 registers of 384 x 2; cfg 0 = 256 x 2).
static cudaError_t dispatch_render(const RenderArgs& a, int variant, int cfg, bool use_smem, bool stats, int sms, cudaStream_t s, KernelInfo* info, bool launch, int* lanes) {
#define RT_DISPATCH(FN, ...)                                                                                                   \
    (use_smem ? (stats ? FN<SmemSpace, true __VA_ARGS__>(a, sms, s, info, launch, lanes) : FN<SmemSpace, false __VA_ARGS__>(a, sms, s, info, launch, lanes)) \
              : (stats ? FN<GmemSpace, true __VA_ARGS__>(a, sms, s, info, launch, lanes) : FN<GmemSpace, false __VA_ARGS__>(a, sms, s, info, launch, lanes)))
#define RT_WAVE(MODE)                                                     \
    switch (cfg) {                                                        \
        case 1: return RT_DISPATCH(launch_wave_t, , MODE, 256, 3);        \
        case 2: return RT_DISPATCH(launch_wave_t, , MODE, 384, 2);        \
        case 3: return RT_DISPATCH(launch_wave_t, , MODE, 256, 4);        \
        case 4: return RT_DISPATCH(launch_wave_t, , MODE, 512, 2);        \
        case 5: return RT_DISPATCH(launch_wave_t, , MODE, 352, 2);        \
        case 6: return RT_DISPATCH(launch_wave_t, , MODE, 320, 2);        \
        default: return RT_DISPATCH(launch_wave_t, , MODE, 256, 2);       \
    }
    if (a.L.quant && !a.L.general && !use_smem && variant == 3) {      // global-memory triangle scenes walk the quantised nodes (rt_api.cu flatten_scene)
#define RT_QUANT(...) (stats ? launch_wave_t<GmemSpace, true, __VA_ARGS__, false, 2>(a, sms, s, info, launch, lanes) : launch_wave_t<GmemSpace, false, __VA_ARGS__, false, 2>(a, sms, s, info, launch, lanes))
        switch (cfg) {
            case 0: return RT_QUANT(1, 256, 2);
            case 1: return RT_QUANT(1, 256, 3);
            case 5: return RT_QUANT(1, 352, 2);
            case 6: return RT_QUANT(1, 320, 2);
            default: return RT_QUANT(1, 384, 2);
        }
#undef RT_QUANT
    }
    if (a.L.general && a.L.flat_scan && use_smem && variant == 3) {    // a handful of primitives: no tree, one scan per trace burst
#define RT_FLAT(...) (stats ? launch_wave_t<SmemSpace, true, __VA_ARGS__, true, 3>(a, sms, s, info, launch, lanes) : launch_wave_t<SmemSpace, false, __VA_ARGS__, true, 3>(a, sms, s, info, launch, lanes))
        switch (cfg) {
            case 0: return RT_FLAT(1, 256, 2);
            case 1: return RT_FLAT(1, 256, 3);
            case 2: return RT_FLAT(1, 384, 2);
            default: return RT_FLAT(1, 320, 2);
        }
#undef RT_FLAT
    }
    if (a.L.general) {
        if (variant != 3) return cudaErrorInvalidValue;
        if (cfg == 0) return RT_DISPATCH(launch_wave_t, , 1, 256, 2, true);
        return RT_DISPATCH(launch_wave_t, , 1, 320, 2, true);      // measured: 320 x 2 (95 registers, no spills) beats 256 / 288 / 352 / 384 x 2 by 4-11 % on the text scenes
    }
    if (variant == 1) return RT_DISPATCH(launch_render_t);
    if (variant == 3) { RT_WAVE(1) }
    RT_WAVE(0)
#undef RT_WAVE
#undef RT_DISPATCH
}
cudaError_t launch_render(const RenderArgs& a, int variant, int cfg, bool use_smem, bool stats, int device_sms, cudaStream_t stream, KernelInfo* info) {
    return dispatch_render(a, variant, cfg, use_smem, stats, device_sms, stream, info, true, nullptr);
}
cudaError_t render_resident_lanes(int variant, int cfg, bool use_smem, bool stats, bool general, int node_format, uint32_t blob_bytes, uint32_t stack_entries, int device_sms, int* lanes) {
    RenderArgs a;
    memset(&a, 0, sizeof(a));
    a.L.total_bytes = blob_bytes; a.stack_entries = stack_entries; a.L.general = general ? 1 : 0; a.L.quant = node_format == 2; a.L.flat_scan = node_format == 3;
    return dispatch_render(a, variant, cfg, use_smem, stats, device_sms, 0, nullptr, false, lanes);
}

// ------------------------------------------------------------------------------------------------ resolve kernels
// Deterministic reduction of the per-chunk layers (fixed order), optionally added to an existing accumulator.
__global__ void sum_layers_kernel(const float4* __restrict__ layers, int n_layers, size_t n_pix, float4* __restrict__ accum, int add) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (size_t)gridDim.x * blockDim.x) {
        float4 s = add ? accum[p] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < n_layers; ++k) { const float4 v = layers[(size_t)k * n_pix + p]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
        accum[p] = s;
    }
}

// ray_depth == 0: get_ray_color returns black before it looks at the scene (rendering.rs:93-95) -- every owned pixel receives
// n_samples samples of zero radiance, no ray is traced.
__global__ void black_layer_kernel(float4* __restrict__ layer, int W, int H, uint32_t tiles_x, uint32_t shard_index, uint32_t shard_count, float n_samples) {
    const size_t n_pix = (size_t)W * (size_t)H;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (size_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(p % (size_t)W), y = (uint32_t)(p / (size_t)W);
        const uint32_t tile = (y >> 2) * tiles_x + (x >> 3);
        layer[p] = make_float4(0.f, 0.f, 0.f, tile % shard_count == shard_index ? n_samples : 0.f);
    }
}

// color_to_pixel (rendering.rs:250-262) in f64 like the reference: mean -> aces_tonemap (:236-248) -> saturate ->
// powf(1/2.2) -> round (half away from zero) -> saturating `as u8` (NaN -> 0).
__device__ __forceinline__ unsigned char tonemap_channel(double x) {
    const double a = 2.51, b = 0.03, c = 2.43, d = 0.59, e = 0.14;
    double y = (x * (a * x + b)) / (x * (c * x + d) + e);
    y = y < 0.0 ? 0.0 : (y > 1.0 ? 1.0 : y);             // f64::clamp keeps NaN
    const double g = pow(y, 1.0 / 2.2) * 255.0;
    const double r = round(g);
    if (!(r == r) || r <= 0.0) return 0;
    if (r >= 255.0) return 255;
    return (unsigned char)r;
}
__global__ void resolve_u8_kernel(const float4* __restrict__ accum, size_t n_pix, unsigned char* __restrict__ rgb) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (size_t)gridDim.x * blockDim.x) {
        const float4 s = accum[p];
        const double n = (double)s.w;                      // total_color / samples (rendering.rs:61)
        rgb[3 * p + 0] = tonemap_channel((double)s.x / n);
        rgb[3 * p + 1] = tonemap_channel((double)s.y / n);
        rgb[3 * p + 2] = tonemap_channel((double)s.z / n);
    }
}
__global__ void resolve_linear_kernel(const float4* __restrict__ accum, size_t n_pix, float* __restrict__ rgb) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (size_t)gridDim.x * blockDim.x) {
        const float4 s = accum[p];
        const bool any = s.w > 0.f;                             // pixels outside this call's tile shard hold no samples: 0
        rgb[3 * p + 0] = any ? s.x / s.w : 0.f; rgb[3 * p + 1] = any ? s.y / s.w : 0.f; rgb[3 * p + 2] = any ? s.z / s.w : 0.f;
    }
}
static int grid_for(size_t n, int block) { size_t g = (n + (size_t)block - 1) / (size_t)block; return (int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1)); }
cudaError_t launch_black_layer(float4* layer, int W, int H, uint32_t tiles_x, uint32_t shard_index, uint32_t shard_count, float n_samples, cudaStream_t stream) {
    black_layer_kernel<<<grid_for((size_t)W * (size_t)H, 256), 256, 0, stream>>>(layer, W, H, tiles_x, shard_index, shard_count, n_samples);
    return cudaGetLastError();
}
cudaError_t launch_sum_layers(const float4* layers, int n_layers, size_t n_pix, float4* accum, bool add, cudaStream_t stream) {
    sum_layers_kernel<<<grid_for(n_pix, 256), 256, 0, stream>>>(layers, n_layers, n_pix, accum, add ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t launch_resolve_u8(const float4* accum, size_t n_pix, uint8_t* rgb, cudaStream_t stream) {
    resolve_u8_kernel<<<grid_for(n_pix, 256), 256, 0, stream>>>(accum, n_pix, rgb);
    return cudaGetLastError();
}
cudaError_t launch_resolve_linear(const float4* accum, size_t n_pix, float* rgb, cudaStream_t stream) {
    resolve_linear_kernel<<<grid_for(n_pix, 256), 256, 0, stream>>>(accum, n_pix, rgb);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ nearest-hit query kernel
// f64 flavour of trace_nearest: FP32 (conservative) boxes, f64 triangle tests on the f64 copies of a/e1/e2.
__device__ __forceinline__ bool tri_test_f64(const double* o, const double* d, const double* tri, double& t, double& u, double& v) {
    const double *a = tri, *e1 = tri + 3, *e2 = tri + 6;
    const double px = d[1] * e2[2] - d[2] * e2[1], py = d[2] * e2[0] - d[0] * e2[2], pz = d[0] * e2[1] - d[1] * e2[0];
    const double det = e1[0] * px + e1[1] * py + e1[2] * pz;
    if (det == 0.0) return false;
    const double inv = 1.0 / det;
    const double tx = o[0] - a[0], ty = o[1] - a[1], tz = o[2] - a[2];
    u = (tx * px + ty * py + tz * pz) * inv;
    const double qx = ty * e1[2] - tz * e1[1], qy = tz * e1[0] - tx * e1[2], qz = tx * e1[1] - ty * e1[0];
    v = (d[0] * qx + d[1] * qy + d[2] * qz) * inv;
    t = (e2[0] * qx + e2[1] * qy + e2[2] * qz) * inv;
    return u >= 0.0 && v >= 0.0 && u + v <= 1.0 && t > 0.0;
}

// hits (optional, 9 doubles per ray): t, normal_geometry, normal_shading (normalised here for comparison), original id,
// is_outer_to_inner -- the vertex frame exactly as the render kernel builds it (hit_vertex).
template <bool F64, bool GEN>
__global__ void __launch_bounds__(128) trace_rays_kernel(const char* blob, const SceneLayout L, const double* __restrict__ tri_d, uint32_t stack_entries,
                                                         const double* __restrict__ rays, long long n, int32_t* __restrict__ tri_id, double* __restrict__ t_out,
                                                         double* __restrict__ hits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    SmemStack st; st.init(smem_base + threadIdx.x * 4u, blockDim.x * 4u, stack_entries);
    GmemSpace sp; sp.base = blob;
#ifdef RT_DEBUG_BOUNDS
    sp.limit = L.total_bytes;
#endif
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double od[3] = {rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]}, dd[3] = {rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]};
    const float3 o = f3((float)od[0], (float)od[1], (float)od[2]), d = f3((float)dd[0], (float)dd[1], (float)dd[2]);
    int best = -1; double best_t = 1.0 / 0.0;
    Hit hit; hit.t = RT_INF_F; hit.u = 0.f; hit.v = 0.f; hit.tri = -1;
    if (!F64) {
        Counters cnt; cnt.node_tests = 0; cnt.tri_tests = 0; cnt.light_tri_tests = 0;
        if (GEN) trace_nearest_gen<GmemSpace, false>(sp, L, st, o, d, -1, hit, cnt);
        else if (L.quant) trace_nearest<GmemSpace, false, 2>(sp, L, st, o, d, -1, hit, cnt);      // the walk the render kernel of this scene runs
        else trace_nearest<GmemSpace, false>(sp, L, st, o, d, -1, hit, cnt);
        best = hit.tri; best_t = hit.tri >= 0 ? (double)hit.t : best_t;
    } else {
        const RaySetup r = ray_setup(o, d, L.nodes);
        float best_tf = RT_INF_F;
        int cur = 0;
        for (;;) {
            while (cur >= 0) pair_step(sp, L.nodes, r, best_tf, cur, st);
            if (cur == RT_CUR_DONE) break;
            const uint32_t code = (uint32_t)~cur;
            const int first = (int)(code >> 3), cntl = (int)(code & 7u) + 1;
            for (int k = first; k < first + cntl; ++k) {
                double t, u, v;
                if (tri_test_f64(od, dd, tri_d + (size_t)k * 9, t, u, v) && t < best_t) {
                    best_t = t; best = k;
                    best_tf = __double2float_ru(t) * 1.000001f;   // prune bound rounded up: boxes stay conservative
                }
            }
            cur = st.pop();
        }
    }
    const int orig = best >= 0 ? __float_as_int(sp.ld4(L.sh_dn1 + (uint32_t)best * 16u).w) : -1;
    if (tri_id) tri_id[i] = orig;
    if (t_out) t_out[i] = best_t;
    if (hits && !F64) {
        double* h = hits + 9 * i;
        if (best < 0) { h[0] = best_t; for (int k = 1; k < 7; ++k) h[k] = 0.0; h[7] = -1.0; h[8] = 0.0; return; }
        const Vertex V = hit_vertex<GmemSpace, GEN>(sp, L, best, sp.ld4(L.sh_n0 + (uint32_t)best * 16u), o, d, hit.u, hit.v);
        const float3 ns = normalize(V.ns);
        h[0] = best_t; h[1] = V.n.x; h[2] = V.n.y; h[3] = V.n.z; h[4] = ns.x; h[5] = ns.y; h[6] = ns.z; h[7] = (double)orig; h[8] = V.outer ? 1.0 : 0.0;
    }
}
cudaError_t launch_trace_rays(const char* blob, const SceneLayout& L, const double* tri_d, uint32_t stack_entries, const double* rays, long long n,
                              bool f64, int32_t* tri_id, double* t, double* hits, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if (L.general && f64) return cudaErrorInvalidValue;          // f64 primitive tests exist for triangles only
    const int block = 128;
    const uint32_t smem = stack_entries * block * 4u;
    const int grid = (int)((n + block - 1) / block);
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, block, smem, stream>>>(blob, L, tri_d, stack_entries, rays, n, tri_id, t, hits);
        return cudaGetLastError();
    };
    if (f64) return go(trace_rays_kernel<true, false>);
    if (L.general) return go(trace_rays_kernel<false, true>);
    return go(trace_rays_kernel<false, false>);
}

__global__ void primary_rays_kernel(const Camera cam, int W, int H, const int32_t* __restrict__ xy, const double* __restrict__ xi, long long n, double* __restrict__ rays) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 o, d;
    camera_ray(cam, W, H, xy[2 * i], xy[2 * i + 1], (float)xi[2 * i], (float)xi[2 * i + 1], o, d);
    rays[6 * i + 0] = o.x; rays[6 * i + 1] = o.y; rays[6 * i + 2] = o.z;
    rays[6 * i + 3] = d.x; rays[6 * i + 4] = d.y; rays[6 * i + 5] = d.z;
}
cudaError_t launch_primary_rays(const Camera& cam, int W, int H, const int32_t* xy, const double* xi, long long n, double* rays, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    primary_rays_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(cam, W, H, xy, xi, n, rays);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ unit-function kernel
int eval_in_width(int fn) {
    switch (fn) {
        case RT_FN_BRDF: return 14; case RT_FN_PDF_COSINE: return 6; case RT_FN_PDF_VNDF: return 10; case RT_FN_PDF_LIGHT: return 6;
        case RT_FN_PDF_MIX: return 13; case RT_FN_SAMPLE_COSINE: return 5; case RT_FN_SAMPLE_VNDF: return 9; case RT_FN_SAMPLE_LIGHT: return 6;
        case RT_FN_PHILOX: return 4; case RT_FN_SAMPLE_LIGHT_GEN: return 8; case RT_FN_DIELECTRIC: return 9; default: return 0;
    }
}
int eval_out_width(int fn) {
    switch (fn) {
        case RT_FN_BRDF: return 3; case RT_FN_PDF_COSINE: case RT_FN_PDF_VNDF: case RT_FN_PDF_LIGHT: case RT_FN_PDF_MIX: return 1;
        case RT_FN_SAMPLE_COSINE: return 6; case RT_FN_SAMPLE_VNDF: case RT_FN_SAMPLE_LIGHT: case RT_FN_SAMPLE_LIGHT_GEN: return 3; case RT_FN_PHILOX: case RT_FN_DIELECTRIC: return 4; default: return 0;
    }
}

// Evaluates the SAME device functions the render kernel calls, on caller-supplied inputs.
__global__ void __launch_bounds__(128) eval_kernel(const char* blob, const SceneLayout L, uint32_t stack_entries, int fn, int win, int wout,
                                                   const float* __restrict__ in, long long n, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    SmemStack st; st.init(smem_base + threadIdx.x * 4u, blockDim.x * 4u, stack_entries);
    GmemSpace sp; sp.base = blob;
#ifdef RT_DEBUG_BOUNDS
    sp.limit = L.total_bytes;
#endif
    Counters cnt; cnt.node_tests = 0; cnt.tri_tests = 0; cnt.light_tri_tests = 0;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* x = in + i * win;
    float* y = out + i * wout;
    if (fn == RT_FN_BRDF) {
        const float3 l = f3(x[0], x[1], x[2]), nn = f3(x[3], x[4], x[5]), v = f3(x[6], x[7], x[8]);
        Material m; m.base = f3(x[9], x[10], x[11]); m.metallic = x[12]; m.roughness = x[13]; m.emission = f3(0.f, 0.f, 0.f);
        const float alpha = m.roughness * m.roughness, alpha2 = alpha * alpha;
        const DirTerms t = dir_terms(nn, v, l, alpha2);
        const float nv = dot(nn, v);
        const float3 f = brdf_eval(m, t.nh > 0.0f ? t.d_nochi : 0.0f, ggx_g1(t.nl, alpha2), ggx_g1(nv, alpha2), t.nl, nv, t.hl);
        y[0] = f.x; y[1] = f.y; y[2] = f.z;
    } else if (fn == RT_FN_PDF_COSINE) {
        y[0] = pdf_cosine(dot(f3(x[0], x[1], x[2]), f3(x[3], x[4], x[5])));
    } else if (fn == RT_FN_PDF_VNDF) {
        const float3 nn = f3(x[0], x[1], x[2]), l = f3(x[3], x[4], x[5]), v = f3(x[6], x[7], x[8]);
        const float alpha = x[9] * x[9], alpha2 = alpha * alpha;
        const DirTerms t = dir_terms(nn, v, l, alpha2);
        const float nv = dot(nn, v);
        y[0] = pdf_vndf(t.d_nochi, ggx_g1(nv, alpha2), nv);
    } else if (fn == RT_FN_PDF_LIGHT) {
        const float3 P = f3(x[0], x[1], x[2]), l = f3(x[3], x[4], x[5]);
        y[0] = L.n_lights <= 0 ? 0.0f : (L.general ? light_pdf<GmemSpace, false, true>(sp, L, st, P, l, cnt) : light_pdf<GmemSpace, false>(sp, L, st, P, l, cnt));
    } else if (fn == RT_FN_PDF_MIX) {
        const float3 P = f3(x[0], x[1], x[2]), nn = f3(x[3], x[4], x[5]), l = f3(x[6], x[7], x[8]), v = f3(x[9], x[10], x[11]);
        const float alpha = x[12] * x[12], alpha2 = alpha * alpha;
        const DirTerms t = dir_terms(nn, v, l, alpha2);
        const float nv = dot(nn, v);
        float pdf = pdf_cosine(t.nl) + pdf_vndf(t.d_nochi, ggx_g1(nv, alpha2), nv);
        const int n_comp = L.n_lights > 0 ? 3 : 2;
        if (n_comp == 3) pdf += L.general ? light_pdf<GmemSpace, false, true>(sp, L, st, P, l, cnt) : light_pdf<GmemSpace, false>(sp, L, st, P, l, cnt);
        y[0] = pdf * (1.0f / (float)n_comp);
    } else if (fn == RT_FN_SAMPLE_COSINE) {
        const float3 nn = f3(x[0], x[1], x[2]);
        const float3 l = sample_cosine(nn, x[3], x[4]);
        const float3 s = sphere_uniform(x[3], x[4]);
        y[0] = l.x; y[1] = l.y; y[2] = l.z; y[3] = s.x; y[4] = s.y; y[5] = s.z;
    } else if (fn == RT_FN_SAMPLE_VNDF) {
        const float3 l = sample_vndf(f3(x[0], x[1], x[2]), f3(x[3], x[4], x[5]), x[6] * x[6], x[7], x[8]);
        y[0] = l.x; y[1] = l.y; y[2] = l.z;
    } else if (fn == RT_FN_SAMPLE_LIGHT) {
        const float3 l = sample_light(sp, L, f3(x[0], x[1], x[2]), (int)x[3], x[4], x[5]);
        y[0] = l.x; y[1] = l.y; y[2] = l.z;
    } else if (fn == RT_FN_SAMPLE_LIGHT_GEN) {
        const uint32_t bits = (x[7] > 0.0f ? 0x80000000u : 0u) | ((uint32_t)(x[6] * 16777216.0f) << 7);     // sign | x01 (24 bits)
        const float3 l = sample_light_gen(sp, L, f3(x[0], x[1], x[2]), (int)x[3], x[4], x[5], bits);
        y[0] = l.x; y[1] = l.y; y[2] = l.z;
    } else if (fn == RT_FN_DIELECTRIC) {
        const float3 nn = f3(x[0], x[1], x[2]), v = f3(x[3], x[4], x[5]);
        float3 l;
        const bool refract = dielectric_sample(nn, v, dot(nn, v), x[6], x[7] > 0.0f, x[8], l);
        y[0] = l.x; y[1] = l.y; y[2] = l.z; y[3] = refract ? 1.0f : 0.0f;
    } else if (fn == RT_FN_PHILOX) {
        const uint4 r = philox4x32_10(make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), RT_PHILOX_TAG),
                                      make_uint2(__float_as_uint(x[3]), 0u));
        y[0] = __uint_as_float(r.x); y[1] = __uint_as_float(r.y); y[2] = __uint_as_float(r.z); y[3] = __uint_as_float(r.w);
    }
}
cudaError_t launch_eval(const char* blob, const SceneLayout& L, uint32_t stack_entries, int fn, const float* in, long long n, float* out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int block = 128;
    const uint32_t smem = stack_entries * block * 4u;
    cudaError_t e = cudaFuncSetAttribute(eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    eval_kernel<<<(int)((n + block - 1) / block), block, smem, stream>>>(blob, L, stack_entries, fn, eval_in_width(fn), eval_out_width(fn), in, n, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ debug bounds counters
// out: RT_BOUNDS_KINDS counters of this translation unit's kernels (all the kernels that touch scene data live here).  Returns
// cudaErrorNotSupported in a regular build.  reset != 0 zeroes them afterwards.
cudaError_t read_bounds_violations(unsigned long long* out, int reset) {
#ifdef RT_DEBUG_BOUNDS
    cudaError_t e = cudaMemcpyFromSymbol(out, g_bounds_violations, sizeof(unsigned long long) * RT_BOUNDS_KINDS);
    if (e == cudaSuccess && reset) {
        unsigned long long zero[RT_BOUNDS_KINDS] = {0};
        e = cudaMemcpyToSymbol(g_bounds_violations, zero, sizeof(zero));
    }
    return e;
#else
    (void)out; (void)reset;
    return cudaErrorNotSupported;
#endif
}

// ------------------------------------------------------------------------------------------------ FP32 issue micro-benchmark
// 16 independent FFMA chains per thread, 3-register form; 2 flop per lane per FFMA.  The measured rate is the
// denominator of the FP32 roofline (MEASURED_PEAKS.json has no FP32 entry).
__global__ void __launch_bounds__(256) ffma_kernel(int iters, float* sink) {
    float acc[16];
    const float a = 1.0000001f + (float)threadIdx.x * 1e-9f, b = 1e-7f * (float)(blockIdx.x + 1);
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = (float)k + (float)threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += acc[k];
    if (s == 123.456f) sink[0] = s;   // never true in practice; keeps the chains alive
}
cudaError_t launch_ffma(int blocks, int threads, int iters, float* sink, cudaStream_t stream) {
    ffma_kernel<<<blocks, threads, 0, stream>>>(iters, sink);
    return cudaGetLastError();
}

}  // namespace rtd
