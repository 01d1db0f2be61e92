// rt_render_pool.cuh -- render kernel v2: the warp-local wavefront.  Included by rt_kernels.cu after the shading
// helpers (DirTerms, mix_sample_and_pdf, camera_ray, load_material, SpaceMaker, IsSmem).
//
// ncu of v1 (profiles/r1_v1_*): 31 % warp execution efficiency -- the node loop issues 48 % of all instructions at
// 8.3/32 active lanes and the rejection loop 23 % at 9.5/32, because every lane owns ONE path and the warp waits for
// its slowest ray / its unluckiest rejection in every iteration.  v2 decouples lanes from paths:
//   * every warp owns a POOL of P path slots (P = 64 or 96) in shared memory (SoA, 22 words per slot);
//   * S phase (shade): P/32 rounds, lane i handles slot r*32+i: finished items are stored and replaced, ended paths
//     are regenerated, and every slot that holds a hit makes exactly ONE attempt of the reference's rejection loop
//     (rendering.rs:102-110).  A rejected slot simply stays in the HIT state and retries in the next S phase, so no
//     lane ever waits for another lane's retry.  Slots with a ray go on a compacted trace list;
//   * T phase (trace): lanes pull rays from the trace list; the moment a lane's traversal ends it writes the hit to
//     its slot and takes the next ray from the list (warp-synchronous ballot, no atomics), so traversal-length
//     variance is absorbed by the pool instead of idling lanes.
// Same estimator, same Philox counters (pixel, sample, call#) as v1 -> same image.
#pragma once

enum { F_OX = 0, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_SKIP, F_TRI, F_U, F_V, F_TX, F_TY, F_TZ, F_AX, F_AY, F_AZ, F_PIX, F_S, F_SSTOP, F_CHUNK, F_META, F_LIST, NF };
enum { ST_NEED_ITEM = 0, ST_NEED_PATH = 1, ST_GEN = 2, ST_TRACE = 3, ST_HIT = 4, ST_EXHAUSTED = 5 };

template <int P>
struct Pool {
    uint32_t base;
    RT_DEV uint32_t at(int f, int s) const { return base + (uint32_t)(f * P + s) * 4u; }
    RT_DEV float ldf(int f, int s) const { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(at(f, s))); return v; }
    RT_DEV int ldi(int f, int s) const { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(at(f, s))); return v; }
    RT_DEV void stf(int f, int s, float v) const { asm volatile("st.shared.f32 [%0], %1;" ::"r"(at(f, s)), "f"(v)); }
    RT_DEV void sti(int f, int s, int v) const { asm volatile("st.shared.s32 [%0], %1;" ::"r"(at(f, s)), "r"(v)); }
    RT_DEV float3 ld3(int f, int s) const { return f3(ldf(f, s), ldf(f + 1, s), ldf(f + 2, s)); }
    RT_DEV void st3(int f, int s, float3 v) const { stf(f, s, v.x); stf(f + 1, s, v.y); stf(f + 2, s, v.z); }
};
// meta word of a slot: state (3 bits) | segments left (8) | attempts made at this vertex (7) | Philox call counter (14)
RT_DEV uint32_t pack_meta(int state, int depth, int attempt, uint32_t call) { return (uint32_t)state | ((uint32_t)depth << 3) | ((uint32_t)attempt << 11) | (call << 18); }

#ifndef RT_POOL_MIN_BLOCKS
#define RT_POOL_MIN_BLOCKS 2
#endif

template <class Space, bool STATS, int P>
__global__ void __launch_bounds__(RT_BLOCK, RT_POOL_MIN_BLOCKS) render_pool_kernel(const RenderArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t stack_bytes = a.stack_entries * blockDim.x * 4u;
    const uint32_t blob_bytes = IsSmem<Space>::value ? a.L.total_bytes : 0u;
    if (IsSmem<Space>::value) {
        const uint4* src = reinterpret_cast<const uint4*>(a.blob);
        uint4* dst = reinterpret_cast<uint4*>(smem_raw + stack_bytes);
        for (uint32_t i = threadIdx.x; i < a.L.total_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
    }
    const Space sp = SpaceMaker<Space>::make(a.blob, smem_base + stack_bytes);
    const SceneLayout& L = a.L;
    SmemStack st; st.addr = smem_base + threadIdx.x * 4u; st.stride = blockDim.x * 4u;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    Pool<P> pool; pool.base = smem_base + stack_bytes + blob_bytes + warp * (uint32_t)(P * NF * 4);
    const uint2 key = make_uint2(a.seed_lo, a.seed_hi);
    const float3 bg = f3(a.bg[0], a.bg[1], a.bg[2]);
    const bool bg_nonzero = (a.bg[0] != 0.f) | (a.bg[1] != 0.f) | (a.bg[2] != 0.f);
    const size_t n_pix = (size_t)a.W * (size_t)a.H;

    Counters cnt; cnt.node_tests = 0; cnt.tri_tests = 0; cnt.light_tri_tests = 0;
    unsigned long long c_samples = 0, c_segments = 0, c_vertices = 0, c_attempts = 0, c_cap = 0, c_nonfinite = 0;

    for (int s = (int)lane; s < P; s += 32) pool.sti(F_META, s, (int)pack_meta(ST_NEED_ITEM, 0, 0, 0u));
    __syncwarp();

    for (;;) {
        // ======================================================================================= S phase
        int n_trace = 0;
        bool pending = false;
#pragma unroll 1
        for (int r = 0; r < P / 32; ++r) {
            const int s = r * 32 + (int)lane;
            uint32_t meta = (uint32_t)pool.ldi(F_META, s);
            int state = (int)(meta & 7u);
            if (state == ST_HIT) {                                           // get_ray_color (rendering.rs:86-127), one attempt
                int depth = (int)((meta >> 3) & 0xffu), attempt = (int)((meta >> 11) & 0x7fu);
                uint32_t call = meta >> 18;
                const int tri = pool.ldi(F_TRI, s);
                float3 T = pool.ld3(F_TX, s);
                bool end_path = false;
                if (tri < 0) {                                               // :125
                    if (bg_nonzero) pool.st3(F_AX, s, pool.ld3(F_AX, s) + T * bg);
                    end_path = true;
                } else {
                    const uint32_t o16 = (uint32_t)tri * 16u;
                    const float4 n0 = sp.ld4(L.sh_n0 + o16);
                    const Material mat = load_material(sp, L, __float_as_int(n0.w));
                    if (attempt == 0) {
                        if ((mat.emission.x != 0.f) | (mat.emission.y != 0.f) | (mat.emission.z != 0.f))
                            pool.st3(F_AX, s, pool.ld3(F_AX, s) + T * mat.emission);   // :99
                        if (--depth <= 0) end_path = true;                   // :93-95
                        else if (STATS) ++c_vertices;
                    }
                    if (!end_path) {
                        const float3 d = pool.ld3(F_DX, s);
                        const float hu = pool.ldf(F_U, s), hv = pool.ldf(F_V, s);
                        const float3 ng = f3(sp.ld4(L.sh_ng + o16));
                        const float sgn = dot(ng, d) < 0.0f ? 1.0f : -1.0f;  // geometry.rs:115-126
                        const float3 n = ng * sgn;
                        const float3 dn1 = f3(sp.ld4(L.sh_dn1 + o16)), dn2 = f3(sp.ld4(L.sh_dn2 + o16));
                        const float3 ns = (f3(n0) + dn1 * hu + dn2 * hv) * sgn;
                        const float3 ta = f3(sp.ld4(L.tri_a + o16)), te1 = f3(sp.ld4(L.tri_e1 + o16)), te2 = f3(sp.ld4(L.tri_e2 + o16));
                        const float3 Pt = fma3(d, -RT_EPS_F, fma3(te2, hv, fma3(te1, hu, ta)));   // :98
                        const float3 v = -d;
                        const float nv = dot(n, v);
                        const float alpha = mat.roughness * mat.roughness, alpha2 = alpha * alpha;
                        const float g1v = ggx_g1(nv, alpha2);
                        const uint32_t pix = (uint32_t)pool.ldi(F_PIX, s);
                        const int s_this = pool.ldi(F_S, s);
                        const uint4 rnd = philox4x32_10(make_uint4(pix, (uint32_t)s_this, call, RT_PHILOX_TAG), key);
                        call = (call + 1u) & 0x3fffu;
                        float3 l; DirTerms terms;
                        const float pdf = mix_sample_and_pdf<Space, STATS>(sp, L, st, a.n_comp, Pt, n, v, nv, alpha, alpha2, g1v, rnd, l, terms, cnt);
                        ++attempt;
                        if (STATS) ++c_attempts;
                        if (pdf > 0.0f && dot(l, ns) > 0.0f) {               // :107
                            const float d_chi = terms.nh > 0.0f ? terms.d_nochi : 0.0f;
                            const float3 f = brdf_eval(mat, d_chi, ggx_g1(terms.nl, alpha2), g1v, terms.nl, nv, terms.hl);
                            T = T * f * (terms.nl * fast_rcp(pdf));          // :122
                            if (!finite3(T)) { if (STATS) ++c_nonfinite; end_path = true; }
                            else {
                                pool.st3(F_OX, s, Pt); pool.st3(F_DX, s, l); pool.st3(F_TX, s, T);
                                pool.sti(F_SKIP, s, terms.nl > 0.0f ? tri : -1);
                                attempt = 0; state = ST_TRACE;
                            }
                        } else if (attempt >= a.max_attempts || attempt >= 127) {
                            if (STATS) ++c_cap;
                            end_path = true;
                        }
                    }
                }
                if (end_path) state = ST_NEED_PATH;
                meta = pack_meta(state, depth, attempt, call);
            }
            if (state == ST_NEED_PATH) {                                     // next sample of this item, or store the item
                const int s_next = pool.ldi(F_S, s) + 1, s_stop = pool.ldi(F_SSTOP, s);
                if (s_next < s_stop) { pool.sti(F_S, s, s_next); state = ST_GEN; }
                else {
                    const int chunk = pool.ldi(F_CHUNK, s);
                    const uint32_t pix = (uint32_t)pool.ldi(F_PIX, s);
                    const float3 acc = pool.ld3(F_AX, s);
                    const float n_done = (float)(s_stop - (a.s_begin + chunk * a.chunk_size));
                    a.layers[(size_t)chunk * n_pix + pix] = make_float4(acc.x, acc.y, acc.z, n_done);
                    state = ST_NEED_ITEM;
                }
            }
            for (;;) {                                                       // warp-aggregated work fetch
                const bool need = state == ST_NEED_ITEM;
                const unsigned m = __ballot_sync(0xffffffffu, need);
                if (m == 0u) break;
                const int leader = __ffs(m) - 1;
                unsigned int base = 0;
                if ((int)lane == leader) base = atomicAdd(a.work_counter, (unsigned int)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (need) {
                    const uint32_t idx = base + (uint32_t)__popc(m & lt_mask);
                    if (idx >= a.total_items) state = ST_EXHAUSTED;
                    else {
                        const uint32_t chunk = idx / a.n_pix_items;
                        const uint32_t rr = idx - chunk * a.n_pix_items;
                        const uint32_t tile = rr >> 5, w = rr & 31u;
                        const uint32_t ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
                        const int px = (int)(tx * 8u + (w & 7u)), py = (int)(ty * 4u + (w >> 3));
                        if (px < a.W && py < a.H) {
                            const int s0 = a.s_begin + (int)chunk * a.chunk_size;
                            pool.sti(F_PIX, s, (int)((uint32_t)py * (uint32_t)a.W + (uint32_t)px));
                            pool.sti(F_CHUNK, s, (int)chunk);
                            pool.sti(F_S, s, s0);
                            pool.sti(F_SSTOP, s, min(s0 + a.chunk_size, a.s_end));
                            pool.st3(F_AX, s, f3(0.f, 0.f, 0.f));
                            state = ST_GEN;
                        }
                    }
                }
            }
            if (state == ST_GEN) {                                           // get_ray_to_pixel (rendering.rs:71-84)
                const uint32_t pix = (uint32_t)pool.ldi(F_PIX, s);
                const int s_this = pool.ldi(F_S, s);
                const int py = (int)(pix / (uint32_t)a.W), px = (int)(pix - (uint32_t)py * (uint32_t)a.W);
                const uint4 rr = philox4x32_10(make_uint4(pix, (uint32_t)s_this, 0u, RT_PHILOX_TAG), key);
                float3 o, d;
                camera_ray(a.cam, a.W, a.H, px, py, u01(rr.x), u01(rr.y), o, d);
                pool.st3(F_OX, s, o); pool.st3(F_DX, s, d); pool.st3(F_TX, s, f3(1.f, 1.f, 1.f));
                pool.sti(F_SKIP, s, -1);
                state = ST_TRACE;
                meta = pack_meta(state, a.ray_depth > 255 ? 255 : a.ray_depth, 0, 1u);
                if (STATS) ++c_samples;
            } else {
                meta = (meta & ~7u) | (uint32_t)state;
            }
            pool.sti(F_META, s, (int)meta);
            const unsigned tm = __ballot_sync(0xffffffffu, state == ST_TRACE);
            if (state == ST_TRACE) pool.sti(F_LIST, n_trace + __popc(tm & lt_mask), s);
            n_trace += __popc(tm);
            pending |= __ballot_sync(0xffffffffu, state == ST_HIT) != 0u;
        }
        __syncwarp();
        if (n_trace == 0 && !pending) break;

        // ======================================================================================= T phase
        int next = 0, slot = 0, cur = 0, sptr = 0, skip_tri = -1;
        uint32_t tmeta = 0;
        bool has_ray = false;
        float3 o = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 1.f), inv = f3(1.f, 1.f, 1.f), od = f3(0.f, 0.f, 0.f);
        Hit hit; hit.t = RT_INF_F; hit.u = 0.f; hit.v = 0.f; hit.tri = -1;
        for (;;) {
            const unsigned idle = __ballot_sync(0xffffffffu, !has_ray);
            if (idle != 0u && next < n_trace) {                             // hand the next rays of the list to the idle lanes
                if (!has_ray) {
                    const int k = next + __popc(idle & lt_mask);
                    if (k < n_trace) {
                        slot = pool.ldi(F_LIST, k);
                        o = pool.ld3(F_OX, slot); d = pool.ld3(F_DX, slot);
                        skip_tri = pool.ldi(F_SKIP, slot);
                        tmeta = (uint32_t)pool.ldi(F_META, slot);
                        inv = safe_inv_dir(d); od = o * inv;
                        hit.t = RT_INF_F; hit.u = 0.f; hit.v = 0.f; hit.tri = -1;
                        cur = 0; sptr = 0; has_ray = true;
                        if (STATS) ++c_segments;
                    }
                }
                next += __popc(idle);
            }
            if (__ballot_sync(0xffffffffu, has_ray) == 0u) break;
            if (has_ray) {                                                   // one while-while round of trace_nearest (bvh.rs:231-297)
                bool finished = false;
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
                        if (sptr == 0) { finished = true; break; }
                        cur = st.load(--sptr);
                    }
                }
                if (!finished) {
                    const uint32_t code = (uint32_t)~cur;
                    const int first = (int)(code >> 3), n = (int)(code & 7u) + 1;
                    for (int i = first; i < first + n; ++i) {
                        const uint32_t o16 = (uint32_t)i * 16u;
                        const float4 ta = sp.ld4(L.tri_a + o16), e1 = sp.ld4(L.tri_e1 + o16), e2 = sp.ld4(L.tri_e2 + o16);
                        float t, u, v;
                        const bool ok = tri_test(o, d, f3(ta), f3(e1), f3(e2), t, u, v);
                        if (STATS) cnt.tri_tests += 1;
                        if (ok && t < hit.t && i != skip_tri) { hit.t = t; hit.u = u; hit.v = v; hit.tri = i; }
                    }
                    if (sptr == 0) finished = true;
                    else cur = st.load(--sptr);
                }
                if (finished) {
                    pool.sti(F_TRI, slot, hit.tri); pool.stf(F_U, slot, hit.u); pool.stf(F_V, slot, hit.v);
                    pool.sti(F_META, slot, (int)((tmeta & ~7u) | (uint32_t)ST_HIT));
                    has_ray = false;
                }
            }
        }
        __syncwarp();
    }
    if (STATS) {
        atomicAdd(a.stats + RT_STAT_SAMPLES, c_samples); atomicAdd(a.stats + RT_STAT_SEGMENTS, c_segments);
        atomicAdd(a.stats + RT_STAT_VERTICES, c_vertices); atomicAdd(a.stats + RT_STAT_ATTEMPTS, c_attempts);
        atomicAdd(a.stats + RT_STAT_NODE_TESTS, cnt.node_tests); atomicAdd(a.stats + RT_STAT_TRI_TESTS, cnt.tri_tests);
        atomicAdd(a.stats + RT_STAT_LIGHT_TRI_TESTS, cnt.light_tri_tests); atomicAdd(a.stats + RT_STAT_CAP_HITS, c_cap);
        atomicAdd(a.stats + RT_STAT_NONFINITE, c_nonfinite);
    }
}

template <class Space, bool STATS, int P>
static cudaError_t launch_pool_t(const RenderArgs& a, int device_sms, cudaStream_t stream, KernelInfo* info, bool launch, int* lanes) {
    const uint32_t smem = a.stack_entries * RT_BLOCK * 4u + (IsSmem<Space>::value ? a.L.total_bytes : 0u) + (RT_BLOCK / 32) * (uint32_t)(P * NF * 4);
    cudaError_t e = cudaFuncSetAttribute(render_pool_kernel<Space, STATS, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_pool_kernel<Space, STATS, P>, RT_BLOCK, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const int grid = per_sm * device_sms;
    if (lanes) *lanes = grid * (RT_BLOCK / 32) * P;       // resident path slots
    if (info) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, render_pool_kernel<Space, STATS, P>);
        info->block = RT_BLOCK; info->blocks_per_sm = per_sm; info->regs = fa.numRegs; info->smem_bytes = (int)smem; info->grid = grid;
    }
    if (!launch) return cudaSuccess;
    render_pool_kernel<Space, STATS, P><<<grid, RT_BLOCK, smem, stream>>>(a);
    return cudaGetLastError();
}
