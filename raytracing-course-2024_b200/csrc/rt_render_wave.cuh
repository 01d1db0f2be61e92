// rt_render_wave.cuh -- render kernel v3 (the product path): the warp-local wavefront with burst traversal.  Included by
// rt_kernels.cu after the shading helpers (DirTerms, mix_sample_and_pdf, camera_ray, load_material, SpaceMaker, IsSmem).
//
// ncu of v1 (profiles/r1_v1_*): 31 % warp execution efficiency -- the node loop issues 45 % of all instructions at
// 8.3/32 active lanes, the leaf loop 17 % at 10.3/32 and shading 38 % at ~10/32, because every lane owns ONE path and
// the warp waits for its slowest ray / its unluckiest rejection in every iteration.  A first pool whose queues were
// serviced after EVERY traversal step reached 17.7/32 lanes but doubled the thread-instruction count with ballots and
// queue arithmetic.  v3 keeps the pool and moves the bookkeeping out of the inner loop:
//   * every warp owns a POOL of 64 path slots in shared memory (32-bit SoA, 19 words per slot) and two ring queues over
//     them: TQ (slots holding a ray to trace) and SQ (slots waiting for shading) -- warp-synchronous, no atomics;
//   * TRACE BURST: the 32 lanes hold 32 rays (only what the box tests need stays in registers).  MODE 1 (default, phased):
//     box-pair steps run, three per vote, while at least `node_min` lanes want one; otherwise the lanes sitting on a leaf
//     test its triangles (leaves of <= 2 triangles side by side); the burst ends when `burst_exit` lanes are finished.
//     MODE 0: plain while-while with the same exit rule.  Finished lanes write (triangle, u, v) to their slot, push it on
//     SQ and pop the next ray of TQ;
//   * SHADE ROUND: whenever TQ cannot feed the idle lanes, 32 slots of SQ are shaded with all 32 lanes busy (the
//     in-flight traversals of the other slots stay in registers): emission / depth bookkeeping, finished items stored
//     exactly once and replaced from the frame's work counter, then ONE Philox call per live slot -- the camera jitter of a
//     new path or the next attempt of the reference's rejection loop (rendering.rs:102-110) -- and the camera ray or one
//     attempt; accepted directions go to TQ, rejected slots return to SQ and retry in a later round (nobody waits).
//     With 64 slots and 32 lanes, an empty TQ implies >= 32 slots on SQ, so rounds are full until the work counter runs dry.
// Same estimator, same Philox counters (pixel, sample, call#) as v1 -> same image (test_kernel_variants_render_the_same_image).
// ncu of this kernel (profiles/r1_v3d_*): 19.5/32 lanes per instruction, issue slots 75 % busy, shared-memory pipe 76 %.
#pragma once

#define RT_POOL_SLOTS 64
#ifndef RT_NODE_UNROLL
#define RT_NODE_UNROLL 3   /* box-pair steps per phase vote */
#endif
// Slot fields (words, SoA [field][slot]).  A slot is read by the shade round and by the lane that traces its ray, never
// by both at once, so hand-over fields share storage:
//   shade -> trace: O (ray origin), IV (safe reciprocal direction), F_TRI = triangle to skip (or -1)
//   trace -> shade: F_TRI = nearest triangle (or -1), (F_U, F_V) = its barycentrics; IV still gives the ray direction
// While a ray is in flight its lane keeps only (1/d, o/d, octant offsets, best t, best triangle, node, stack pointer) in
// registers -- they must survive the shade rounds the warp runs for OTHER slots -- rebuilds origin and direction from
// them at a leaf and re-reads only the skip triangle from the slot.
enum { F_OX = 0, F_OY, F_OZ, F_IX, F_IY, F_IZ, F_TRI, F_U, F_V, F_TX, F_TY, F_TZ, F_AX, F_AY, F_AZ, F_PIX, F_S, F_CHUNK, F_META, NF };
#define RT_POOL_QUEUE_BYTES (2u * RT_POOL_SLOTS)                       /* TQ and SQ: one byte per entry */
#define RT_POOL_WARP_BYTES (RT_POOL_SLOTS * NF * 4u + RT_POOL_QUEUE_BYTES)
enum { ST_NEED_ITEM = 0, ST_ENDED = 1, ST_GEN = 2, ST_TRACE = 3, ST_RETRY = 4, ST_EXHAUSTED = 5, ST_NONE = 6 };

struct Pool {
    uint32_t base;
    RT_DEV uint32_t at(int f, int s) const {
        RT_BOUNDS(f >= 0 && f < NF && s >= 0 && s < RT_POOL_SLOTS, RT_BOUNDS_POOL);
        return base + (uint32_t)(f * RT_POOL_SLOTS + s) * 4u;
    }
    RT_DEV float ldf(int f, int s) const { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(at(f, s))); return v; }
    RT_DEV int ldi(int f, int s) const { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(at(f, s))); return v; }
    RT_DEV void stf(int f, int s, float v) const { asm volatile("st.shared.f32 [%0], %1;" ::"r"(at(f, s)), "f"(v)); }
    RT_DEV void sti(int f, int s, int v) const { asm volatile("st.shared.s32 [%0], %1;" ::"r"(at(f, s)), "r"(v)); }
    RT_DEV float3 ld3(int f, int s) const { return f3(ldf(f, s), ldf(f + 1, s), ldf(f + 2, s)); }
    RT_DEV void st3(int f, int s, float3 v) const { stf(f, s, v.x); stf(f + 1, s, v.y); stf(f + 2, s, v.z); }
    // queues: q = 0 (TQ) or 1 (SQ), pos in [0, RT_POOL_SLOTS)
    RT_DEV uint32_t qat(int q, int pos) const {
        RT_BOUNDS((q == 0 || q == 1) && pos >= 0 && pos < RT_POOL_SLOTS, RT_BOUNDS_QUEUE);
        return base + RT_POOL_SLOTS * NF * 4u + (uint32_t)(q * RT_POOL_SLOTS + pos);
    }
    RT_DEV int qld(int q, int pos) const { int v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(qat(q, pos))); return v; }
    RT_DEV void qst(int q, int pos, int v) const { asm volatile("st.shared.u8 [%0], %1;" ::"r"(qat(q, pos)), "r"(v)); }
};
// meta word of a slot: state (3 bits) | segments left (8) | attempts made at this vertex (7) | Philox call counter (14)
RT_DEV uint32_t pack_meta(int state, int depth, int attempt, uint32_t call) { return (uint32_t)state | ((uint32_t)depth << 3) | ((uint32_t)attempt << 11) | (call << 18); }

// MODE 0: while-while burst.  MODE 1: phased burst (box-pair steps while >= node_min lanes want one, else leaf phase).
// BLOCK x MINB = resident threads per SM (occupancy / register budget trade-off, picked at run time from a few builds).
// GEN: general-primitive scenes (SceneLayout::general): leaves hold 64-byte primitive records (triangles, boxes, ellipsoids),
// the infinite primitives are scanned when a ray retires (rendering.rs:215-224), vertices may be dielectric (own spec).
// NODES: 0 = octant-ordered 112-byte pair nodes; 2 = quantised 32-byte pair nodes read with one 256-bit load (global-memory
// triangle scenes: SceneLayout::qnodes, pair_step_quant); 3 = NO tree (general scenes of a handful of primitives, SceneLayout::flat_scan):
// a trace burst is one straight scan over all records -- every lane reads the same record (broadcast loads), the primitive kind is
// warp-uniform, all 32 rays finish together: no box tests, no phase votes, no plane scan at retire.
template <class Space, bool STATS, int MODE, int BLOCK, int MINB, bool GEN = false, int NODES = 0>
__global__ void __launch_bounds__(BLOCK, MINB) render_wave_kernel(const RenderArgs a) {
    constexpr int P = RT_POOL_SLOTS;
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
    Space sp = SpaceMaker<Space>::make(a.blob, smem_base + stack_bytes);
#ifdef RT_DEBUG_BOUNDS
    sp.limit = a.debug_blob_limit ? a.debug_blob_limit : a.L.total_bytes;
#endif
    const SceneLayout& L = a.L;
    SmemStack st; st.init(smem_base + threadIdx.x * 4u, BLOCK * 4u, a.stack_entries);
    const uint32_t stack0 = smem_base + threadIdx.x * 4u;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned FULL = 0xffffffffu;
    Pool pool; pool.base = smem_base + stack_bytes + blob_bytes + warp * RT_POOL_WARP_BYTES;
    const uint2 key = make_uint2(a.seed_lo, a.seed_hi);
    const float3 bg = f3(a.bg[0], a.bg[1], a.bg[2]);
    const bool bg_nonzero = (a.bg[0] != 0.f) | (a.bg[1] != 0.f) | (a.bg[2] != 0.f);
    const size_t n_pix = (size_t)a.W * (size_t)a.H;
    const int burst_exit = a.burst_exit, node_min = a.node_min;
    const bool small_leaves = a.L.max_leaf <= 2;

    Counters cnt; cnt.node_tests = 0; cnt.tri_tests = 0; cnt.light_tri_tests = 0;
    unsigned long long c_samples = 0, c_segments = 0, c_vertices = 0, c_attempts = 0, c_cap = 0, c_nonfinite = 0;

    // ring queues: positions in [0, P), counts separate (all warp-uniform)
    int tq_head = 0, tq_n = 0, sq_head = 0, sq_n = P;
    for (int s = (int)lane; s < P; s += 32) { pool.sti(F_META, s, (int)pack_meta(ST_NEED_ITEM, 0, 0, 0u)); pool.qst(1, s, s); }
    __syncwarp();

    // traversal state of this lane's in-flight ray (the stack pointer lives in `st`)
    int slot = -1, cur = RT_CUR_DONE, hit_tri = -1;
    float t_best = RT_INF_F;
    RaySetup rs; rs.inv = f3(1.f, 1.f, 1.f); rs.od = f3(0.f, 0.f, 0.f); rs.ox = 0u; rs.oy = 32u; rs.oz = 64u;

    // box-pair step and leaf step of the lane's ray (callers predicate them)
    auto node_step = [&]() {
        if constexpr (NODES == 2) pair_step_quant(sp, L.qnodes, rs, t_best, cur, st);
        else pair_step<Space, true, IsSmem<Space>::value>(sp, L.nodes, rs, t_best, cur, st);   // shared-memory scenes always carry packed references
        if (STATS) cnt.node_tests += 2;
    };
    auto tri_rec = [&](int i) { return load_tri_test(sp, L.tri_t, i); };
    // (origin, direction) of the lane's ray from the registers the box tests use: d = 1/(1/d), o = (o/d) d (3 MUFU + 3 FMUL, 2^-22
    // relative) instead of six more shared-memory loads per leaf visit -- the kernel is shared-memory bound
    auto ray_of_lane = [&](float3& o, float3& d) {
        if constexpr (NODES == 2) ray_from_quant(rs, L.qorg, L.qcell, o, d);
        else { d = f3(fast_rcp(rs.inv.x), fast_rcp(rs.inv.y), fast_rcp(rs.inv.z)); o = rs.od * d; }
    };
    auto leaf_step = [&]() {
        const uint32_t code = (uint32_t)~cur;
        const int first = (int)(code >> 3), n = (int)(code & 7u) + 1;
        float3 o, d;
        ray_of_lane(o, d);
        const int skip_tri = pool.ldi(F_TRI, slot);
        if (GEN) {
            for (int i = first; i < first + n; ++i) {
                float t, u, v;
                const bool ok = prim_first_hit(load_prim(sp, L.prims, i), o, d, t, u, v);
                if (STATS) cnt.tri_tests += 1;
                if (ok && t < t_best && i != skip_tri) { t_best = t; hit_tri = i; pool.stf(F_U, slot, u); pool.stf(F_V, slot, v); }
            }
            cur = st.pop();
            return;
        }
        if (small_leaves) {                                                // leaves of 1-2 triangles: both tests side by side, one update
            const int second = first + (n > 1 ? 1 : 0);
            float t0, u0, v0, t1, u1, v1;
            const bool ok0 = tri_test(o, d, tri_rec(first), t0, u0, v0) && first != skip_tri;
            const bool ok1 = tri_test(o, d, tri_rec(second), t1, u1, v1) && second != skip_tri && n > 1;
            if (STATS) cnt.tri_tests += (unsigned long long)n;
            const bool take1 = ok1 && (!ok0 || t1 < t0);                   // strict <: the first triangle keeps exact ties (bvh.rs:269)
            const float tw = take1 ? t1 : t0;
            if ((ok0 || ok1) && tw < t_best) { t_best = tw; hit_tri = take1 ? second : first; pool.stf(F_U, slot, take1 ? u1 : u0); pool.stf(F_V, slot, take1 ? v1 : v0); }
            cur = st.pop();
            return;
        }
        for (int i = first; i < first + n; ++i) {
            float t, u, v;
            const bool ok = tri_test(o, d, tri_rec(i), t, u, v);
            if (STATS) cnt.tri_tests += 1;
            if (ok && t < t_best && i != skip_tri) { t_best = t; hit_tri = i; pool.stf(F_U, slot, u); pool.stf(F_V, slot, v); }
        }
        cur = st.pop();
    };

    for (;;) {
        // ---------------------------------------------------------------------------------- retire finished rays, refill from TQ
        {
            const bool finished = slot >= 0 && cur == RT_CUR_DONE;
            const unsigned fin = __ballot_sync(FULL, finished);
            if (fin != 0u) {
                if (finished) {
                    if (GEN && NODES != 3 && L.n_planes > 0) {                    // infinite primitives after the BVH (rendering.rs:215-224)
                        float3 o, d;
                        ray_of_lane(o, d);
                        const int skip_tri = pool.ldi(F_TRI, slot);
                        for (int i = L.n_tris; i < L.n_tris + L.n_planes; ++i) {
                            float t, u, v;
                            const bool ok = prim_first_hit(load_prim(sp, L.prims, i), o, d, t, u, v);
                            if (STATS) cnt.tri_tests += 1;
                            if (ok && t < t_best && i != skip_tri) { t_best = t; hit_tri = i; pool.stf(F_U, slot, u); pool.stf(F_V, slot, v); }
                        }
                    }
                    pool.sti(F_TRI, slot, hit_tri);
                    pool.qst(1, (sq_head + sq_n + __popc(fin & lt_mask)) & (P - 1), slot);
                    slot = -1;
                }
                sq_n += __popc(fin);
            }
            const unsigned idle = __ballot_sync(FULL, slot < 0);
            const int n_idle = __popc(idle);
            if (n_idle > 0 && tq_n > 0) {
                const int rank = __popc(idle & lt_mask);
                if (slot < 0 && rank < tq_n) {
                    slot = pool.qld(0, (tq_head + rank) & (P - 1));
                    if constexpr (NODES == 2) rs = ray_setup_quant(pool.ld3(F_OX, slot), pool.ld3(F_IX, slot), L.qorg, L.qcell);
                    else if constexpr (NODES != 3) { rs.inv = pool.ld3(F_IX, slot); rs.od = pool.ld3(F_OX, slot) * rs.inv; ray_octant(rs, L.nodes); }
                    t_best = RT_INF_F; hit_tri = -1;
                    cur = 0; st.reset(stack0);
                    if (STATS) ++c_segments;
                }
                const int took = min(n_idle, tq_n);
                tq_head = (tq_head + took) & (P - 1);
                tq_n -= took;
            }
            __syncwarp();
            // ------------------------------------------------------------------------------ shade round when TQ ran dry
            if (n_idle > 0 && tq_n == 0 && sq_n > 0 && __ballot_sync(FULL, slot < 0) != 0u) {
                const int n_take = min(32, sq_n);
                int s = 0, state = ST_NONE;
                uint32_t meta = 0;
                if ((int)lane < n_take) {
                    s = pool.qld(1, (sq_head + (int)lane) & (P - 1));
                    meta = (uint32_t)pool.ldi(F_META, s);
                    state = (int)(meta & 7u);
                }
                sq_head = (sq_head + n_take) & (P - 1);
                sq_n -= n_take;
                int depth = (int)((meta >> 3) & 0xffu), attempt = (int)((meta >> 11) & 0x7fu);
                uint32_t call = meta >> 18;
                int tri = -1, mat_id = 0;
                float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f);
                float3 T = f3(1.f, 1.f, 1.f);

                // (1) what the nearest-hit query returned: get_ray_color (rendering.rs:86-127) up to the sampling loop
                bool ends = state == ST_ENDED;
                if (state == ST_TRACE || state == ST_RETRY) {
                    tri = pool.ldi(F_TRI, s);
                    T = pool.ld3(F_TX, s);
                    if (tri < 0) {                                               // :125
                        if (bg_nonzero) pool.st3(F_AX, s, pool.ld3(F_AX, s) + T * bg);
                        ends = true;
                    } else {
                        RT_BOUNDS(tri < L.n_tris + L.n_planes || (L.n_tris == 0 && tri == 0), RT_BOUNDS_PRIM);
                        n0 = sp.ld4(L.sh_n0 + (uint32_t)tri * 16u);
                        mat_id = __float_as_int(n0.w);
                        if (attempt == 0) {
                            const float3 em = f3(sp.ld4(L.mat1 + (uint32_t)mat_id * 16u));
                            if ((em.x != 0.f) | (em.y != 0.f) | (em.z != 0.f)) pool.st3(F_AX, s, pool.ld3(F_AX, s) + T * em);   // :99
                            if (--depth <= 0) ends = true;                       // :93-95
                            else if (STATS) ++c_vertices;
                        }
                    }
                }
                // (2) path bookkeeping for the slots whose path ended: next sample of the item, or store the item and
                //     fetch a new one (warp-aggregated atomic on the frame's work counter)
                uint32_t xy = 0;
                int s_this = 0;
                if (state != ST_NONE && state != ST_NEED_ITEM) { xy = (uint32_t)pool.ldi(F_PIX, s); s_this = pool.ldi(F_S, s); }
                if (ends) {
                    const int chunk = pool.ldi(F_CHUNK, s);
                    const int s_stop = a.s_begin + __ldg(a.chunk_begin + chunk + 1);
                    if (++s_this < s_stop) state = ST_GEN;
                    else {
                        const uint32_t pix = (xy >> 16) * (uint32_t)a.W + (xy & 0xffffu);
                        const float3 acc = pool.ld3(F_AX, s);
                        const float n_done = (float)(__ldg(a.chunk_begin + chunk + 1) - __ldg(a.chunk_begin + chunk));
                        RT_BOUNDS(chunk >= 0 && chunk < a.n_chunks, RT_BOUNDS_CHUNK);
                        RT_BOUNDS((size_t)pix < n_pix, RT_BOUNDS_LAYER);
                        a.layers[(size_t)chunk * n_pix + pix] = make_float4(acc.x, acc.y, acc.z, n_done);
                        state = ST_NEED_ITEM;
                    }
                }
                for (;;) {
                    const bool need = state == ST_NEED_ITEM;
                    const unsigned m = __ballot_sync(FULL, need);
                    if (m == 0u) break;
                    const int leader = __ffs(m) - 1;
                    unsigned int base = 0;
                    if ((int)lane == leader) base = atomicAdd(a.work_counter, (unsigned int)__popc(m));
                    base = __shfl_sync(FULL, base, leader);
                    if (need) {
                        const uint32_t idx = base + (uint32_t)__popc(m & lt_mask);
                        if (idx >= a.total_items) state = ST_EXHAUSTED;
                        else {
                            const uint32_t chunk = idx / a.n_pix_items;
                            const uint32_t rr = idx - chunk * a.n_pix_items;
                            const uint32_t tile = rr >> 5, w = rr & 31u;
                            const uint32_t ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
                            const uint32_t px = tx * 8u + (w & 7u), py = ty * 4u + (w >> 3);
                            if ((int)px < a.W && (int)py < a.H && tile % a.shard_count == a.shard_index) {
                                xy = (py << 16) | px;
                                s_this = a.s_begin + __ldg(a.chunk_begin + chunk);
                                pool.sti(F_PIX, s, (int)xy);
                                pool.sti(F_CHUNK, s, (int)chunk);
                                pool.st3(F_AX, s, f3(0.f, 0.f, 0.f));
                                state = ST_GEN;
                            }
                        }
                    }
                }
                if (state == ST_GEN) {
                    pool.sti(F_S, s, s_this);
                    call = 0u; attempt = 0; depth = a.ray_depth;                    // <= 255, checked on the host
                    if (STATS) ++c_samples;
                }
                // (3) ONE Philox call per live slot: the camera jitter of a new path (call 0) or the next attempt of the
                //     rejection loop (call >= 1)
                const bool live = state == ST_GEN || state == ST_TRACE || state == ST_RETRY;
                uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
                const uint32_t pix = (xy >> 16) * (uint32_t)a.W + (xy & 0xffffu);
                if (live) { rnd = philox4x32_10(make_uint4(pix, (uint32_t)s_this, call, RT_PHILOX_TAG), key); call = (call + 1u) & 0x3fffu; }
                // (4) the new ray: camera ray (rendering.rs:71-84), or one attempt of the rejection loop (:102-110)
                float3 ro = f3(0.f, 0.f, 0.f), rd = f3(0.f, 0.f, 1.f);
                int skip = -1;
                bool accepted = false;
                if (state == ST_GEN) {
                    camera_ray(a.cam, a.W, a.H, (int)(xy & 0xffffu), (int)(xy >> 16), u01(rnd.x), u01(rnd.y), ro, rd);
                    T = f3(1.f, 1.f, 1.f);
                    accepted = true;
                } else if (live) {
                    const Material mat = load_material(sp, L, mat_id);
                    const float3 iv = pool.ld3(F_IX, s);                          // the ray direction is not stored: d = 1 / (1/d) (2^-22, shared-memory bound)
                    const float3 din = f3(fast_rcp(iv.x), fast_rcp(iv.y), fast_rcp(iv.z));
                    const float hu = pool.ldf(F_U, s), hv = pool.ldf(F_V, s);
                    const Vertex V = hit_vertex<Space, GEN>(sp, L, tri, n0, GEN ? pool.ld3(F_OX, s) : f3(0.f, 0.f, 0.f), din, hu, hv);
                    const float3 n = V.n, ns = V.ns, Pt = V.P;
                    const float3 v = -din;
                    const float nv = dot(n, v);
                    bool dielectric = false;
                    float ior = 1.0f;
                    if (GEN && L.has_dielectric) {
                        const float4 m2 = sp.ld4(L.mat2 + (uint32_t)mat_id * 16u);
                        dielectric = __float_as_int(m2.y) == RT_MAT_DIELECTRIC; ior = m2.x;
                    }
                    if (GEN && dielectric) {
                        float3 l;
                        const bool refract = dielectric_sample(n, v, nv, ior, V.outer, u01(rnd.x), l);
                        if (STATS) ++c_attempts;
                        if (refract) { ro = fma3(din, RT_EPS_F, V.surf); if (V.outer) T = T * mat.base; skip = V.flat ? tri : -1; }
                        else { ro = Pt; skip = (V.flat || V.outer) ? tri : -1; }
                        rd = l; attempt = 0; accepted = true;
                    } else {
                    const float alpha = mat.roughness * mat.roughness, alpha2 = alpha * alpha;
                    const float g1v = ggx_g1(nv, alpha2);
                    float3 l; DirTerms terms;
                    const float pdf = mix_sample_and_pdf<Space, STATS, GEN>(sp, L, st, a.n_comp, a.inv_n_comp, Pt, n, v, nv, alpha, alpha2, g1v, rnd, l, terms, cnt);
                    ++attempt;
                    if (STATS) ++c_attempts;
                    if (pdf > 0.0f && dot(l, ns) > 0.0f) {                       // :107
                        const float d_chi = terms.nh > 0.0f ? terms.d_nochi : 0.0f;
                        const float3 f = brdf_eval(mat, d_chi, ggx_g1(terms.nl, alpha2), g1v, terms.nl, nv, terms.hl);
                        T = T * f * (terms.nl * fast_rcp(pdf));                  // :122
                        if (!finite3(T)) { if (STATS) ++c_nonfinite; state = ST_ENDED; }
                        else { ro = Pt; rd = l; skip = (terms.nl > 0.0f && (V.flat || V.outer)) ? tri : -1; attempt = 0; accepted = true; }
                    } else if (attempt >= a.max_attempts || attempt >= 127) {
                        if (STATS) ++c_cap;
                        state = ST_ENDED;                                        // the path ends in the next round
                    } else {
                        state = ST_RETRY;
                    }
                    }
                }
                if (accepted) {
                    pool.st3(F_OX, s, ro); pool.st3(F_IX, s, safe_inv_dir(rd)); pool.st3(F_TX, s, T);
                    pool.sti(F_TRI, s, skip);
                    state = ST_TRACE;
                }
                if (state != ST_NONE) pool.sti(F_META, s, (int)pack_meta(state, depth, attempt, call));
                const unsigned to_tq = __ballot_sync(FULL, state == ST_TRACE);
                if (state == ST_TRACE) pool.qst(0, (tq_head + tq_n + __popc(to_tq & lt_mask)) & (P - 1), s);
                tq_n += __popc(to_tq);
                const bool again = state == ST_RETRY || state == ST_ENDED;       // rejected attempt / late path end: a later round
                const unsigned to_sq = __ballot_sync(FULL, again);
                if (again) pool.qst(1, (sq_head + sq_n + __popc(to_sq & lt_mask)) & (P - 1), s);
                sq_n += __popc(to_sq);
                __syncwarp();
                continue;
            }
        }
        if (__ballot_sync(FULL, slot >= 0) == 0u) break;                         // nothing in flight, nothing queued: pool exhausted

        // ---------------------------------------------------------------------------------- trace burst
        if constexpr (NODES == 3) {
            // intersect_ray_with_scene (rendering.rs:201-226) as one scan: finite primitives in BVH order, then the planes, strict `<`
            if (slot >= 0) {
                const float3 o = pool.ld3(F_OX, slot), iv = pool.ld3(F_IX, slot);
                const float3 d = f3(fast_rcp(iv.x), fast_rcp(iv.y), fast_rcp(iv.z));
                const int skip_tri = pool.ldi(F_TRI, slot);
                float hu = 0.f, hv = 0.f;
                for (int i = 0; i < L.n_tris + L.n_planes; ++i) {
                    float t, u, v;
                    const bool ok = prim_first_hit(load_prim(sp, L.prims, i), o, d, t, u, v);
                    if (ok && t < t_best && i != skip_tri) { t_best = t; hit_tri = i; hu = u; hv = v; }
                }
                if (STATS) cnt.tri_tests += (unsigned long long)(L.n_tris + L.n_planes);
                pool.stf(F_U, slot, hu); pool.stf(F_V, slot, hv);
                cur = RT_CUR_DONE;
            }
            continue;
        }
        if (MODE == 0) {
            for (;;) {
                while (cur >= 0) node_step();
                if (cur != RT_CUR_DONE) leaf_step();
                if (__popc(__ballot_sync(FULL, cur == RT_CUR_DONE)) >= burst_exit) break;
            }
        } else {
            // phased burst: box-pair steps run while at least `nmin` lanes want one; otherwise the lanes that sit on a leaf
            // test its triangles; the minority keeps its state and waits for its phase.  nmin <= live lanes, so a burst
            // always advances some ray; it ends when `burst_exit` lanes are finished (or idle).
            const int nmin = min(node_min, __popc(__ballot_sync(FULL, slot >= 0)));
            for (;;) {
                for (;;) {
                    if (__popc(__ballot_sync(FULL, cur >= 0)) < nmin) break;
                    if (cur >= 0) node_step();
#if RT_NODE_UNROLL >= 2
                    if (cur >= 0) node_step();
#endif
#if RT_NODE_UNROLL >= 3
                    if (cur >= 0) node_step();
#endif
#if RT_NODE_UNROLL >= 4
                    if (cur >= 0) node_step();
#endif
#if RT_NODE_UNROLL >= 6
                    if (cur >= 0) node_step();
                    if (cur >= 0) node_step();
#endif
                }
                const bool at_leaf = cur < 0 && cur != RT_CUR_DONE;
                if (__ballot_sync(FULL, at_leaf) == 0u) break;                   // fewer than nmin box-pair lanes, the rest finished
                if (at_leaf) leaf_step();
                if (__popc(__ballot_sync(FULL, cur == RT_CUR_DONE)) >= burst_exit) break;
            }
        }
    }
    if (STATS) {
        atomicAdd(a.stats + RT_STAT_SAMPLES, c_samples); atomicAdd(a.stats + RT_STAT_SEGMENTS, c_segments);
        atomicAdd(a.stats + RT_STAT_VERTICES, c_vertices); atomicAdd(a.stats + RT_STAT_ATTEMPTS, c_attempts);
        atomicAdd(a.stats + RT_STAT_NODE_TESTS, cnt.node_tests); atomicAdd(a.stats + RT_STAT_TRI_TESTS, cnt.tri_tests);
        atomicAdd(a.stats + RT_STAT_LIGHT_TRI_TESTS, cnt.light_tri_tests); atomicAdd(a.stats + RT_STAT_CAP_HITS, c_cap);
        atomicAdd(a.stats + RT_STAT_NONFINITE, c_nonfinite);
    }
}

template <class Space, bool STATS, int MODE, int BLOCK, int MINB, bool GEN = false, int NODES = 0>
static cudaError_t launch_wave_t(const RenderArgs& a, int device_sms, cudaStream_t stream, KernelInfo* info, bool launch, int* lanes) {
    const uint32_t smem = a.stack_entries * BLOCK * 4u + (IsSmem<Space>::value ? a.L.total_bytes : 0u) + (BLOCK / 32) * RT_POOL_WARP_BYTES;
    auto kern = render_wave_kernel<Space, STATS, MODE, BLOCK, MINB, GEN, NODES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const int grid = per_sm * device_sms;
    if (lanes) *lanes = grid * (BLOCK / 32) * RT_POOL_SLOTS;       // resident path slots
    if (info) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kern);
        info->block = BLOCK; info->blocks_per_sm = per_sm; info->regs = fa.numRegs; info->smem_bytes = (int)smem; info->grid = grid;
    }
    if (!launch) return cudaSuccess;
    kern<<<grid, BLOCK, smem, stream>>>(a);
    return cudaGetLastError();
}
