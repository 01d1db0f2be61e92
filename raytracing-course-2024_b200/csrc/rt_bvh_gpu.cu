// rt_bvh_gpu.cu -- GPU builder for the device BVH: takes the place of create_bvh_tree (bvh.rs:26-144)
// when scene load time matters more than tree quality.  Produces the same FlatBvh (host_scene.h) as the host SAH builder
// (bvh_builder.cpp), so everything downstream -- validate_flat_bvh (bvh.rs:299-322 restated), the octant-ordered node
// packing, the kernels -- is shared.  Node order and shape are not observable through render_scene; nearest-hit results are
// identical to any other valid tree except at exact ties.
//
// Pipeline (all on the device, one stream):
//   1. per triangle: the reference's EPS-padded box (aabb.rs:53-65) rounded outward to f32 + centroid; scene bounds by
//      block reduction + ordered-int atomics;
//   2. 30-bit Morton code of the centroid, made unique by appending the triangle's position (64-bit key); CUB radix sort;
//   3. Karras 2012 ("Maximizing parallelism in the construction of BVHs, octrees and k-d trees"): every inner node finds
//      its key range and split by longest common prefixes, fully in parallel;
//   4. bottom-up refit with one atomic visit counter per inner node: boxes and subtree sizes;
//   5. collapse: an inner node whose subtree holds <= max_leaf triangles (while its parent's holds more) becomes a leaf over
//      its contiguous key range; surviving inner nodes are compacted by a prefix sum (root stays node 0) and emitted as
//      child-pair nodes (boxes of both children + two references).
#include <cub/cub.cuh>

#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "host_scene.h"

namespace rtb {
namespace {

#define GPU_TRY(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess) { if (err) *err = std::string(#expr) + ": " + cudaGetErrorString(e__); ok = false; goto done; } \
    } while (0)

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }   // monotonic float -> int
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct Box6 { float mn[3], mx[3]; };

// 1. boxes, centroids, scene bounds (of the centroids) ----------------------------------------------------------------
__global__ void k_boxes(const double* __restrict__ tri_v, const int32_t* __restrict__ ids, int n, Box6* __restrict__ boxes, float3* __restrict__ cen, int* __restrict__ bounds) {
    const double EPS = 0.00001;                                        // geometry.rs:49
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {0.f, 0.f, 0.f};
    if (k < n) {
        const double* v = tri_v + (size_t)ids[k] * 9;
        Box6 b;
        for (int a = 0; a < 3; ++a) {
            const double lo = fmin(fmin(v[a], v[3 + a]), v[6 + a]) - EPS, hi = fmax(fmax(v[a], v[3 + a]), v[6 + a]) + EPS;   // aabb.rs:53-65
            b.mn[a] = __double2float_rd(lo); b.mx[a] = __double2float_ru(hi);
            c[a] = (float)(0.5 * (lo + hi));
        }
        boxes[k] = b;
        cen[k] = make_float3(c[0], c[1], c[2]);
    }
    // block reduction of the centroid bounds, then one atomic per block and component
    __shared__ int smn[3][8], smx[3][8];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int a = 0; a < 3; ++a) {
        int lo = k < n ? f2ord(c[a]) : 0x7fffffff, hi = k < n ? f2ord(c[a]) : (int)0x80000000;
        for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if (lane == 0) { smn[a][warp] = lo; smx[a][warp] = hi; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int lo = 0x7fffffff, hi = (int)0x80000000;
        for (unsigned w = 0; w < blockDim.x / 32u; ++w) { lo = min(lo, smn[threadIdx.x][w]); hi = max(hi, smx[threadIdx.x][w]); }
        atomicMin(bounds + threadIdx.x, lo); atomicMax(bounds + 3 + threadIdx.x, hi);
    }
}

// 2. Morton keys ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expand10(uint32_t v) {            // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__global__ void k_morton(const float3* __restrict__ cen, int n, const int* __restrict__ bounds, unsigned long long* __restrict__ keys) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float lo[3] = {ord2f(bounds[0]), ord2f(bounds[1]), ord2f(bounds[2])}, hi[3] = {ord2f(bounds[3]), ord2f(bounds[4]), ord2f(bounds[5])};
    const float c[3] = {cen[k].x, cen[k].y, cen[k].z};
    uint32_t q[3];
    for (int a = 0; a < 3; ++a) {
        const float ext = hi[a] - lo[a];
        const float t = ext > 0.f ? (c[a] - lo[a]) / ext : 0.f;
        q[a] = (uint32_t)fminf(fmaxf(t * 1024.f, 0.f), 1023.f);
    }
    const uint32_t m = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    keys[k] = ((unsigned long long)m << 32) | (unsigned long long)(uint32_t)k;   // unique keys: position as tie-breaker
}

// 3. Karras hierarchy ------------------------------------------------------------------------------------------------------
// Node numbering: inner nodes 0 .. n-2, leaves n-1 .. 2n-2 (leaf j = sorted position j).  parent[] covers both.
__device__ __forceinline__ int lcp(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(keys[i] ^ keys[j]));                    // keys are unique -> never 64
}
__global__ void k_hierarchy(const unsigned long long* __restrict__ keys, int n, int2* __restrict__ children, int2* __restrict__ range, int* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = lcp(keys, n, i, i + 1) > lcp(keys, n, i, i - 1) ? 1 : -1;
    const int dmin = lcp(keys, n, i, i - d);
    int lmax = 2;
    while (lcp(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t > 0; t >>= 1) if (lcp(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lcp(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {                   // binary search of the split with ceil halving
        if (lcp(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = lo == gamma ? (n - 1) + gamma : gamma;            // leaf or inner
    const int right = hi == gamma + 1 ? (n - 1) + gamma + 1 : gamma + 1;
    children[i] = make_int2(left, right);
    range[i] = make_int2(lo, hi);
    parent[left] = i; parent[right] = i;
    if (i == 0) parent[0] = -1;
}

// 4. refit ---------------------------------------------------------------------------------------------------------------------
__global__ void k_refit(const unsigned long long* __restrict__ keys, const Box6* __restrict__ tri_boxes, int n, const int2* __restrict__ children,
                        const int* __restrict__ parent, Box6* __restrict__ node_boxes /* 2n-1 */, int* __restrict__ visits) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    node_boxes[(n - 1) + j] = tri_boxes[(uint32_t)(keys[j] & 0xffffffffull)];
    __threadfence();
    int cur = parent[(n - 1) + j];
    while (cur >= 0) {
        if (atomicAdd(visits + cur, 1) == 0) return;                   // first arrival: the sibling subtree is not done yet
        const int2 ch = children[cur];
        const float* pa = reinterpret_cast<const float*>(node_boxes + ch.x);
        const float* pb = reinterpret_cast<const float*>(node_boxes + ch.y);
        Box6 u;                                                        // __ldcg: the children were written by other SMs, bypass L1
        for (int k = 0; k < 3; ++k) { u.mn[k] = fminf(__ldcg(pa + k), __ldcg(pb + k)); u.mx[k] = fmaxf(__ldcg(pa + 3 + k), __ldcg(pb + 3 + k)); }
        node_boxes[cur] = u;
        __threadfence();
        cur = parent[cur];
    }
}

// 5. collapse + emit ----------------------------------------------------------------------------------------------------------
// alive[i] = 1 iff inner node i survives as a pair node: its own subtree holds more than max_leaf triangles.
__global__ void k_alive(const int2* __restrict__ range, int n, int max_leaf, int* __restrict__ alive) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    alive[i] = (range[i].y - range[i].x + 1) > max_leaf ? 1 : 0;
}
__global__ void k_emit(const int2* __restrict__ children, const int2* __restrict__ range, const Box6* __restrict__ node_boxes, const int* __restrict__ alive,
                       const int* __restrict__ new_index, int n, float4* __restrict__ box_a, float4* __restrict__ box_b, float4* __restrict__ box_c, int2* __restrict__ child) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !alive[i]) return;
    const int2 ch = children[i];
    int ref[2];
    const int c2[2] = {ch.x, ch.y};
    for (int k = 0; k < 2; ++k) {
        const int c = c2[k];
        if (c >= n - 1) ref[k] = ~(((c - (n - 1)) << 3) | 0);                                     // single triangle
        else if (alive[c]) ref[k] = new_index[c];
        else ref[k] = ~((range[c].x << 3) | (range[c].y - range[c].x));                              // collapsed subtree: its key range
    }
    const Box6 a = node_boxes[ch.x], b = node_boxes[ch.y];
    const int o = new_index[i];
    box_a[o] = make_float4(a.mn[0], a.mx[0], a.mn[1], a.mx[1]);
    box_b[o] = make_float4(b.mn[0], b.mx[0], b.mn[1], b.mx[1]);
    box_c[o] = make_float4(a.mn[2], a.mx[2], b.mn[2], b.mx[2]);
    child[o] = make_int2(ref[0], ref[1]);
}

}  // namespace

bool build_bvh_gpu(const double* tri_v, int n_total_tris, const std::vector<int32_t>& ids, const BvhBuildParams& p, int device, FlatBvh* out, double* build_ms,
                   std::string* err) {
    const int n = (int)ids.size();
    const int max_leaf = std::max(1, std::min(8, p.max_leaf_size));
    if (n <= 2 * max_leaf) { if (err) *err = "too few triangles for the GPU builder"; return false; }
    bool ok = true;
    double* d_v = nullptr; int32_t* d_ids = nullptr; Box6* d_tb = nullptr; float3* d_cen = nullptr; int* d_bounds = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr; void* d_tmp = nullptr; size_t tmp_bytes = 0, scan_bytes = 0;
    int2 *d_children = nullptr, *d_range = nullptr; int *d_parent = nullptr, *d_visits = nullptr, *d_alive = nullptr, *d_newidx = nullptr;
    Box6* d_nb = nullptr; float4 *d_a = nullptr, *d_b = nullptr, *d_c = nullptr; int2* d_child = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const int B = 256, G = (n + B - 1) / B;
    int n_alive = 0, last_alive = 0, last_idx = 0;
    std::vector<unsigned long long> keys_h;
    GPU_TRY(cudaSetDevice(device));
    GPU_TRY(cudaMalloc((void**)&d_v, (size_t)n_total_tris * 9 * sizeof(double)));
    GPU_TRY(cudaMalloc((void**)&d_ids, (size_t)n * sizeof(int32_t)));
    GPU_TRY(cudaMalloc((void**)&d_tb, (size_t)n * sizeof(Box6)));
    GPU_TRY(cudaMalloc((void**)&d_cen, (size_t)n * sizeof(float3)));
    GPU_TRY(cudaMalloc((void**)&d_bounds, 6 * sizeof(int)));
    GPU_TRY(cudaMalloc((void**)&d_keys, (size_t)n * 8));
    GPU_TRY(cudaMalloc((void**)&d_keys2, (size_t)n * 8));
    GPU_TRY(cudaMalloc((void**)&d_children, (size_t)n * sizeof(int2)));
    GPU_TRY(cudaMalloc((void**)&d_range, (size_t)n * sizeof(int2)));
    GPU_TRY(cudaMalloc((void**)&d_parent, (size_t)(2 * n) * sizeof(int)));
    GPU_TRY(cudaMalloc((void**)&d_visits, (size_t)n * sizeof(int)));
    GPU_TRY(cudaMalloc((void**)&d_alive, (size_t)n * sizeof(int)));
    GPU_TRY(cudaMalloc((void**)&d_newidx, (size_t)n * sizeof(int)));
    GPU_TRY(cudaMalloc((void**)&d_nb, (size_t)(2 * n) * sizeof(Box6)));
    GPU_TRY(cudaMalloc((void**)&d_a, (size_t)n * sizeof(float4)));
    GPU_TRY(cudaMalloc((void**)&d_b, (size_t)n * sizeof(float4)));
    GPU_TRY(cudaMalloc((void**)&d_c, (size_t)n * sizeof(float4)));
    GPU_TRY(cudaMalloc((void**)&d_child, (size_t)n * sizeof(int2)));
    GPU_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_keys2, n));
    GPU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_alive, d_newidx, n - 1));
    tmp_bytes = std::max(tmp_bytes, scan_bytes);
    GPU_TRY(cudaMalloc(&d_tmp, tmp_bytes));
    GPU_TRY(cudaMemcpy(d_v, tri_v, (size_t)n_total_tris * 9 * sizeof(double), cudaMemcpyHostToDevice));
    GPU_TRY(cudaMemcpy(d_ids, ids.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    GPU_TRY(cudaEventCreate(&e0)); GPU_TRY(cudaEventCreate(&e1));
    {
        const int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
        GPU_TRY(cudaMemcpy(d_bounds, init, sizeof(init), cudaMemcpyHostToDevice));
    }
    GPU_TRY(cudaEventRecord(e0));
    k_boxes<<<G, B>>>(d_v, d_ids, n, d_tb, d_cen, d_bounds);
    k_morton<<<G, B>>>(d_cen, n, d_bounds, d_keys);
    GPU_TRY(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_keys2, n));
    GPU_TRY(cudaMemsetAsync(d_visits, 0, (size_t)n * sizeof(int)));
    k_hierarchy<<<G, B>>>(d_keys2, n, d_children, d_range, d_parent);
    k_refit<<<G, B>>>(d_keys2, d_tb, n, d_children, d_parent, d_nb, d_visits);
    k_alive<<<G, B>>>(d_range, n, max_leaf, d_alive);
    GPU_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_alive, d_newidx, n - 1));
    k_emit<<<G, B>>>(d_children, d_range, d_nb, d_alive, d_newidx, n, d_a, d_b, d_c, d_child);
    GPU_TRY(cudaEventRecord(e1));
    GPU_TRY(cudaEventSynchronize(e1));
    GPU_TRY(cudaGetLastError());
    if (build_ms) { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *build_ms = ms; }
    GPU_TRY(cudaMemcpy(&last_alive, d_alive + (n - 2), sizeof(int), cudaMemcpyDeviceToHost));
    GPU_TRY(cudaMemcpy(&last_idx, d_newidx + (n - 2), sizeof(int), cudaMemcpyDeviceToHost));
    n_alive = last_idx + last_alive;
    out->n_nodes = n_alive;
    out->box_a.resize((size_t)n_alive * 4); out->box_b.resize((size_t)n_alive * 4); out->box_c.resize((size_t)n_alive * 4); out->child.resize((size_t)n_alive * 2);
    GPU_TRY(cudaMemcpy(out->box_a.data(), d_a, (size_t)n_alive * sizeof(float4), cudaMemcpyDeviceToHost));
    GPU_TRY(cudaMemcpy(out->box_b.data(), d_b, (size_t)n_alive * sizeof(float4), cudaMemcpyDeviceToHost));
    GPU_TRY(cudaMemcpy(out->box_c.data(), d_c, (size_t)n_alive * sizeof(float4), cudaMemcpyDeviceToHost));
    GPU_TRY(cudaMemcpy(out->child.data(), d_child, (size_t)n_alive * sizeof(int2), cudaMemcpyDeviceToHost));
    keys_h.resize((size_t)n);
    GPU_TRY(cudaMemcpy(keys_h.data(), d_keys2, (size_t)n * 8, cudaMemcpyDeviceToHost));
    out->tri_order.resize((size_t)n);
    for (int k = 0; k < n; ++k) out->tri_order[(size_t)k] = ids[(size_t)(keys_h[(size_t)k] & 0xffffffffull)];
    {   // depth, leaves, largest leaf: one pass over the emitted tree (node 0 = root)
        int depth = 0, leaves = 0, mx = 0;
        std::vector<std::pair<int, int>> stack;
        stack.push_back({0, 1});
        while (!stack.empty()) {
            const std::pair<int, int> t = stack.back(); stack.pop_back();
            depth = std::max(depth, t.second);
            for (int c = 0; c < 2; ++c) {
                const int r = out->child[(size_t)t.first * 2 + (size_t)c];
                if (r >= 0) stack.push_back({r, t.second + 1});
                else { ++leaves; mx = std::max(mx, (int)((uint32_t)~r & 7u) + 1); }
            }
        }
        out->depth = depth; out->n_leaves = leaves; out->max_leaf = mx;
    }
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_v); cudaFree(d_ids); cudaFree(d_tb); cudaFree(d_cen); cudaFree(d_bounds); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_tmp);
    cudaFree(d_children); cudaFree(d_range); cudaFree(d_parent); cudaFree(d_visits); cudaFree(d_alive); cudaFree(d_newidx); cudaFree(d_nb);
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_c); cudaFree(d_child);
    return ok;
}

}  // namespace rtb
