// rt_kernels.h -- host-callable launchers of the sm_100a kernels (rt_kernels.cu).  C++ linkage, internal to the
// library; the public surface is include/rt_api.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_device.cuh"

namespace rtd {

struct Camera {
    float pos[3], right[3], up[3], fwd[3];
    float tan_x, tan_y;  // tan(fov_x/2), tan(fov_y/2) evaluated in f64 on the host (rendering.rs:76-77)
    float two_over_w, two_over_h;  // 2/width, 2/height (rendering.rs:74-75 without a division per sample)
};

#define RT_MAX_CHUNKS 65536
struct RenderArgs {
    const char* blob;            // device scene blob (global memory)
    SceneLayout L;
    Camera cam;
    float bg[3];
    int32_t W, H, ray_depth, max_attempts, n_comp;
    float inv_n_comp;            // 1 / n_comp (MixDistribution::pdf, distributions.rs:194-201)
    int32_t s_begin, s_end, n_chunks;
    // Sample chunks: chunk k renders samples [s_begin + chunk_begin[k], s_begin + chunk_begin[k+1]) of a pixel into layer k.  The first
    // chunks are large and equal, the last ones shrink geometrically: work items are handed out in chunk order, so the items that are
    // still running when the frame's work counter runs dry are short ones and the tail of the persistent kernel stays short.
    const int32_t* chunk_begin;  // device array of n_chunks + 1 entries (read once per work item)
    uint32_t tiles_x, n_pix_items, total_items;
    uint32_t shard_index, shard_count;   // interleaved tile sharding: this launch owns tiles t with t % shard_count == shard_index (count >= 1)
    uint32_t stack_entries;
    int32_t node_min;            // v3 phased bursts: box-pair steps run while at least this many lanes want one
    int32_t burst_exit;          // v3: a trace burst ends when this many lanes of the warp have finished their ray (1..32)
    uint32_t debug_blob_limit;   // RT_DEBUG_BOUNDS builds: blob bytes the loads are checked against (0 = L.total_bytes; RT_DEBUG_BLOB_LIMIT lowers it to prove the instrument fires)
    uint32_t seed_lo, seed_hi;
    float4* layers;              // n_chunks x W*H  (rgb sums, sample count)
    unsigned int* work_counter;  // zeroed before launch
    unsigned long long* stats;   // RT_N_STATS counters (stats kernels only)
};

enum { RT_STAT_SAMPLES = 0, RT_STAT_SEGMENTS, RT_STAT_VERTICES, RT_STAT_ATTEMPTS, RT_STAT_NODE_TESTS, RT_STAT_TRI_TESTS,
       RT_STAT_LIGHT_TRI_TESTS, RT_STAT_CAP_HITS, RT_STAT_NONFINITE, RT_N_STATS };

struct KernelInfo { int block, blocks_per_sm, regs, smem_bytes, grid; };

// Launch the persistent render kernel.  variant: 1 = v1 per-lane megakernel, 2 / 3 = v3 warp-local wavefront (64 path slots per warp) with
// while-while / phased trace bursts; cfg picks the v3 build (block size x min blocks per SM).  use_smem: stage the blob in shared memory.  Returns cudaError_t.
cudaError_t launch_render(const RenderArgs& a, int variant, int cfg, bool use_smem, bool stats, int device_sms, cudaStream_t stream, KernelInfo* info);
// How many lanes the render kernel keeps resident (grid * block) -- used to size the sample chunks.
cudaError_t render_resident_lanes(int variant, int cfg, bool use_smem, bool stats, bool general, int node_format, uint32_t blob_bytes, uint32_t stack_entries, int device_sms, int* lanes);

cudaError_t launch_black_layer(float4* layer, int W, int H, uint32_t tiles_x, uint32_t shard_index, uint32_t shard_count, float n_samples, cudaStream_t stream);
cudaError_t launch_sum_layers(const float4* layers, int n_layers, size_t n_pix, float4* accum, bool add, cudaStream_t stream);
cudaError_t launch_resolve_u8(const float4* accum, size_t n_pix, uint8_t* rgb, cudaStream_t stream);
cudaError_t launch_resolve_linear(const float4* accum, size_t n_pix, float* rgb, cudaStream_t stream);

cudaError_t launch_trace_rays(const char* blob, const SceneLayout& L, const double* tri_d, uint32_t stack_entries, const double* rays, long long n,
                              bool f64, int32_t* tri_id, double* t, double* hits, cudaStream_t stream);
cudaError_t launch_primary_rays(const Camera& cam, int W, int H, const int32_t* xy, const double* xi, long long n, double* rays, cudaStream_t stream);
cudaError_t launch_eval(const char* blob, const SceneLayout& L, uint32_t stack_entries, int fn, const float* in, long long n, float* out, cudaStream_t stream);
cudaError_t launch_ffma(int blocks, int threads, int iters, float* sink, cudaStream_t stream);

cudaError_t read_bounds_violations(unsigned long long* out, int reset);   // RT_DEBUG_BOUNDS builds only (rt_device.cuh)
int eval_in_width(int fn);
int eval_out_width(int fn);

}  // namespace rtd
