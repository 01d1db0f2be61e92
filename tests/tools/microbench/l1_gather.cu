// Micro-benchmark: what does one SM's L1 sustain for DIVERGENT record gathers (every lane its own address), by load width and by
// where the table lives (L1-resident, L2-resident)?  Prints lane-loads per clock per SM and bytes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_gather l1_gather.cu && ./l1_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>   // 0: 1 x LDG.128, 1: 2 x LDG.128 (same sector), 2: 1 x LDG.256, 3: 4 x LDG.128 (64 B), 4: 2 x LDG.256 (64 B), 5: LDG.64, 6: 3 x LDG.128 + LDG.64 spread over 112 B
__global__ void __launch_bounds__(352, 2) gather(const char* __restrict__ table, uint32_t mask, int iters, float* out) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const char* p = table + (size_t)((x >> 8) & mask) * 128u;
        float4 a, b, c, d;
        if (MODE == 0) { asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p)); acc += a.x + a.w; }
        if (MODE == 1) { asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
                         asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 16)); acc += a.x + b.w; }
        if (MODE == 2) { asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p)); acc += a.x + b.w; }
        if (MODE == 3) { asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
                         asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 16));
                         asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(p + 32));
                         asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(p + 48)); acc += a.x + b.w + c.y + d.z; }
        if (MODE == 4) { asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
                         asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w), "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(p + 32)); acc += a.x + b.w + c.y + d.z; }
        if (MODE == 5) { float2 e; asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(e.x), "=f"(e.y) : "l"(p)); acc += e.x + e.y; }
        if (MODE == 6) { float2 e;
                         asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p + (x & 16u)));
                         asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 32 + ((x >> 1) & 16u)));
                         asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(p + 64 + ((x >> 2) & 16u)));
                         asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(e.x), "=f"(e.y) : "l"(p + 96)); acc += a.x + b.w + c.y + e.x; }
    }
    if (acc == 123.456f) out[0] = acc;
}
template <int MODE> void run(const char* name, const char* table, uint32_t mask, int sms, float* out, double clock_ghz, int loads) {
    const int iters = 2000, grid = sms * 2, block = 352;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<MODE><<<grid, block>>>(table, mask, iters, out);
    cudaEventRecord(e0);
    gather<MODE><<<grid, block>>>(table, mask, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_iters = (double)grid * block * iters;
    const double clocks = ms * 1e-3 * clock_ghz * 1e9;
    printf("%-34s table %8u KB: %7.3f ms  %6.3f lane-records/clk/SM  %6.3f lane-loads/clk/SM\n", name, (mask + 1) / 8, ms, lane_iters / clocks / sms, lane_iters * loads / clocks / sms);
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount; const double ghz = pr.clockRate * 1e-6;
    char* table; float* out; cudaMalloc(&table, 64u << 20); cudaMemset(table, 0, 64u << 20); cudaMalloc(&out, 4);
    printf("%s, %d SMs, %.3f GHz (nominal)\n", pr.name, sms, ghz);
    for (uint32_t recs : {64u, 256u, 65536u, 262144u}) {           // 8 KB, 32 KB (L1), 8 MB, 32 MB (L2)
        const uint32_t mask = recs - 1;
        run<5>("1 x LDG.64", table, mask, sms, out, ghz, 1);
        run<0>("1 x LDG.128", table, mask, sms, out, ghz, 1);
        run<1>("2 x LDG.128 (one sector)", table, mask, sms, out, ghz, 2);
        run<2>("1 x LDG.256 (one sector)", table, mask, sms, out, ghz, 1);
        run<3>("4 x LDG.128 (64 B)", table, mask, sms, out, ghz, 4);
        run<4>("2 x LDG.256 (64 B)", table, mask, sms, out, ghz, 2);
        run<6>("3 x LDG.128 + LDG.64 (pair node)", table, mask, sms, out, ghz, 4);
    }
    return 0;
}
