// ubench5.cu -- does the order of 2-input and 3-input packed min/max matter?  Same instruction counts per
// iteration (16 x VIMNMX.U16x2 + 8 x VIMNMX3.U16x2 on independent registers), different interleavings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench5 tools/ubench5.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
// min in even iterations, max in odd ones: consecutive operations on one register cannot be re-fused by ptxas
template <int ODD> __device__ __forceinline__ void v2(uint32_t &a, uint32_t b)
{
    if (ODD) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    else asm volatile("min.u16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
}
template <int ODD> __device__ __forceinline__ void v3(uint32_t &a, uint32_t b, uint32_t c)
{
    if (ODD) asm volatile("{ .reg .b32 t; max.u16x2 t, %1, %2; max.u16x2 %0, %0, t; }" : "+r"(a) : "r"(b), "r"(c));
    else asm volatile("{ .reg .b32 t; min.u16x2 t, %1, %2; min.u16x2 %0, %0, t; }" : "+r"(a) : "r"(b), "r"(c));
}
template <int PAT, int ODD>
__device__ __forceinline__ void body(uint32_t (&a)[16], uint32_t (&d)[8], uint32_t b, uint32_t c);

template <int PAT>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed)
{
    uint32_t a[16], d[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed * (threadIdx.x + 1) + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = a[i] ^ 0x5555u;
    uint32_t b = seed ^ 0x00ff00ffu, c = seed + 77u;
    for (int it = 0; it < ITER; it += 2) {
        body<PAT, 0>(a, d, b, c);
        body<PAT, 1>(a, d, b, c);
        b += 0x00010001u;
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r ^= a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= d[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}
template <int PAT, int ODD>
__device__ __forceinline__ void body(uint32_t (&a)[16], uint32_t (&d)[8], uint32_t b, uint32_t c)
{
        if (PAT == 0) {          // runs: 16 x V2, then 8 x V3
#pragma unroll
            for (int i = 0; i < 16; ++i) v2<ODD>(a[i], b);
#pragma unroll
            for (int i = 0; i < 8; ++i) v3<ODD>(d[i], b, c);
        } else if (PAT == 1) {   // V2 V2 V3 repeated
#pragma unroll
            for (int i = 0; i < 8; ++i) { v2<ODD>(a[2 * i], b); v2<ODD>(a[2 * i + 1], b); v3<ODD>(d[i], b, c); }
        } else if (PAT == 2) {   // V2 V3 V2 repeated (no two V2 adjacent within a group boundary... V2 V3 V2 | V2 V3 V2)
#pragma unroll
            for (int i = 0; i < 8; ++i) { v2<ODD>(a[2 * i], b); v3<ODD>(d[i], b, c); v2<ODD>(a[2 * i + 1], b); }
        } else if (PAT == 3) {   // runs of 4 V2 + 2 V3
#pragma unroll
            for (int i = 0; i < 4; ++i) { v2<ODD>(a[4 * i], b); v2<ODD>(a[4 * i + 1], b); v2<ODD>(a[4 * i + 2], b); v2<ODD>(a[4 * i + 3], b); v3<ODD>(d[2 * i], b, c); v3<ODD>(d[2 * i + 1], b, c); }
        } else if (PAT == 4) {   // only the 16 V2
#pragma unroll
            for (int i = 0; i < 16; ++i) v2<ODD>(a[i], b);
        } else if (PAT == 5) {   // only the 8 V3
#pragma unroll
            for (int i = 0; i < 8; ++i) v3<ODD>(d[i], b, c);
        } else if (PAT == 6) {   // 32 V2 instead (what an all-2-input gate would issue: 16 + 2 x 8)
#pragma unroll
            for (int i = 0; i < 16; ++i) v2<ODD>(a[i], b);
#pragma unroll
            for (int i = 0; i < 16; ++i) v2<ODD>(a[i], c);
        }
}

template <int PAT>
void run(const char *name, uint32_t *d, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int blocks = sms * 8;
    k<PAT><<<blocks, 256>>>(d, 3u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<PAT><<<blocks, 256>>>(d, 5u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double iters_per_smsp = (double)blocks * 8 * ITER / (sms * 4);   // warp-iterations per scheduler
    printf("%-34s %8.3f ms -> %6.2f cycles per warp-iteration per SMSP (at %d MHz nominal)\n", name, ms,
           ms * 1e-3 * clk_khz * 1e3 / iters_per_smsp, clk_khz / 1000);
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    uint32_t *d;
    cudaMalloc(&d, 4096);
    int s = p.multiProcessorCount;
    printf("%s: 16 x VIMNMX.U16x2 + 8 x VIMNMX3.U16x2 per iteration, by interleaving\n", p.name);
    run<4>("16 V2 only", d, s);
    run<5>("8 V3 only", d, s);
    run<0>("16 V2 then 8 V3", d, s);
    run<1>("(V2 V2 V3) x 8", d, s);
    run<2>("(V2 V3 V2) x 8", d, s);
    run<3>("(V2 V2 V2 V2 V3 V3) x 4", d, s);
    run<6>("32 V2 (all-2-input equivalent)", d, s);
    return 0;
}
