// ubench4.cu -- the packed preamble gate (airgpu_scan.cuh) on registers only: cycles per call for the
// shipped min3-based form and for a 2-input-only form (VIMNMX3 mixed with 2-input VIMNMX runs the
// 2-input ones at half rate: tools/ubench2.cu "VMNX2 + VMNX3").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench4 tools/ubench4.cu [-DAIRGPU_GATE_2IN=1]
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../air_rs_b200/csrc/airgpu_scan.cuh"
#define ITER 512
template <bool BF>
__global__ void __launch_bounds__(128, 8) k(uint32_t *out, uint32_t seed, uint32_t minus_one)
{
    uint32_t R[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) R[i] = (seed * (threadIdx.x + 1) + i * 0x00070003u) & 0x3FFF3FFFu;
    uint32_t acc = 0;
    for (int it = 0; it < ITER; ++it) {
        uint32_t h[2];
        airgpu::gate_scan<BF>(R, h, minus_one);
        acc += h[0] ^ h[1];
#pragma unroll
        for (int i = 0; i < 48; i += 6) R[i] ^= (acc & 0x00010001u);   // keep the gate inside the loop
    }
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}
template <bool BF>
void run(const char *name, uint32_t *d, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int blocks = sms * 8 * 4;
    k<BF><<<blocks, 128>>>(d, 3u, 0xFFFFFFFFu);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<BF><<<blocks, 128>>>(d, 5u, 0xFFFFFFFFu);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double calls_per_smsp = (double)blocks * 4 * ITER / (sms * 4);
    printf("%-10s %8.3f ms -> %7.1f cycles per warp-level gate call per SMSP (at %d MHz nominal)\n", name, ms,
           ms * 1e-3 * clk_khz * 1e3 / calls_per_smsp, clk_khz / 1000);
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    uint32_t *d;
    cudaMalloc(&d, 4096);
#ifdef AIRGPU_GATE_2IN
    printf("%s: gate with 2-input min/max only\n", p.name);
#else
    printf("%s: gate as shipped (min3)\n", p.name);
#endif
    run<true>("bf16/U8", d, p.multiProcessorCount);
    run<false>("u16/CS16", d, p.multiProcessorCount);
    return 0;
}
