// ubench.cu -- per-SM issue rates of the integer instructions the decode kernel leans on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
// Each kernel runs ITER x UNROLL dependent-free chains per thread, 148*k CTAs x 256 threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed)
{
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
    uint32_t b = seed ^ 0x00ff00ffu, c = seed + 77u;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = __vimax3_u16x2(a[i], b, c);                 // VIMNMX3.U16x2
            if (OP == 1) a[i] = __vminu2(a[i], b);                          // VIMNMX.U16x2
            if (OP == 2) a[i] = __byte_perm(a[i], b, 0x5432);               // PRMT
            if (OP == 3) a[i] = __dp4a(a[i], b, c);                         // IDP.4A
            if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));   // LOP3
            if (OP == 5) a[i] = a[i] + 0x80008000u - b;                     // IADD3
            if (OP == 6) a[i] = a[i] * b + c;                               // IMAD
            if (OP == 7) a[i] = max(a[i], b);                               // IMNMX / VIMNMX.U32
            if (OP == 8) { a[i] = __vimax3_u16x2(a[i], b, c); a[(i+1)&7] = a[(i+1)&7] * b + c; }  // mix ALU+FMA
            if (OP == 9) { a[i] = __vimax3_u16x2(a[i], b, c); a[(i+1)&7] = __dp4a(a[(i+1)&7], b, c); }
            if (OP == 10) a[i] = __vimax3_u32(a[i], b, c);
            if (OP == 11) a[i] = __funnelshift_r(a[i], b, 16);              // SHF
            if (OP == 12) { float f = __uint_as_float(a[i]); f = fmaf(f, 1.0001f, 0.5f); a[i] = __float_as_uint(f); } // FFMA
            if (OP == 13) { a[i] = __vimax3_u16x2(a[i], b, c); float f = __uint_as_float(a[(i+1)&7]); f = fmaf(f, 1.0001f, 0.5f); a[(i+1)&7] = __float_as_uint(f); }
        }
        b += 0x00010001u;
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}

template <int OP>
void run(const char *name, int ops_per_slot, uint32_t *d, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(d, 3u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(d, 5u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = (double)blocks * 256 * ITER * 8 * ops_per_slot;
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double per_clk_sm_at_max = lane_ops / (ms * 1e-3) / sms / (clk_khz * 1e3);
    printf("%-28s %8.3f ms  %7.2f Tlane-op/s  %6.1f lane-ops/clk/SM (at %d MHz nominal)\n", name, ms,
           lane_ops / (ms * 1e-3) / 1e12, per_clk_sm_at_max, clk_khz / 1000);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    uint32_t *d;
    cudaMalloc(&d, 4096);
    int sms = p.multiProcessorCount;
    run<0>("VIMNMX3.U16x2", 1, d, sms);
    run<1>("VIMNMX.U16x2", 1, d, sms);
    run<2>("PRMT", 1, d, sms);
    run<3>("IDP.4A", 1, d, sms);
    run<4>("LOP3", 1, d, sms);
    run<5>("IADD3", 1, d, sms);
    run<6>("IMAD", 1, d, sms);
    run<7>("IMNMX.U32", 1, d, sms);
    run<10>("VIMNMX3.U32", 1, d, sms);
    run<11>("SHF", 1, d, sms);
    run<12>("FFMA", 1, d, sms);
    run<8>("VIMNMX3.U16x2 + IMAD", 2, d, sms);
    run<9>("VIMNMX3.U16x2 + IDP.4A", 2, d, sms);
    run<13>("VIMNMX3.U16x2 + FFMA", 2, d, sms);
    return 0;
}
