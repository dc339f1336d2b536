// ubench2.cu -- which pipe does what: pairs of instruction kinds issued together.
// If two kinds share a pipe the pair takes the SUM of their solo times; if they sit on
// different pipes it takes about the MAX.  asm volatile keeps the compiler from folding.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define ITER 2048
enum { VMNX3, VMNX2, PRMT_, LOP3_, SHF_, IADD3_, IMAD_, IMADHI, IDP_, HMNMX_BF, HMNMX_H, FFMA_, IMADW, HSUBBF, NOP_ };

template <int OP>
__device__ __forceinline__ void op(uint32_t &a, uint32_t b, uint32_t c)
{
    if (OP == VMNX3) { a = __vimax3_u16x2(a, b, c); asm volatile("" : "+r"(a)); }
    if (OP == VMNX2) { a = __vminu2(a, b); asm volatile("" : "+r"(a)); }
    if (OP == PRMT_) asm volatile("prmt.b32 %0, %0, %1, 0x5432;" : "+r"(a) : "r"(b));
    if (OP == LOP3_) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == SHF_) asm volatile("shf.r.wrap.b32 %0, %0, %1, 16;" : "+r"(a) : "r"(b));
    if (OP == IADD3_) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; sub.u32 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(c));
    if (OP == IMAD_) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == IDP_) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == HMNMX_BF) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == HMNMX_H) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == IMADW) { unsigned long long t; asm volatile("mul.wide.u32 %0, %1, 65536;" : "=l"(t) : "r"(a)); a = (uint32_t)(t >> 32) + (uint32_t)t + b; asm volatile("" : "+r"(a)); }
    if (OP == HSUBBF) asm volatile("sub.rn.bf16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == FFMA_) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
}

template <int A, int B>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed)
{
    uint32_t a[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + 1) + i; d[i] = a[i] ^ 0x5555u; }
    uint32_t b = seed ^ 0x00ff00ffu, c = seed + 77u;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            op<A>(a[i], b, c);
            if (B != NOP_) op<B>(d[i], c, b);
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ d[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}

template <int A, int B>
float run(uint32_t *d, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<A, B><<<sms * 8, 256>>>(d, 3u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<A, B><<<sms * 8, 256>>>(d, 5u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
#define SOLO(X) { float t = run<X, NOP_>(d, sms); printf("%-10s solo %.3f ms  -> %.1f lane-ops/clk/SM\n", #X, t, lane / (t*1e-3) / sms / clk); }
#define PAIR(X, Y) { float t = run<X, Y>(d, sms); printf("%-10s + %-10s %.3f ms\n", #X, #Y, t); }

// exhaustive check: bf16x2 min/max on 15-bit patterns == integer min/max (denormals included)
__global__ void bfcheck(unsigned long long *bad)
{
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;   // 0..32767
    if (x > 32512u) return;
    unsigned long long nb = 0;
    for (uint32_t y = 0; y <= 32512u; ++y) {
        uint32_t p = x | (y << 16), q = y | (x << 16), mx, mn;
        asm volatile("max.bf16x2 %0, %1, %2;" : "=r"(mx) : "r"(p), "r"(q));
        asm volatile("min.bf16x2 %0, %1, %2;" : "=r"(mn) : "r"(p), "r"(q));
        uint32_t hi = x > y ? x : y, lo = x < y ? x : y;
        if (mx != (hi | (hi << 16)) || mn != (lo | (lo << 16))) nb++;
    }
    if (nb) atomicAdd(bad, nb);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double clk = clk_khz * 1e3; double lane = (double)sms * 8 * 256 * ITER * 8;
    uint32_t *d; cudaMalloc(&d, 4096);
    printf("%s, %d SMs, solo reference: 128 lane-ops/clk/SM = full rate\n", p.name, sms);
    SOLO(VMNX3) SOLO(VMNX2) SOLO(PRMT_) SOLO(LOP3_) SOLO(SHF_) SOLO(IADD3_) SOLO(IMAD_) SOLO(IMADHI) SOLO(IDP_) SOLO(HMNMX_BF) SOLO(HMNMX_H) SOLO(FFMA_)
    SOLO(IMADW) SOLO(HSUBBF) PAIR(VMNX3, IMADW) PAIR(VMNX3, HSUBBF) PAIR(HSUBBF, IMAD_) PAIR(HSUBBF, LOP3_) PAIR(IMADW, IMAD_)
    PAIR(VMNX3, LOP3_) PAIR(VMNX3, IMAD_) PAIR(VMNX3, VMNX2) PAIR(VMNX3, HMNMX_BF) PAIR(VMNX3, IADD3_) PAIR(VMNX3, FFMA_)
    PAIR(VMNX2, LOP3_) PAIR(VMNX2, IMAD_) PAIR(VMNX2, HMNMX_BF) PAIR(VMNX2, FFMA_) PAIR(VMNX2, VMNX2)
    PAIR(HMNMX_BF, LOP3_) PAIR(HMNMX_BF, IMAD_) PAIR(HMNMX_BF, FFMA_) PAIR(HMNMX_BF, HMNMX_BF)
    PAIR(PRMT_, IMAD_) PAIR(PRMT_, IDP_) PAIR(PRMT_, LOP3_) PAIR(LOP3_, IADD3_) PAIR(IADD3_, IMAD_) PAIR(IADD3_, IADD3_)
    PAIR(IMAD_, IDP_) PAIR(IMADHI, LOP3_) PAIR(FFMA_, IMAD_) PAIR(FFMA_, LOP3_)
    unsigned long long *bad; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
    bfcheck<<<128, 256>>>(bad);
    unsigned long long hb; cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
    printf("bf16x2 min/max vs integer on all 15-bit level pairs: %llu mismatches\n", hb);
    return 0;
}
