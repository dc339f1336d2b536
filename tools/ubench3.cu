// ubench3.cu -- issue rates of the conversion / special-function / compare instructions the CS16
// level (exact isqrt) and the survivor path lean on, and which pipe HSET2 / HFMA2 share.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3 tools/ubench3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
enum { I2F_U32, F2I_RZ, MUFU_SQRT, MUFU_RSQ, I2F_S16, FADD_RZ, FFMA_, HSET2_BF, HFMA2_BF, IMAD_, LOP3_, VMNX2, VMNX3, POPC_, FLO_, BREV_,
       ISETP_SEL, VOTE_, REDUX_, SHFL_, PRMT_, IADD3_, VMNX_PRED, LEA_, NOP_ };

template <int OP>
__device__ __forceinline__ void op(uint32_t &a, uint32_t b, uint32_t c)
{
    if (OP == I2F_U32) asm volatile("cvt.rz.f32.u32 %0, %0;" : "+r"(a));
    if (OP == F2I_RZ) asm volatile("cvt.rzi.u32.f32 %0, %0;" : "+r"(a));
    if (OP == MUFU_SQRT) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+r"(a));
    if (OP == MUFU_RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+r"(a));
    if (OP == I2F_S16) asm volatile("{ .reg .s16 h, g; mov.b32 {h, g}, %0; cvt.rn.f32.s16 %0, h; }" : "+r"(a));
    if (OP == FADD_RZ) asm volatile("add.rz.f32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == FFMA_) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == HSET2_BF) asm volatile("set.gt.u32.bf16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == HFMA2_BF) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == IMAD_) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == LOP3_) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == VMNX2) { a = __vminu2(a, b); asm volatile("" : "+r"(a)); }
    if (OP == VMNX3) { a = __vimax3_u16x2(a, b, c); asm volatile("" : "+r"(a)); }
    if (OP == POPC_) asm volatile("popc.b32 %0, %0;" : "+r"(a));
    if (OP == FLO_) asm volatile("bfind.u32 %0, %0;" : "+r"(a));
    if (OP == BREV_) asm volatile("brev.b32 %0, %0;" : "+r"(a));
    if (OP == ISETP_SEL) asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(a) : "r"(b), "r"(c));
    if (OP == VOTE_) asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "+r"(a) : "r"(b));
    if (OP == REDUX_) asm volatile("redux.sync.xor.b32 %0, %0, 0xffffffff;" : "+r"(a));
    if (OP == SHFL_) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(a) : "r"(b));
    if (OP == PRMT_) asm volatile("prmt.b32 %0, %0, %1, 0x5432;" : "+r"(a) : "r"(b));
    if (OP == IADD3_) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; sub.u32 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(c));
    if (OP == VMNX_PRED) { bool p, q; a = __vibmax_u16x2(a, b, &p, &q); a += p ? 1u : 0u; asm volatile("" : "+r"(a)); }
    if (OP == LEA_) asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(a) : "r"(b));
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
            op<B>(d[i], b, c);
        }
        b += 0x00010001u;
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ d[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}

template <int A, int B>
void run(const char *name, uint32_t *d, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int blocks = sms * 8;
    k<A, B><<<blocks, 256>>>(d, 3u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<A, B><<<blocks, 256>>>(d, 5u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double slots = (double)blocks * 256 * ITER * 8;     // lane-slots of each of A and B
    printf("%-26s %8.3f ms  -> %6.1f lane-slots/clk/SM (at %d MHz nominal)\n", name, ms,
           slots / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs; solo rows: one op per slot; pair rows: both ops per slot\n", p.name, p.multiProcessorCount);
    uint32_t *d;
    cudaMalloc(&d, 4096);
    int s = p.multiProcessorCount;
#define SOLO(X) run<X, NOP_>(#X, d, s)
#define PAIR(X, Y) run<X, Y>(#X " + " #Y, d, s)
    SOLO(I2F_U32); SOLO(F2I_RZ); SOLO(MUFU_SQRT); SOLO(MUFU_RSQ); SOLO(I2F_S16); SOLO(FADD_RZ); SOLO(FFMA_);
    SOLO(HSET2_BF); SOLO(HFMA2_BF); SOLO(IMAD_); SOLO(LOP3_); SOLO(VMNX2); SOLO(VMNX3); SOLO(POPC_); SOLO(FLO_); SOLO(BREV_);
    SOLO(ISETP_SEL); SOLO(VOTE_); SOLO(REDUX_); SOLO(SHFL_); SOLO(PRMT_); SOLO(IADD3_); SOLO(VMNX_PRED); SOLO(LEA_);
    PAIR(HSET2_BF, IMAD_); PAIR(HSET2_BF, LOP3_); PAIR(HSET2_BF, FFMA_); PAIR(HSET2_BF, VMNX2);
    PAIR(HFMA2_BF, IMAD_); PAIR(HFMA2_BF, LOP3_);
    PAIR(I2F_U32, MUFU_SQRT); PAIR(I2F_U32, F2I_RZ); PAIR(I2F_U32, LOP3_); PAIR(I2F_U32, IMAD_);
    PAIR(MUFU_SQRT, LOP3_); PAIR(MUFU_SQRT, IMAD_); PAIR(POPC_, LOP3_); PAIR(POPC_, MUFU_SQRT);
    PAIR(VOTE_, LOP3_); PAIR(ISETP_SEL, IMAD_); PAIR(VMNX2, VMNX3); PAIR(VMNX2, LOP3_);
    return 0;
}
