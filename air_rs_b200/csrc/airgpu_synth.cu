// airgpu_synth.cu -- device twin of air_rs_b200/synth.py (workload generator).
//
// Not part of the decode path: it only fills HBM with the synthetic capture
// SURVEY.md 8(d) defines, so that multi-gigabyte inputs never cross PCIe.  Every
// sample is integer arithmetic on a counter hash; tests/test_synth.py checks the
// bytes against the numpy renderer.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/airgpu.h"

namespace {

thread_local char g_serr[256] = "";

#define SCU(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            snprintf(g_serr, sizeof g_serr, "%s failed: %s", #call, cudaGetErrorString(e_));   \
            return e_ == cudaErrorMemoryAllocation ? AIRGPU_ERR_NOMEM : AIRGPU_ERR_CUDA;       \
        }                                                                                      \
    } while (0)

constexpr unsigned long long kGolden = 0x9E3779B97F4A7C15ull;
constexpr int kSlots = 116;            // 4 preamble pulses + 112 data bits
constexpr int kFrameSpan = 242;        // 240 samples + 1 for the half-sample smear, +1 slack
constexpr size_t kChunk = (size_t)16 << 20;

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ int noise_of(unsigned long long seed_mul, unsigned long long ctr, int gain)
{
    unsigned long long h = mix64(seed_mul + ctr);
    int g = (int)(__dp4a((unsigned)h, 0x01010101u, 0u) + __dp4a((unsigned)(h >> 32), 0x01010101u, 0u)) - 1020;
    return (g * gain) >> 16;   // arithmetic shift == numpy's floor shift
}

// one thread per (frame, slot): add the pulse amplitude into the chunk accumulator
__global__ void scatter_kernel(const long long *start, const int *nbits, const unsigned char *payload,
                               const int *amp_i, const int *amp_q, const unsigned char *smear,
                               size_t f_lo, size_t f_count, long long shift, long long j0, long long n,
                               int2 *sig)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= f_count * kSlots) return;
    size_t f = f_lo + t / kSlots;
    int s = (int)(t % kSlots);
    int pos;
    if (s < 4) {
        pos = s == 0 ? 0 : s == 1 ? 2 : s == 2 ? 7 : 9;
    } else {
        int k = s - 4;
        if (k >= nbits[f]) return;
        int bit = (payload[f * 14 + (k >> 3)] >> (7 - (k & 7))) & 1;
        pos = 16 + 2 * k + (1 - bit);
    }
    long long p = start[f] + shift + pos - j0;
    int ai = amp_i[f], aq = amp_q[f];
    int li = 0, lq = 0;
    if (smear[f]) {
        li = ai >> 1;
        lq = aq >> 1;
    }
    if (p >= 0 && p < n) {
        atomicAdd(&sig[p].x, ai - li);
        atomicAdd(&sig[p].y, aq - lq);
    }
    if (smear[f] && p + 1 >= 0 && p + 1 < n) {
        atomicAdd(&sig[p + 1].x, li);
        atomicAdd(&sig[p + 1].y, lq);
    }
}

template <int FMT>
__global__ void compose_kernel(unsigned long long seed_mul, unsigned long long j0, unsigned long long n,
                               int gain, const int2 *sig, void *out)
{
    unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    unsigned long long j = j0 + k;
    int2 s = sig[k];
    int vi = noise_of(seed_mul, 2ull * j, gain) + s.x;
    int vq = noise_of(seed_mul, 2ull * j + 1ull, gain) + s.y;
    if (FMT == AIRGPU_FMT_U8) {
        vi = min(max(vi + 128, 0), 255);
        vq = min(max(vq + 128, 0), 255);
        reinterpret_cast<uchar2 *>(out)[k] = make_uchar2((unsigned char)vi, (unsigned char)vq);
    } else {
        vi = min(max(vi, -32768), 32767);
        vq = min(max(vq, -32768), 32767);
        reinterpret_cast<short2 *>(out)[k] = make_short2((short)vi, (short)vq);
    }
}

}  // namespace

struct airgpu_synth_table {
    int device = 0;
    size_t n = 0;
    std::vector<long long> h_start;   // sorted ascending
    long long *start = nullptr;
    int *nbits = nullptr;
    unsigned char *payload = nullptr;
    int *amp_i = nullptr, *amp_q = nullptr;
    unsigned char *smear = nullptr;
    int2 *sig = nullptr;              // chunk accumulator, allocated on first render
};

extern "C" {

const char *airgpu_synth_last_error(void) { return g_serr; }

void airgpu_synth_table_destroy(airgpu_synth_table *t)
{
    if (!t) return;
    cudaSetDevice(t->device);
    for (void *p : {(void *)t->start, (void *)t->nbits, (void *)t->payload, (void *)t->amp_i, (void *)t->amp_q,
                    (void *)t->smear, (void *)t->sig})
        if (p) cudaFree(p);
    delete t;
}

// Upload a frame table (air_rs_b200/synth.py: FrameTable), sorted by start sample.
int airgpu_synth_table_create(int device, const int64_t *start, const int32_t *nbits, const uint8_t *payload,
                              const int32_t *amp_i, const int32_t *amp_q, const uint8_t *smear, size_t n_frames,
                              airgpu_synth_table **out)
{
    if (!out) return AIRGPU_ERR_INVALID;
    *out = nullptr;
    SCU(cudaSetDevice(device));
    for (size_t k = 1; k < n_frames; ++k)
        if (start[k] < start[k - 1]) {
            snprintf(g_serr, sizeof g_serr, "frame table is not sorted by start");
            return AIRGPU_ERR_INVALID;
        }
    airgpu_synth_table *t = new (std::nothrow) airgpu_synth_table();
    if (!t) return AIRGPU_ERR_NOMEM;
    t->device = device;
    t->n = n_frames;
    t->h_start.assign(start, start + n_frames);
    const size_t m = std::max<size_t>(n_frames, 1);
#define UP(field, src, bytes)                                                              \
    do {                                                                                   \
        cudaError_t e_ = cudaMalloc(&t->field, (bytes) ? (bytes) : 16);                    \
        if (e_ == cudaSuccess && n_frames)                                                 \
            e_ = cudaMemcpy(t->field, src, bytes, cudaMemcpyHostToDevice);                 \
        if (e_ != cudaSuccess) {                                                           \
            snprintf(g_serr, sizeof g_serr, "table upload failed: %s", cudaGetErrorString(e_)); \
            airgpu_synth_table_destroy(t);                                                 \
            return AIRGPU_ERR_CUDA;                                                        \
        }                                                                                  \
    } while (0)
    (void)m;
    UP(start, start, n_frames * sizeof(long long));
    UP(nbits, nbits, n_frames * sizeof(int));
    UP(payload, payload, n_frames * 14);
    UP(amp_i, amp_i, n_frames * sizeof(int));
    UP(amp_q, amp_q, n_frames * sizeof(int));
    UP(smear, smear, n_frames);
#undef UP
    *out = t;
    return AIRGPU_OK;
}

// Render samples [j0, j0+n) into d_out (device): interleaved u8 or i16.  period > 0
// repeats the frame schedule every `period` samples; the noise never repeats.
int airgpu_synth_render(airgpu_synth_table *t, uint64_t seed, uint64_t j0, uint64_t n, uint32_t format,
                        int32_t noise_gain, uint64_t period, void *d_out, void *stream)
{
    if (!t || (n && !d_out)) return AIRGPU_ERR_INVALID;
    if (format != AIRGPU_FMT_U8 && format != AIRGPU_FMT_CS16) return AIRGPU_ERR_INVALID;
    if (noise_gain < 0 || noise_gain >= (1 << 21)) {
        snprintf(g_serr, sizeof g_serr, "noise gain out of range");
        return AIRGPU_ERR_INVALID;
    }
    if (period && period < (uint64_t)kFrameSpan) return AIRGPU_ERR_INVALID;
    SCU(cudaSetDevice(t->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!t->sig) SCU(cudaMalloc(&t->sig, kChunk * sizeof(int2)));
    const unsigned long long seed_mul = (unsigned long long)seed * kGolden;
    const size_t bps = format == AIRGPU_FMT_U8 ? 2 : 4;

    for (uint64_t c0 = 0; c0 < n; c0 += kChunk) {
        const long long cj0 = (long long)(j0 + c0);
        const long long cn = (long long)std::min<uint64_t>(kChunk, n - c0);
        SCU(cudaMemsetAsync(t->sig, 0, (size_t)cn * sizeof(int2), s));
        // repetitions of the schedule that can touch [cj0, cj0 + cn)
        long long rep_lo = 0, rep_hi = 0;
        if (period) {
            rep_lo = cj0 / (long long)period - 1;
            if (rep_lo < 0) rep_lo = 0;
            rep_hi = (cj0 + cn - 1) / (long long)period;
        }
        for (long long rep = rep_lo; rep <= rep_hi; ++rep) {
            const long long shift = rep * (long long)period;
            // frames with start + shift in (cj0 - kFrameSpan, cj0 + cn)
            auto lo = std::lower_bound(t->h_start.begin(), t->h_start.end(), cj0 - shift - kFrameSpan + 1);
            auto hi = std::lower_bound(t->h_start.begin(), t->h_start.end(), cj0 - shift + cn);
            const size_t f_lo = (size_t)(lo - t->h_start.begin());
            const size_t f_count = hi > lo ? (size_t)(hi - lo) : 0;
            if (!f_count) continue;
            const size_t threads = f_count * kSlots;
            scatter_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
                t->start, t->nbits, t->payload, t->amp_i, t->amp_q, t->smear, f_lo, f_count, shift, cj0, cn, t->sig);
            SCU(cudaGetLastError());
        }
        void *dst = static_cast<char *>(d_out) + c0 * bps;
        const unsigned blocks = (unsigned)((cn + 255) / 256);
        if (format == AIRGPU_FMT_U8)
            compose_kernel<AIRGPU_FMT_U8><<<blocks, 256, 0, s>>>(seed_mul, (unsigned long long)cj0, (unsigned long long)cn,
                                                                  noise_gain, t->sig, dst);
        else
            compose_kernel<AIRGPU_FMT_CS16><<<blocks, 256, 0, s>>>(seed_mul, (unsigned long long)cj0, (unsigned long long)cn,
                                                                    noise_gain, t->sig, dst);
        SCU(cudaGetLastError());
    }
    return AIRGPU_OK;
}

}  // extern "C"
