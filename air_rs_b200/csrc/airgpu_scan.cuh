// airgpu_scan.cuh -- the preamble gate (reference src/adsb/demod.rs:17-44) over one lane's
// 32 + 32 candidate offsets, on packed u16x2 words.
//
// Compiles for the device (the real instructions) and for the host (plain C++ emulation of the
// same packed operations) so that tools/emu_scan.cpp can check the index arithmetic and the
// hit-bit layout against a direct evaluation of the gate without a GPU.  The host build is test
// infrastructure; the product only ever runs the device build.
//
// Layout ("two streams"): a warp tile is 2048 candidate offsets = 2 streams of 1024.  Word w
// of the tile's level array holds (level[w], level[1024 + w]) in its (low, high) u16 halves, so
// one packed instruction works on offset x of stream 0 and offset 1024 + x of stream 1 at once,
// and -- unlike packing neighbouring offsets -- every operand of offset x + d is simply word
// x + d: no realignment (PRMT) and the sliding minima are shared between offsets.
//
// Levels are inverted (smaller level = larger magnitude).  With R[d] = word x + d:
//   highs 0,2,7,9 (demod.rs:20-25)          H = max(G[x], G[x+7]),  G[i] = max(R[i], R[i+2])
//   lows 1,3..6 | 8,10..13 | 14,15 (:26-31)  L = min3(C[x], C[x+7], P2[x+14]),
//                                           C[i] = min3(R[i+1], P2[i+3], P2[i+5]), P2[i] = min(R[i], R[i+1])
//   the offset fails the preamble test iff H > L (a high below a low; ties pass).
// 5.8 packed instructions per offset PAIR (185 for 32 pairs).
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define AIRGPU_HD __host__ __device__ __forceinline__
#else
#define AIRGPU_HD inline
#endif

namespace airgpu {

constexpr int kStream = 1024;            // candidate offsets per stream (two streams per warp tile)
constexpr int kLaneX = kStream / 32;     // 32 offsets of each stream per lane
constexpr int kLaneWords = kLaneX + 15;  // words a lane reads: x .. x + 31 + 15

namespace packed {
#if defined(__CUDA_ARCH__)
AIRGPU_HD uint32_t min2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
AIRGPU_HD uint32_t max2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
#if defined(AIRGPU_GATE_2IN)
AIRGPU_HD uint32_t min3(uint32_t a, uint32_t b, uint32_t c)   // A/B: tools/ubench4.cu (the empty asm keeps ptxas from re-fusing)
{
    uint32_t t = __vminu2(a, b);
    asm volatile("" : "+r"(t));
    return __vminu2(t, c);
}
#else
AIRGPU_HD uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
#endif
AIRGPU_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
// 1 in every half where hi > lo, any 16-bit levels: max - lo is non-zero exactly there (no borrow
// between the halves because max >= lo in each), clamped to 1.  The subtraction is written as
// lo * minus_one + max so that it issues on the FMA pipe (minus_one = 0xFFFFFFFF is a kernel
// parameter: ptxas cannot turn it back into an integer-pipe IADD3).
AIRGPU_HD uint32_t fail_flags(uint32_t lo, uint32_t hi, uint32_t minus_one)
{
    return __vminu2(lo * minus_one + __vmaxu2(hi, lo), 0x00010001u);
}
// 0xFFFF in every half where hi > lo (bf16 compare of valid, ordered level patterns)
AIRGPU_HD uint32_t fail_mask_bf16(uint32_t lo, uint32_t hi)
{
    uint32_t d;
    asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}
// c + sum over the four bytes of (signed byte of a) * (unsigned byte of b)
AIRGPU_HD uint32_t dp4a_su(uint32_t a, uint32_t b, uint32_t c)
{
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"((int)c));
    return (uint32_t)d;
}
AIRGPU_HD uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
#else
inline uint32_t lo16(uint32_t a) { return a & 0xFFFFu; }
inline uint32_t hi16(uint32_t a) { return a >> 16; }
inline uint32_t pack(uint32_t l, uint32_t h) { return l | (h << 16); }
inline uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
inline uint32_t min2(uint32_t a, uint32_t b) { return pack(umin(lo16(a), lo16(b)), umin(hi16(a), hi16(b))); }
inline uint32_t max2(uint32_t a, uint32_t b) { return pack(umax(lo16(a), lo16(b)), umax(hi16(a), hi16(b))); }
inline uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return min2(min2(a, b), c); }
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)((v >> (8 * ((s >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return r;
}
inline uint32_t fail_flags(uint32_t lo, uint32_t hi, uint32_t minus_one)
{
    return min2(lo * minus_one + max2(hi, lo), 0x00010001u);
}
inline uint32_t fail_mask_bf16(uint32_t lo, uint32_t hi)
{
    return (hi16(hi) > hi16(lo) ? 0xFFFF0000u : 0u) | (lo16(hi) > lo16(lo) ? 0xFFFFu : 0u);
}
inline uint32_t dp4a_su(uint32_t a, uint32_t b, uint32_t c)
{
    int d = (int)c;
    for (int k = 0; k < 4; ++k) d += (int)(int8_t)(a >> (8 * k)) * (int)((b >> (8 * k)) & 0xFF);
    return (uint32_t)d;
}
inline uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c)
{
    for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 0xFF) * ((b >> (8 * k)) & 0xFF);
    return c;
}
#endif
}  // namespace packed

// R[d] = word x0 + d of the tile (d = 0 .. kLaneWords-1; R has 48 entries, the last is unused).
// hits[h] bit b (h = 0, 1): the preamble test PASSED for stream hit_stream(b), offset x0 + hit_x(h, b).
//
// The hit bits are gathered on the FMA pipe: one dot product per offset pair adds that pair's two
// bit weights (1 << e for stream 0, 16 << e for stream 1) into a byte-wide accumulator, so byte g of
// hits[h] holds x = 16 h + 4 g .. + 3 of both streams.
//   kBf16 (U8 levels, <= 0x7F00: valid, ordered bf16 patterns): HSET2.BF16 leaves 0xFFFF (two bytes
//     of -1) per failing half; signed x unsigned dot products subtract the weights from -1.
//   otherwise (CS16 levels use all 16 bits): fail_flags() leaves 1 per failing half; unsigned dot
//     products add the weights up from 0 and the result is inverted once.
template <bool kBf16>
AIRGPU_HD void gate_scan(const uint32_t (&R)[48], uint32_t (&hits)[2], uint32_t minus_one)
{
    using namespace packed;
    uint32_t P2[46], G[39], C[39];
#pragma unroll
    for (int i = 3; i < 46; ++i) P2[i] = min2(R[i], R[i + 1]);
#pragma unroll
    for (int i = 0; i < 39; ++i) G[i] = max2(R[i], R[i + 2]);
#pragma unroll
    for (int i = 0; i < 39; ++i) C[i] = min3(R[i + 1], P2[i + 3], P2[i + 5]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t acc[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            acc[g] = (kBf16 && g == 0) ? 0xFFFFFFFFu : 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int x = 16 * h + 4 * g + e;
                const uint32_t hi = max2(G[x], G[x + 7]);
                const uint32_t lo = min3(C[x], C[x + 7], P2[x + 14]);
                const uint32_t wgt = (1u << e) | (0x100000u << e);
                if (kBf16) acc[g] = dp4a_su(fail_mask_bf16(lo, hi), wgt, acc[g]);
                else acc[g] = dp4a_uu(fail_flags(lo, hi, minus_one), wgt, acc[g]);
            }
        }
        const uint32_t v = ((acc[3] * 256u + acc[2]) * 256u + acc[1]) * 256u + acc[0];
        hits[h] = kBf16 ? v : ~v;
    }
}

// bit b of hits[h]  ->  which stream, which of the lane's 32 offsets
AIRGPU_HD int hit_stream(int b) { return (b >> 2) & 1; }
AIRGPU_HD int hit_x(int h, int b) { return 16 * h + 4 * (b >> 3) + (b & 3); }

// ---- shared-memory layout of the word array ----------------------------------------------------
// 16-byte chunks of 4 words; one pad chunk after every 8 keeps both access patterns conflict
// free: phase 1 stores chunks 2c and 2c+1 from lane c, phase 2 loads chunks 8*lane + k.
constexpr int kTileWords = kStream + 240;                       // 1264: words 0 .. 1023 + 239 (+1)
constexpr int kTileChunks8 = kTileWords / 8;                    // 158 phase-1 units of 8 words
constexpr int kTileWordsPadded = kTileWords + 4 * ((kTileWords + 31) / 32);
AIRGPU_HD int phys_chunk4(int c) { return c + (c >> 3); }
AIRGPU_HD int phys_word(int w) { return w + ((w >> 5) << 2); }

// ---- where the scalar readers find a level (indices into the tile's array viewed as u16) --------
// Candidate = (stream st, word xw): offset i = st * kStream + xw reads the (st ? high : low) u16
// halves of consecutive words starting at word xw.
AIRGPU_HD int level_index(int xw, int st, int k) { return 2 * phys_word(xw + k) + st; }

// DF test (demod.rs:45-54): levels 16..25 of the candidate.  The ten words are consecutive except
// that one pad (4 words = 8 u16) may fall inside the run; `cross` is the first k behind it
// (>= 10: none).  Level k is at level_index(xw, st, 16) + 2k, + 8 when k >= cross.
AIRGPU_HD int df_cross(int xw) { return 32 - ((xw + 16) & 31); }

// Slicer (demod.rs:92-131, 180-201): lane `lane` handles frame bits 31 - lane + 32 r (so that a
// ballot is the big-endian frame word as it stands), i.e. levels 16 + 2k and 17 + 2k.  Rounds are
// 64 words = 72 padded words = 144 u16 apart; the second level is the next word, one pad further
// when the first is the last word before a pad.  Round 3 only has bits 96..111: lanes 16..31.
AIRGPU_HD int slicer_word(int xw, int lane) { return xw + 16 + 2 * (31 - lane); }
AIRGPU_HD int slicer_step(int wj) { return (wj & 31) == 31 ? 10 : 2; }
constexpr int kSlicerRoundStride = 144;

}  // namespace airgpu
