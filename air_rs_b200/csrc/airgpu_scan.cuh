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

#ifndef AIRGPU_HITS_IDP
#define AIRGPU_HITS_IDP 1     // 1: gather the hit bits with IDP.4A on the FMA pipe (bf16-comparable levels only; measured 4.37 vs 4.46 ms), 0: PRMT + LOP3 merge
#endif

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
AIRGPU_HD uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
AIRGPU_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
// bits 15 and 31: set iff hi > lo in that half.  The other bits are unspecified.
template <bool kBf16>
AIRGPU_HD uint32_t fail_bits(uint32_t lo, uint32_t hi)
{
    if (kBf16) {
        // levels <= 0x7F00 read as bf16 are finite, non-negative and ordered like the integers
        // (subnormals included), so sign(lo - hi) is the comparison and a tie gives +0: one
        // HFMA2.BF16 on the FMA pipe instead of an integer-pipe instruction
        uint32_t d;
        asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(0xBF80BF80u), "r"(lo));
        return d;
    } else {
        bool ph, pl;
        (void)__vibmax_u16x2(lo, hi, &ph, &pl);          // predicates: lo >= hi
        return (ph ? 0u : 0x80000000u) | (pl ? 0u : 0x00008000u);
    }
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
template <bool kBf16>
inline uint32_t fail_bits(uint32_t lo, uint32_t hi)
{
    // the unspecified bits are filled with ones so that a consumer relying on them shows up
    return 0x7FFF7FFFu | (hi16(hi) > hi16(lo) ? 0x80000000u : 0u) | (lo16(hi) > lo16(lo) ? 0x8000u : 0u);
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
#endif
}  // namespace packed

// R[d] = word x0 + d of the tile (d = 0 .. kLaneWords-1; R has 48 entries, the last is unused).
// hits[h] bit b (h = 0, 1): the preamble test PASSED for stream hit_stream(b), offset x0 + hit_x(h, b).
template <bool kBf16>
AIRGPU_HD void gate_scan(const uint32_t (&R)[48], uint32_t (&hits)[2])
{
    using namespace packed;
    uint32_t P2[46], G[39], C[39];
#pragma unroll
    for (int i = 3; i < 46; ++i) P2[i] = min2(R[i], R[i + 1]);
#pragma unroll
    for (int i = 0; i < 39; ++i) G[i] = max2(R[i], R[i + 2]);
#pragma unroll
    for (int i = 0; i < 39; ++i) C[i] = min3(R[i + 1], P2[i + 3], P2[i + 5]);
#if AIRGPU_HITS_IDP
    if (kBf16) {
        // Hit bits gathered on the FMA pipe: the compare leaves 0xFFFF (= two bytes of -1) per
        // failing half, and one signed x unsigned dot product per offset pair subtracts that
        // pair's two weights (1 << e for stream 0, 16 << e for stream 1) from a byte-wide
        // accumulator.  Starting from -1, byte g of the result is ~(fail flags) of x = 4g .. 4g+3.
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t acc[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                acc[g] = g == 0 ? 0xFFFFFFFFu : 0u;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int x = 16 * h + 4 * g + e;
                    const uint32_t hi = max2(G[x], G[x + 7]);
                    const uint32_t lo = min3(C[x], C[x + 7], P2[x + 14]);
                    acc[g] = dp4a_su(fail_mask_bf16(lo, hi), (1u << e) | (0x100000u << e), acc[g]);
                }
            }
            hits[h] = ((acc[3] * 256u + acc[2]) * 256u + acc[1]) * 256u + acc[0];
        }
    } else
#endif
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t fails[2] = {0u, 0u};    // [g], bit 8*j + 7 - r: see below
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            uint32_t d[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int x = 2 * (8 * h + q) + e;
                const uint32_t hi = max2(G[x], G[x + 7]);
                const uint32_t lo = min3(C[x], C[x + 7], P2[x + 14]);
                d[e] = fail_bits<kBf16>(lo, hi);
            }
            // top bits of the four bytes: (stream 0, x even), (stream 1, x even), (stream 0, x odd), (stream 1, x odd)
            const uint32_t F = prmt(d[0], d[1], 0x7531);
            // keep bits 7..8-r of every byte, take bit 7-r from F (select: one LOP3)
            const int r = q & 3;
            const uint32_t keep = 0x01010101u * (0xFFu & ~(0xFFu >> r));
            fails[q >> 2] = r == 0 ? F : ((fails[q >> 2] & keep) | ((F >> r) & ~keep));
        }
        hits[h] = (~fails[0] & 0xF0F0F0F0u) | ((~fails[1] & 0xF0F0F0F0u) >> 4);
    }
}

// bit b of hits[h]  ->  which stream, which of the lane's 32 offsets
template <bool kBf16>
AIRGPU_HD int hit_stream(int b)
{
#if AIRGPU_HITS_IDP
    if (kBf16) return (b >> 2) & 1;
#endif
    return (b >> 3) & 1;
}
template <bool kBf16>
AIRGPU_HD int hit_x(int h, int b)
{
#if AIRGPU_HITS_IDP
    if (kBf16) return 16 * h + 4 * (b >> 3) + (b & 3);
#endif
    return 2 * (8 * h + 7 - (b & 7)) + (b >> 4);
}

// ---- shared-memory layout of the word array ----------------------------------------------------
// 16-byte chunks of 4 words; one pad chunk after every 8 keeps both access patterns conflict
// free: phase 1 stores chunks 2c and 2c+1 from lane c, phase 2 loads chunks 8*lane + k.
constexpr int kTileWords = kStream + 240;                       // 1264: words 0 .. 1023 + 239 (+1)
constexpr int kTileChunks8 = kTileWords / 8;                    // 158 phase-1 units of 8 words
constexpr int kTileWordsPadded = kTileWords + 4 * ((kTileWords + 31) / 32);
AIRGPU_HD int phys_chunk4(int c) { return c + (c >> 3); }
AIRGPU_HD int phys_word(int w) { return w + ((w >> 5) << 2); }

// ---- where the scalar readers find a level (indices into the tile's array viewed as u16) --------
// Level `k` levels after candidate i's first sample: candidates of stream s = i >> 10 read the
// (s ? high : low) u16 halves of consecutive words starting at word (i & 1023).
AIRGPU_HD int level_index(int i, int k) { return 2 * phys_word((i & (kStream - 1)) + k) + (i >> 10); }

// DF test (demod.rs:45-54): levels 16..25 of candidate i.  The ten words are consecutive except
// that one pad (4 words = 8 u16) may fall inside the run; `cross` is the first k behind it
// (>= 10: none).  Level k is at df_base + 2k, + 8 when k >= cross.
AIRGPU_HD int df_cross(int i) { return 32 - (((i & (kStream - 1)) + 16) & 31); }

// Slicer (demod.rs:92-131, 180-201): lane `lane` handles frame bits lane + 32 r, i.e. levels
// 16 + 2k and 17 + 2k.  Rounds are 64 words = 72 padded words = 144 u16 apart; the second level
// is the next word, one pad further when the first is the last word before a pad.
AIRGPU_HD int slicer_word(int i, int lane) { return (i & (kStream - 1)) + 16 + 2 * lane; }
AIRGPU_HD int slicer_step(int wj) { return (wj & 31) == 31 ? 10 : 2; }
constexpr int kSlicerRoundStride = 144;

}  // namespace airgpu
