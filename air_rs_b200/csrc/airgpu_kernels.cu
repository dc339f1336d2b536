// airgpu_kernels.cu -- fused ADS-B decode kernel for sm_100a.
//
// Reference path reproduced bit for bit (file:line in jaxsonpd/air_rs):
//   magnitude      src/utils.rs:46-52
//   offset loop    src/adsb.rs:98-114  (every i in [0, len-240), ascending, no skip)
//   gate           src/adsb/demod.rs:17-57
//   bit slicer     src/adsb/demod.rs:92-131 + 180-201  (bit k = m[2k] > m[2k+1])
//   CRC-24         src/adsb/crc.rs:10-40
//   1-bit repair   src/adsb/crc.rs:49-65
//
// Design:
//   * a WARP is the unit of work: one tile of kWarpTile candidate offsets at a time, in a
//     private slice of shared memory (no __syncthreads anywhere); IQ is read from HBM once
//     (the 240-sample halo of each stream comes from L2), with 16-byte coalesced loads;
//     nothing but frame records is written back;
//   * the per-sample "level" is kept in shared memory as u16 and is INVERTED
//     (smaller level = larger magnitude):
//       U8  : level = I*(255-I) + Q*(255-Q).  With re = (2I-255)*128 the
//             reference magnitude is isqrt(16384*(130050 - 4*level)); distinct
//             levels differ by >= 2, so 128*sqrt(k) moves by > 1.4 and the floor
//             never merges two of them: comparing levels is EXACTLY comparing
//             reference magnitudes, ties included.  Two IDP.4A per sample pair,
//             no square root, no LUT.
//       CS16: level = 65535 - isqrt(re^2 + im^2) (MUFU sqrt + integer fix-up).
//   * the preamble test runs on packed u16x2 words (two offsets per instruction,
//     airgpu_scan.cuh): a warp tile is two streams of 1024 offsets and word w holds
//     (level[w], level[1024 + w]); a lane owns 32 consecutive offsets of each stream
//     so the window lives in registers; warps leave the fast path only when some
//     lane saw a preamble (5.6e-4 per offset on noise);
//   * survivors are spread over the lanes for the DF test, then sliced with warp
//     ballots, checked with a 112-entry single-bit syndrome table and written to
//     the tile's fixed scratch slots in offset order (no atomic with a return value).
#include <algorithm>
#include <cstdlib>

#include "airgpu_kernels.cuh"
#include "airgpu_scan.cuh"

namespace airgpu {
namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;

#ifndef AIRGPU_STAGE_DEFAULT
#define AIRGPU_STAGE_DEFAULT 0      // 1: TMA-staged U8 kernel on aligned single-segment captures (A/B: DESIGN.md section 5)
#endif

// One store to an NVSwitch multicast address: the switch writes it to the same offset of every rank's mapping.
__device__ __forceinline__ void multimem_st_u64(unsigned long long *mc, unsigned long long v)
{
    asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(mc), "l"(v) : "memory");
}

// ---- CRC-24 single-bit syndromes -------------------------------------------
// Frame bit p (MSB first, p = 0..111) has weight x^(111-p) in (data * x^24 + parity),
// so its syndrome is x^(111-p) mod 0x1FFF409.  For p >= 88 that is the parity bit
// itself.  XOR over the set bits == crc(data) ^ received_crc (crc.rs:10-40,
// demod.rs:70-74); the repair of crc.rs:49-65 is "find p < 88 with syn[p] == that".
struct SynTable {
    uint32_t v[112];
};
constexpr uint32_t mul_x_mod_g(uint32_t r)
{
    return (r & 0x800000u) ? (((r << 1) ^ 0xFFF409u) & 0xFFFFFFu) : ((r << 1) & 0xFFFFFFu);
}
constexpr SynTable make_syn()
{
    SynTable t{};
    uint32_t r = 1;
    for (int e = 0; e < 112; ++e) {
        t.v[111 - e] = r;
        r = mul_x_mod_g(r);
    }
    return t;
}
// compile-time known answers (T[0] = 0x3935EA ... T[87] = 0xFFF409; a parity bit is its own syndrome)
static_assert(make_syn().v[0] == 0x3935EAu && make_syn().v[87] == 0xFFF409u, "CRC-24 single-bit syndromes (crc.rs:10-40)");
static_assert(make_syn().v[88] == 0x800000u && make_syn().v[111] == 0x000001u, "parity-bit syndromes");

// The slicer gives lane l the frame bits 31 - l + 32 r (r = 0..3), so a lane needs exactly four
// syndromes: one 16-byte load.  Entry .w is zero for the lanes that have no bit in round 3.
struct SynLanes {
    uint4 v[32];
};
constexpr SynLanes make_syn_lanes()
{
    const SynTable t = make_syn();
    SynLanes s{};
    for (int l = 0; l < 32; ++l) {
        s.v[l].x = t.v[31 - l];
        s.v[l].y = t.v[63 - l];
        s.v[l].z = t.v[95 - l];
        s.v[l].w = l >= 16 ? t.v[127 - l] : 0u;
    }
    return s;
}
static_assert(make_syn_lanes().v[31].x == 0x3935EAu && make_syn_lanes().v[8].z == 0xFFF409u && make_syn_lanes().v[15].w == 0u &&
              make_syn_lanes().v[16].w == 0x000001u, "lane layout of the syndromes: frame bit 32 r + 31 - lane");
__device__ const SynLanes g_syn_lanes = make_syn_lanes();

// ---- per-sample level -------------------------------------------------------
// U8: one 32-bit word = (I0, Q0, I1, Q1).  Returns (level0 | level1 << 16).
__device__ __forceinline__ uint32_t levels_u8_pair(uint32_t w, uint32_t minus_one)
{
    // z = I*(255-I) + Q*(255-Q) <= 32512 per sample.  The integer ALU pipe is the busy
    // one in this kernel, so the complement is taken as w * -1 + -1 (IMAD, FMA pipe):
    // with `minus_one` = 0xFFFFFFFF passed as a kernel parameter so that ptxas cannot turn
    // it back into an integer-pipe negate
    const uint32_t nw = w * minus_one + minus_one;            // ~w
    const uint32_t zsum = __dp4a(nw, w, 0u);                  // z0 + z1
    const uint32_t z1 = __dp4a(nw, w & 0xFFFF0000u, 0u);      // z1
    return zsum + z1 * 65535u;                                // z0 + (z1 << 16)
}

// U8, two streams: wa = (I, Q, I', Q') of samples (s, s+1), wb = the same for samples (1024+s, 1024+s+1).
// out0 = level[s] | level[1024+s] << 16, out1 = level[s+1] | level[1024+s+1] << 16.  Same instruction
// count per sample as levels_u8_pair (4 IMAD + 4 IDP.4A + 2 LOP3 per four samples).
__device__ __forceinline__ void levels_u8_streams(uint32_t wa, uint32_t wb, uint32_t minus_one, uint32_t &out0, uint32_t &out1)
{
    const uint32_t nwa = wa * minus_one + minus_one;            // ~wa on the FMA pipe
    const uint32_t nwb = wb * minus_one + minus_one;
    const uint32_t za1 = __dp4a(nwa, wa & 0xFFFF0000u, 0u);
    const uint32_t zb1 = __dp4a(nwb, wb & 0xFFFF0000u, 0u);
    out1 = zb1 * 65536u + za1;
    const uint32_t sumb = __dp4a(nwb, wb, 0u);                  // zb0 + zb1
    out0 = __dp4a(nwa, wa, sumb * 65536u - out1);               // (za0 + za1) + (zb0 + zb1) << 16 - out1
    // (measured alternative without the two masks -- sample 0 isolated by shifting both operands up
    // 16 bits with multiplies, +2 FMA-pipe instructions per call -- was slower: 4.66 vs 4.36 ms)
}

// CS16: one 32-bit word = (re, im) little-endian i16.  Returns -floor(sqrt(re^2+im^2)) (mod 2^32), exactly;
// the level is 0xFFFF + that.
//   n = re^2 + im^2 <= 2^31.  MUFU sqrt of float(n) is within a relative 2^-21 of the true root s; scaled UP by
//   (1 + 2^-20) it lies in [s, s + 1), so its integer part r is isqrt(n) or isqrt(n) + 1, and the sign of
//   n - r^2 (|n - r^2| < 2^17: no wrap-around) says which.
// Instruction budget (tools/ubench3.cu: F2I, I2F.S16, POPC share the 16-lane XU pipe with MUFU; ISETP, SEL, LOP3
// the 16-lane integer pipe): the integer part is taken with one FFMA.RZ against 2^23 (the sum's ulp is 1, so
// round-toward-zero IS the floor) instead of FADD + F2I, and the fix-up is IMADs plus one shift instead of
// IMAD + ISETP + SEL.  `one` / `minus_one` are kernel parameters so that ptxas keeps the additions on the FMA pipe.
__device__ __forceinline__ uint32_t neg_isqrt_cs16(uint32_t w, uint32_t minus_one, uint32_t one)
{
    // n = re^2 + im^2 without unpacking: a 16-bit value is lowbyte (unsigned) + 256 * highbyte (signed), so with
    // bp = (re.lo, im.lo, re.hi, im.hi) two 16 x 8-bit dot products give (re, im).(re.lo, im.lo) and
    // (re, im).(re.hi, im.hi): one PRMT on the integer pipe, the rest on the FMA pipe
    const uint32_t bp = __byte_perm(w, 0u, 0x3120);
    int hi, lo;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(hi) : "r"(w), "r"(bp), "r"(0));
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(lo) : "r"(w), "r"(bp), "r"(hi * 256));
    const uint32_t n = (uint32_t)lo;
    float f, t;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(__uint2float_rz(n)));
    asm("fma.rz.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(f), "f"(1.00000095367431640625f), "f"(8388608.0f));
    const uint32_t q = __float_as_uint(t);                       // 0x4B000000 + r
    uint32_t r, nr, d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(q), "r"(one), "r"(0xB5000000u));          // q - 0x4B000000
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(nr) : "r"(q), "r"(minus_one), "r"(0x4B000000u));   // -r
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(r), "r"(nr), "r"(n));                     // n - r^2
    uint32_t u;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(d >> 31), "r"(one), "r"(nr));             // -(r - [n < r^2])
    return u;
}

// two samples of the two streams -> one word (level of stream 0 | level of stream 1 << 16), level = 0xFFFF - isqrt
__device__ __forceinline__ uint32_t level_word_cs16(uint32_t wa, uint32_t wb, uint32_t minus_one, uint32_t one)
{
    const uint32_t ua = neg_isqrt_cs16(wa, minus_one, one), ub = neg_isqrt_cs16(wb, minus_one, one);
    // (0xFFFF + ua) + ((0xFFFF + ub) << 16) = ua + (ub << 16) - 1  (mod 2^32)
    uint32_t v;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(ub), "r"(65536u), "r"(ua));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(v), "r"(one), "r"(minus_one));
    return v;
}

// Eight words of the tile from eight samples of each stream.
// U8: a0 / b0 = the 16 bytes of stream 0 / 1 (a1, b1 unused); CS16: a0,a1 / b0,b1 = 2 x 16 bytes each.
template <int FMT>
__device__ __forceinline__ void words_of_chunk8(uint4 a0, uint4 a1, uint4 b0, uint4 b1, uint32_t minus_one, uint4 &o0, uint4 &o1)
{
    if (FMT == AIRGPU_FMT_U8) {
        levels_u8_streams(a0.x, b0.x, minus_one, o0.x, o0.y);
        levels_u8_streams(a0.y, b0.y, minus_one, o0.z, o0.w);
        levels_u8_streams(a0.z, b0.z, minus_one, o1.x, o1.y);
        levels_u8_streams(a0.w, b0.w, minus_one, o1.z, o1.w);
    } else {
        const uint32_t one = 0u - minus_one;
        o0.x = level_word_cs16(a0.x, b0.x, minus_one, one);
        o0.y = level_word_cs16(a0.y, b0.y, minus_one, one);
        o0.z = level_word_cs16(a0.z, b0.z, minus_one, one);
        o0.w = level_word_cs16(a0.w, b0.w, minus_one, one);
        o1.x = level_word_cs16(a1.x, b1.x, minus_one, one);
        o1.y = level_word_cs16(a1.y, b1.y, minus_one, one);
        o1.z = level_word_cs16(a1.z, b1.z, minus_one, one);
        o1.w = level_word_cs16(a1.w, b1.w, minus_one, one);
    }
}

__device__ __forceinline__ uint4 ldg_stream(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 16 bytes at byte offset `off`, zero beyond `avail` bytes; any alignment (edge tiles only).
__device__ __noinline__ uint4 load16_guarded(const uint8_t *src, long long off, long long avail)
{
    unsigned long long lo = 0ull, hi = 0ull;     // no indexed local array: the kernel keeps a zero-byte stack frame
#pragma unroll 1
    for (int b = 0; b < 8; ++b) {
        if (off + b < avail) lo |= (unsigned long long)src[off + b] << (8 * b);
        if (off + b + 8 < avail) hi |= (unsigned long long)src[off + b + 8] << (8 * b);
    }
    return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

// DF = 17 test on the first five data bits (demod.rs:45-54), inverted levels, candidate = word xw
// of stream st.  Words xw+16 .. xw+25 are consecutive in shared memory except that one pad (4 words
// = 16 bytes) may fall inside the run (28 % of the hits): level k is read from one of two bases,
// 16 bytes apart, chosen by k >= cross, so that lanes with and without a pad run the same code.
__device__ __forceinline__ bool df17_ok(const uint16_t *s16, int xw, int st)
{
    const uint16_t *qa = s16 + level_index(xw, st, 16);
    uint32_t v[10];
    const int cross = df_cross(xw);            // first k behind the pad (>= 10: none)
    const uint16_t *qb = qa + 8;
#pragma unroll
    for (int k = 0; k < 10; ++k) v[k] = (k >= cross ? qb : qa)[2 * k];
    const uint32_t hi = max(max(max(v[0], v[3]), max(v[5], v[7])), v[8]);
    const uint32_t lo = min(min(min(v[1], v[2]), min(v[4], v[6])), v[9]);
    return hi <= lo;
}

struct Cand {
    uint32_t w0, w1, w2, w3;   // frame bits, big-endian words (bit 31 of w0 = frame bit 0)
    uint32_t fixed;            // 0xFF or repaired bit
    bool valid;
};

// Warp-cooperative slice + CRC + repair for the candidate at word xw of stream st
// (demod.rs:65-82).  All lanes return the same value.  Lane l owns frame bits 31 - l + 32 r, so
// the four ballots ARE the big-endian frame words.
__device__ __forceinline__ Cand process_candidate(const uint16_t *s, int xw, int st, int lane)
{
    Cand c;
    const int wj = slicer_word(xw, lane);
    const uint16_t *p0 = s + 2 * phys_word(wj) + st;
    const uint16_t *p1 = p0 + slicer_step(wj);
    const bool tail = lane >= 16;                                 // round 3 only has bits 96..111
    const uint4 syn_of = __ldg(&g_syn_lanes.v[lane]);
    // m[2k] > m[2k+1]  (demod.rs:104), inverted levels
    const bool b0 = p0[0] < p1[0], b1 = p0[kSlicerRoundStride] < p1[kSlicerRoundStride],
               b2 = p0[2 * kSlicerRoundStride] < p1[2 * kSlicerRoundStride];
    bool b3 = false;
    if (tail) b3 = p0[3 * kSlicerRoundStride] < p1[3 * kSlicerRoundStride];
    c.w0 = __ballot_sync(kFull, b0);
    c.w1 = __ballot_sync(kFull, b1);
    c.w2 = __ballot_sync(kFull, b2);
    c.w3 = __ballot_sync(kFull, b3);
    const uint32_t part = (b0 ? syn_of.x : 0u) ^ (b1 ? syn_of.y : 0u) ^ (b2 ? syn_of.z : 0u) ^ (b3 ? syn_of.w : 0u);
    const uint32_t syn = __reduce_xor_sync(kFull, part);
    c.fixed = 0xFFu;
    c.valid = true;
    if (syn != 0u) {
        // crc.rs:49-65: only a flip of one of the 88 data bits can match (round 2: bits 64..87 = lanes 8..31).
        // Nearly every non-zero syndrome is unrepairable: one ballot settles that.
        const bool h0 = syn_of.x == syn, h1 = syn_of.y == syn, h2 = lane >= 8 && syn_of.z == syn;
        const unsigned any = __ballot_sync(kFull, h0 || h1 || h2);
        if (any == 0u) {
            c.valid = false;
        } else {
            // the syndromes are distinct: exactly one (lane, round) matches.  Frame bit = 32 r + 31 - lane.
            const int l = __ffs(any) - 1;
            const int r = __shfl_sync(kFull, h0 ? 0 : (h1 ? 1 : 2), l);
            c.fixed = 32 * r + 31 - l;
            const uint32_t flip = 1u << l;
            if (r == 0) c.w0 ^= flip;
            else if (r == 1) c.w1 ^= flip;
            else c.w2 ^= flip;
        }
    }
    return c;
}

// Where the frames a warp finds go.  Every tile owns kSlotsPerTile fixed 32-byte slots in scratch
// (slot index = tile * kSlotsPerTile): the four big-endian frame words, then (offset within the
// tile | fixed_bit << 16).  The common case therefore needs no reservation, no atomic with a return
// value and no staging, and the fast path may fill the slots in ANY order -- gather_kernel ranks
// the (at most kSlotsPerTile) records of a tile by offset.  A tile with more frames than slots
// (degenerate input such as a constant buffer) is redone by the ordered path, which writes every
// record in ascending offset order: the first kSlotsPerTile into the slots, the rest into an
// overflow range reserved with one atomic.
struct Sink {
    uint4 *slots;                   // this tile's fixed slots in scratch (2 x uint4 per slot)
    uint4 *overflow;                // overflow range, nullptr on the fast path
    unsigned long long ovf_room;    // records that fit the overflow range
    uint32_t seq;                   // valid frames seen so far in this pass
    uint32_t gate;                  // gate passes (reference num_processed)
};

// One survivor of the gate: slice, CRC, repair, emit (warp-cooperative, warp-uniform candidate).
__device__ __forceinline__ void emit_candidate(const uint16_t *lv, int i, int lane, Sink &sink)
{
    const Cand c = process_candidate(lv, i & (kStream - 1), i >> 10, lane);
    if (!c.valid) return;
    uint4 *dst = nullptr;
    if (sink.seq < (uint32_t)kSlotsPerTile) {
        dst = sink.slots + 2 * sink.seq;
    } else if (sink.overflow != nullptr) {
        const unsigned long long r = sink.seq - kSlotsPerTile;
        if (r < sink.ovf_room) dst = sink.overflow + 2 * r;
    }
    if (lane == 0 && dst != nullptr) {
        dst[0] = make_uint4(c.w0, c.w1, c.w2, c.w3);
        reinterpret_cast<uint32_t *>(dst + 1)[0] = (uint32_t)i | (c.fixed << 16);
    }
    sink.seq += 1;
}

// The lane's 48 words of the tile -> one bit per owned offset that passes the preamble test.
template <int FMT>
__device__ __forceinline__ void gate_of_lane(const uint16_t *lv, int lane, uint32_t minus_one, uint32_t (&pm)[2])
{
    // lane owns offsets x0 .. x0+31 of both streams, x0 = 32 * lane: words x0 .. x0+46, i.e.
    // 16-byte chunks 8*lane .. 8*lane+11, which sit at padded chunks 9*lane + k + (k >> 3)
    const uint4 *lv4 = reinterpret_cast<const uint4 *>(lv) + 9 * lane;
    uint32_t R[48];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const uint4 v = lv4[k + (k >> 3)];
        R[4 * k + 0] = v.x;
        R[4 * k + 1] = v.y;
        R[4 * k + 2] = v.z;
        R[4 * k + 3] = v.w;
    }
    gate_scan<FMT == AIRGPU_FMT_U8>(R, pm, minus_one);
}

// Gate + slice + CRC over the warp's candidates [0, wcands): the reference's offset loop
// (adsb.rs:98-114) for this range.  FAST PATH: records are emitted in no particular order.
//
// The preamble test runs over the whole tile first (one straight-line pass, no divergence); each
// lane only keeps WHICH of its 64 offsets passed, as bits.  Then, in rounds, every lane pops one of
// its hits and runs the DF test on it (all lanes at once), and the survivors of the round are
// sliced and checked one after the other by the whole warp.  A round costs the same whether one
// lane or all of them have a hit; in dense traffic a tile has ~6 hits in ~1.3 rounds.
template <int FMT>
__device__ __forceinline__ void scan_tile_fast(const uint16_t *lv, int wcands, int lane, uint32_t minus_one, Sink &sink)
{
    uint32_t pm[2];
    gate_of_lane<FMT>(lv, lane, minus_one, pm);
    uint32_t a = pm[0], b = pm[1];
    for (;;) {
        const bool from_a = a != 0u;
        const uint32_t w = from_a ? a : b;
        if (__ballot_sync(kFull, w != 0u) == 0u) break;
        int i = 0;
        bool ok = false;
        if (w != 0u) {
            const int bit = __ffs(w) - 1;
            const uint32_t rest = w & (w - 1u);
            if (from_a) a = rest;
            else b = rest;
            const int xw = lane * kLaneX + hit_x(from_a ? 0 : 1, bit), st = hit_stream(bit);
            i = st * kStream + xw;
            ok = i < wcands && df17_ok(lv, xw, st);
        }
        unsigned surv = __ballot_sync(kFull, ok);
        sink.gate += __popc(surv);
        while (surv) {
            const int src = __ffs(surv) - 1;
            surv &= surv - 1u;
            emit_candidate(lv, __shfl_sync(kFull, i, src), lane, sink);
        }
    }
}

// The same range in ASCENDING OFFSET ORDER (adsb.rs:98): only for tiles with more frames than
// fixed slots, i.e. degenerate input.  Not inlined: it must stay out of the hot loop's code.
template <int FMT>
__device__ __noinline__ unsigned long long scan_tile_ordered(const uint16_t *lv, int wcands, int lane, uint32_t minus_one,
                                                             uint4 *slots, uint4 *overflow, unsigned long long ovf_room)
{   // everything by value (a Sink passed by reference would live on the stack); returns gate passes << 32 | frames
    Sink sink;
    sink.slots = slots;
    sink.overflow = overflow;
    sink.ovf_room = ovf_room;
    sink.seq = 0;
    sink.gate = 0;
    uint32_t pm[2];
    gate_of_lane<FMT>(lv, lane, minus_one, pm);
    uint32_t cm[2] = {0u, 0u};   // cm[s] bit x: offset s*1024 + lane*32 + x passes the gate
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        uint32_t m = half ? pm[1] : pm[0];
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int st = hit_stream(b), x = hit_x(half, b);
            const int xw = lane * kLaneX + x;
            if (st * kStream + xw < wcands && df17_ok(lv, xw, st)) {
                if (st) cm[1] |= 1u << x;
                else cm[0] |= 1u << x;
            }
        }
    }
    // survivors, in ascending offset order: stream, then lane, then bit
#pragma unroll 1
    for (int st = 0; st < 2; ++st) {
        const uint32_t mine = st ? cm[1] : cm[0];
        unsigned lanes = __ballot_sync(kFull, mine != 0u);
        while (lanes) {
            const int src_lane = __ffs(lanes) - 1;
            lanes &= lanes - 1;
            uint32_t bits = __shfl_sync(kFull, mine, src_lane);
            sink.gate += __popc(bits);
            while (bits) {
                const int o = __ffs(bits) - 1;
                bits &= bits - 1;
                emit_candidate(lv, st * kStream + src_lane * kLaneX + o, lane, sink);
            }
        }
    }
    return ((unsigned long long)sink.gate << 32) | sink.seq;
}

// ---- one tile = kWarpTile candidate offsets, done by one warp in its private slice of shared memory ----

// Geometry of a tile.  `rem` = samples from this warp's first sample to the end of its segment;
// everything else follows from it.  All tile starts are multiples of 2048 samples, so 16-byte
// alignment of the loads is a per-launch property (p.vec_ok, computed by the host).  Tiles below
// p.full_tiles (single-segment launches: all but the last two) are complete and need none of the
// 64-bit arithmetic.
struct TileGeom {
    const uint8_t *src;          // first sample of the tile
    unsigned long long rem;      // samples to the end of the segment
    int wcands;                  // candidate offsets of this tile: min(kWarpTile, rem - 240)
};

template <int FMT, bool kSingleSegment>
__device__ __forceinline__ TileGeom tile_geometry(const DecodeParams &p, const unsigned tile)
{
    constexpr int BPS = (FMT == AIRGPU_FMT_U8) ? 2 : 4;
    unsigned long long seg_start = 0, wpos = (unsigned long long)tile * kWarpTile;
    TileGeom g;
    g.rem = (unsigned long long)(kStream + kTileWords);
    g.wcands = kWarpTile;
    if (!kSingleSegment || tile >= p.full_tiles) {
        unsigned seg = 0, tile_in_seg = tile;
        if (!kSingleSegment) {
            seg = tile / p.tiles_per_seg;
            tile_in_seg = tile - seg * p.tiles_per_seg;
        }
        seg_start = kSingleSegment ? 0ull : (unsigned long long)seg * p.seg_len;
        const unsigned long long seg_n = kSingleSegment ? p.n_samples : min(p.seg_len, p.n_samples - seg_start);
        wpos = (unsigned long long)tile_in_seg * kWarpTile;
        g.rem = seg_n > wpos ? seg_n - wpos : 0ull;
        g.wcands = g.rem > (unsigned long long)kFrameSamples
                       ? (int)min((unsigned long long)kWarpTile, g.rem - kFrameSamples) : 0;
    }
    g.src = static_cast<const uint8_t *>(p.iq) + (seg_start + wpos) * BPS;
    return g;
}

constexpr int kRounds = (kTileChunks8 + 31) / 32;           // 5 rounds of 32 units; the last has 30

// ---- phase 1: IQ -> inverted levels in the warp's shared-memory slice ----
// Unit of work: 8 words of the tile = 8 samples of stream 0 (tile samples 8c ..) and 8 samples
// of stream 1 (1024 + 8c ..).  Words 1024 .. 1263 repeat, in their low halves, levels that
// words 0 .. 239 hold in their high halves: those loads hit L2.
template <int FMT>
__device__ __forceinline__ void load_levels(const DecodeParams &p, const TileGeom &g, const int lane, uint16_t *lv)
{
    constexpr int BPS = (FMT == AIRGPU_FMT_U8) ? 2 : 4;
    constexpr int kChunkBytes = 8 * BPS;
    constexpr int kStreamBytes = kStream * BPS;
    const uint8_t *src = g.src;
    if (p.vec_ok && g.rem >= (unsigned long long)(kStream + kTileWords)) {
        // unit c = lane + 32 m: global address and shared address are both "per-lane base +
        // compile-time offset" (padded 16-byte chunks 2c + (c >> 2) = 2 lane + (lane >> 2) + 72 m)
        const uint8_t *gsrc = src + lane * kChunkBytes;
        uint4 *sdst = reinterpret_cast<uint4 *>(lv) + (2 * lane + (lane >> 2));
        if (FMT == AIRGPU_FMT_CS16) {
            // rolled (fully unrolled it is far too much code for the instruction cache)
            uint4 a0 = ldg_stream(gsrc), a1 = ldg_stream(gsrc + 16);
            uint4 b0 = ldg_stream(gsrc + kStreamBytes), b1 = ldg_stream(gsrc + kStreamBytes + 16);
#pragma unroll 1
            for (int m = 0; m < kRounds; ++m) {
                uint4 na0 = a0, na1 = a1, nb0 = b0, nb1 = b1;      // (never converted when not reloaded)
                if (m + 1 < kRounds && lane + 32 * (m + 1) < kTileChunks8) {
                    const uint8_t *gn = gsrc + 32 * (m + 1) * kChunkBytes;
                    na0 = ldg_stream(gn);
                    na1 = ldg_stream(gn + 16);
                    nb0 = ldg_stream(gn + kStreamBytes);
                    nb1 = ldg_stream(gn + kStreamBytes + 16);
                }
                if (lane + 32 * m < kTileChunks8) {
                    uint4 o0, o1;
                    words_of_chunk8<FMT>(a0, a1, b0, b1, p.minus_one, o0, o1);
                    sdst[72 * m] = o0;
                    sdst[72 * m + 1] = o1;
                }
                a0 = na0;
                a1 = na1;
                b0 = nb0;
                b1 = nb1;
            }
        } else {
            // all ten loads of the lane are issued before the first conversion: a warp waits for
            // HBM once per tile
            uint4 a[kRounds], b[kRounds];
#pragma unroll
            for (int m = 0; m < kRounds; ++m) {
                if (32 * m + 31 < kTileChunks8 || lane + 32 * m < kTileChunks8) {
                    a[m] = ldg_stream(gsrc + 32 * m * kChunkBytes);
                    b[m] = ldg_stream(gsrc + 32 * m * kChunkBytes + kStreamBytes);
                }
            }
#pragma unroll
            for (int m = 0; m < kRounds; ++m) {
                if (32 * m + 31 < kTileChunks8 || lane + 32 * m < kTileChunks8) {
                    uint4 o0, o1;
                    words_of_chunk8<FMT>(a[m], a[m], b[m], b[m], p.minus_one, o0, o1);
                    sdst[72 * m] = o0;
                    sdst[72 * m + 1] = o1;
                }
            }
        }
    } else {
        // edge of a segment or an unaligned buffer: same arithmetic, guarded byte loads
        const long long avail = (long long)g.rem * BPS;   // bytes to the end of the segment
#pragma unroll 1
        for (int c = lane; c < kTileChunks8; c += 32) {
            const long long oa = (long long)c * kChunkBytes, ob = oa + kStreamBytes;
            uint4 a0 = load16_guarded(src, oa, avail), b0 = load16_guarded(src, ob, avail), a1 = a0, b1 = b0;
            if (FMT == AIRGPU_FMT_CS16) {
                a1 = load16_guarded(src, oa + 16, avail);
                b1 = load16_guarded(src, ob + 16, avail);
            }
            uint4 o0, o1;
            words_of_chunk8<FMT>(a0, a1, b0, b1, p.minus_one, o0, o1);
            uint4 *d = reinterpret_cast<uint4 *>(lv) + phys_chunk4(2 * c);
            d[0] = o0;
            d[1] = o1;
        }
    }
}

// ---- phase 2..4: gate, slice, CRC; frames go straight to the tile's scratch slots; publish the count ----
template <int FMT>
__device__ __forceinline__ void finish_tile(const DecodeParams &p, const unsigned tile, const int wcands, const int lane,
                                            const uint16_t *lv)
{
    uint4 *scratch = reinterpret_cast<uint4 *>(p.scratch);
    Sink sink;
    sink.slots = scratch + (unsigned long long)tile * (kSlotsPerTile * 2);
    sink.overflow = nullptr;
    sink.ovf_room = 0;
    sink.seq = 0;
    sink.gate = 0;
    uint32_t nvalid, gate;
    if (!p.force_ordered) {
        scan_tile_fast<FMT>(lv, wcands, lane, p.minus_one, sink);
        nvalid = sink.seq;
        gate = sink.gate;
    } else {                                                               // tests only
        const unsigned long long r = scan_tile_ordered<FMT>(lv, wcands, lane, p.minus_one, sink.slots, nullptr, 0ull);
        nvalid = (uint32_t)r;
        gate = (uint32_t)(r >> 32);
    }
    unsigned long long ovf_base = 0;
    if (nvalid > (uint32_t)kSlotsPerTile) {
        // rare (degenerate input): redo the range in ascending offset order; frames kSlotsPerTile..
        // go to the overflow area, which starts after all the fixed slots
        if (lane == 0) ovf_base = atomicAdd(p.ovf_counter, (unsigned long long)(nvalid - kSlotsPerTile));
        ovf_base = __shfl_sync(kFull, ovf_base, 0);
        (void)scan_tile_ordered<FMT>(lv, wcands, lane, p.minus_one, sink.slots,
                                     scratch + ((unsigned long long)p.n_tiles * kSlotsPerTile + ovf_base) * 2,
                                     p.ovf_cap > ovf_base ? p.ovf_cap - ovf_base : 0ull);
    }
    if (lane == 0) {
        p.tile_tab[tile] = make_uint2((unsigned)min(ovf_base, 0xFFFFFFFFull), nvalid);
        // per-group sums, gate passes << 32 | frames: one fire-and-forget RED (no atomic with a return value)
        if (gate) atomicAdd(&p.group_sum[tile / kGroupTiles], ((unsigned long long)gate << 32) | nvalid);
    }
}

template <int FMT, bool kSingleSegment>
__device__ __forceinline__ void decode_tile(const DecodeParams &p, const unsigned tile, const int lane, uint16_t *lv)
{
    const TileGeom g = tile_geometry<FMT, kSingleSegment>(p, tile);
    if (g.wcands == 0) {
        if (lane == 0) p.tile_tab[tile] = make_uint2(0u, 0u);
        return;
    }
    load_levels<FMT>(p, g, lane, lv);
    __syncwarp();
    finish_tile<FMT>(p, tile, g.wcands, lane, lv);
}

#ifndef AIRGPU_MIN_CTAS
#define AIRGPU_MIN_CTAS 8
#endif
#ifndef AIRGPU_MIN_CTAS_CS16
#define AIRGPU_MIN_CTAS_CS16 4       // ptxas then settles on ~70 registers without spills (7 CTAs per SM still fit); a cap of 64 or 72 spills
#endif
template <int FMT, bool kSingleSegment>
__global__ void __launch_bounds__(kThreads, FMT == AIRGPU_FMT_U8 ? AIRGPU_MIN_CTAS : AIRGPU_MIN_CTAS_CS16)
decode_kernel(const DecodeParams p)
{
    // Warps never talk to each other: each owns one tile at a time, a private slice of shared
    // memory and its own output slots.  The CTA is only a packaging unit (4 warps keep the
    // per-CTA footprint small: 8 CTAs per SM).  A warp does p.tiles_per_warp tiles in a row
    // (a CTA covers 4 * tiles_per_warp consecutive tiles, the four warps always on neighbouring
    // ones): a warp's start-up (special registers, parameter loads, CTA launch) took 20 % of its
    // life with one tile per warp.
    __shared__ __align__(128) uint16_t s_lvl[kWarps][2 * kTileWordsPadded];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned tile = blockIdx.x * (kWarps * p.tiles_per_warp) + warp;
    // (the bound is recomputed from the block index every round -- the volatile asm keeps ptxas from
    // hoisting it: with 64 registers a loop counter or a hoisted bound is spilled to local memory)
#pragma unroll 1
    for (;;) {
        unsigned cta;
        asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(cta));
        if (tile >= min(p.n_tiles, (cta + 1) * (kWarps * p.tiles_per_warp))) break;
        decode_tile<FMT, kSingleSegment>(p, tile, lane, s_lvl[warp]);
        tile += kWarps;
        __syncwarp();    // every lane is done reading the slice before the next tile overwrites it
    }
}

// ---- A/B variant (AIRGPU_STAGE=1): the same path with the raw IQ of a warp's NEXT tile staged in
// shared memory by the TMA unit (1-D cp.async.bulk completing on an mbarrier) while the warp runs
// the gate and the survivor path of the current one.  The 2288 samples a tile reads are one
// contiguous 4576-byte range (U8), so one bulk copy per tile, issued by lane 0, replaces the ten
// LDG.128 per lane and their forty registers; the price is 4.6 KB more shared memory per warp
// (5 CTAs = 20 warps per SM instead of 32).  U8, single segment, aligned buffers only; every
// other launch and the ragged last tiles take the loader above.  Measured: DESIGN.md section 5.
constexpr int kRawBytesU8 = (kStream + kTileWords) * 2;      // 4576: a multiple of 16
constexpr int kRawPitch = 4608;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_issue(uint32_t dst, const void *src, uint32_t bar)
{
    // generic-proxy reads of the buffer (phase 1 of the previous tile) are ordered before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kRawBytesU8) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"((uint32_t)kRawBytesU8), "r"(bar) : "memory");
}

__device__ __forceinline__ void stage_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

#ifndef AIRGPU_STAGED_MIN_CTAS
#define AIRGPU_STAGED_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(kThreads, AIRGPU_STAGED_MIN_CTAS)
decode_kernel_staged_u8(const DecodeParams p)
{
    constexpr int FMT = AIRGPU_FMT_U8;
    __shared__ __align__(128) uint16_t s_lvl[kWarps][2 * kTileWordsPadded];
    __shared__ __align__(128) uint8_t s_raw[kWarps][kRawPitch];
    __shared__ __align__(8) unsigned long long s_bar[kWarps];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint16_t *lv = s_lvl[warp];
    const uint32_t raw = smem_u32(s_raw[warp]), bar = smem_u32(&s_bar[warp]);
    unsigned tile = blockIdx.x * (kWarps * p.tiles_per_warp) + warp;
    const unsigned bound = min(p.n_tiles, (blockIdx.x + 1) * (kWarps * p.tiles_per_warp));
    if (tile >= bound) return;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // complete tiles (tile < p.full_tiles) are staged; the launcher guarantees p.vec_ok
    bool staged = tile < p.full_tiles;
    if (staged && lane == 0) stage_issue(raw, static_cast<const uint8_t *>(p.iq) + (unsigned long long)tile * (kWarpTile * 2), bar);
    uint32_t parity = 0;
#pragma unroll 1
    for (; tile < bound; tile += kWarps) {
        const unsigned next = tile + kWarps;
        const bool next_staged = next < bound && next < p.full_tiles;
        int wcands = kWarpTile;
        if (staged) {
            stage_wait(bar, parity);
            parity ^= 1u;
            // unit c = lane + 32 m reads 16 bytes of each stream from the staged tile (consecutive lanes,
            // consecutive 16-byte chunks: conflict free) and stores its two chunks of words as the loader does
            const uint4 *rsrc = reinterpret_cast<const uint4 *>(s_raw[warp]) + lane;
            uint4 *sdst = reinterpret_cast<uint4 *>(lv) + (2 * lane + (lane >> 2));
#pragma unroll
            for (int m = 0; m < kRounds; ++m) {
                if (32 * m + 31 < kTileChunks8 || lane + 32 * m < kTileChunks8) {
                    const uint4 a = rsrc[32 * m], b = rsrc[32 * m + kStream * 2 / 16];
                    uint4 o0, o1;
                    words_of_chunk8<FMT>(a, a, b, b, p.minus_one, o0, o1);
                    sdst[72 * m] = o0;
                    sdst[72 * m + 1] = o1;
                }
            }
            __syncwarp();      // every lane has read the staged bytes and written its words
        } else {
            const TileGeom g = tile_geometry<FMT, true>(p, tile);
            wcands = g.wcands;
            if (wcands) load_levels<FMT>(p, g, lane, lv);
            __syncwarp();
        }
        // the staging buffer is free again: the next tile's bytes arrive during the gate and the survivor path
        if (next_staged && lane == 0)
            stage_issue(raw, static_cast<const uint8_t *>(p.iq) + (unsigned long long)next * (kWarpTile * 2), bar);
        if (wcands == 0) {
            if (lane == 0) p.tile_tab[tile] = make_uint2(0u, 0u);
        } else {
            finish_tile<FMT>(p, tile, wcands, lane, lv);
        }
        staged = next_staged;
        __syncwarp();    // every lane is done reading the slice before the next tile overwrites it
    }
}

// ---- ordering: two-level exclusive scan of the per-tile counts, then gather ------
// decode_kernel already added every tile's count into group_sum[tile / kGroupTiles];
// group_scan_kernel turns those few sums into bases (one CTA), and gather_kernel re-scans
// the counts of its own group in shared memory and copies the records to their final,
// globally ordered position.  Both are tiny next to the decode kernel at any capture size.
constexpr int kScanThreads = 1024;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *s_warp,
                                                                   unsigned long long *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(kFull, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long wsum = lane < nwarps ? s_warp[lane] : 0ull;
        unsigned long long y = wsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long z = __shfl_up_sync(kFull, y, d);
            if (lane >= d) y += z;
        }
        if (lane < nwarps) s_warp[lane] = y - wsum;
        if (lane == 31) s_warp[32] = y;
    }
    __syncthreads();
    const unsigned long long excl = s_warp[warp] + x - v;
    *total = s_warp[32];
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(kScanThreads, 1)
group_scan_kernel(const unsigned long long *group_sum, unsigned n_groups, unsigned long long *group_base,
                  unsigned long long *d_total, unsigned long long *d_gate, const OutSet dst)
{
    __shared__ unsigned long long s_warp[33];
    unsigned long long running = *d_total;   // frames already in `out` (pieces of one call append)
    unsigned long long gate = 0;
    for (unsigned base = 0; base < n_groups; base += kScanThreads) {
        const unsigned g = base + threadIdx.x;
        const unsigned long long both = g < n_groups ? group_sum[g] : 0ull;    // gate passes << 32 | frames
        unsigned long long total;
        const unsigned long long excl = block_exclusive_scan(both & 0xFFFFFFFFull, s_warp, &total);
        if (g < n_groups) group_base[g] = running + excl;
        running += total;
        unsigned long long gsum;
        (void)block_exclusive_scan(both >> 32, s_warp, &gsum);
        gate += gsum;
    }
    if (threadIdx.x == 0) {
        *d_total = running;
        *d_gate += gate;                     // gate passes accumulate over the pieces of a call
        // frame exchange: the count travels with the records (peer-mapped or multicast addresses)
        if (dst.multicast) {
            if (dst.count[0]) multimem_st_u64(dst.count[0], running);
        } else {
            for (unsigned j = 0; j < dst.n; ++j)
                if (dst.count[j] && dst.count[j] != d_total) *dst.count[j] = running;
        }
    }
}

// Scratch slot -> airgpu_frame (three little-endian u64 words): byte-swapped frame words, then
// bytes 12, 13 | fixed_bit | reserved, then the absolute sample offset.
// kMode 0: one destination (plain decode); 1: dst.n destinations, plain stores to local / peer-mapped memory (the
// frame exchange over NVLink, fused: no copy engine, no second kernel); 2: one multimem.st per word to an NVSwitch
// multicast address -- the switch replicates it to every rank, so a rank's NVLink egress is 1x its list, not (n-1)x.
// Copy `n_words` staged u64 words to word index `w0` of every destination with the widest stores the alignment
// allows: a lone 8-byte word in front if needed, then 16-byte stores from consecutive threads (full 128-byte lines:
// what NVLink wants -- 8-byte stores at a 24-byte stride reached a fraction of the link rate), then a lone word.
//   kMode 0: one destination (plain decode); 1: dst.n destinations, plain stores to local / peer-mapped memory (the
//   frame exchange over NVLink, fused: no copy engine, no second kernel); 2: one multimem.st per 16 bytes to an
//   NVSwitch multicast address -- the switch replicates it to every rank, so a rank's NVLink egress is 1x its list.
template <int kMode>
__device__ __forceinline__ void copy_out(const OutSet &o, unsigned long long w0, const unsigned long long *s_words, unsigned n_words)
{
    const unsigned ndst = kMode == 1 ? o.n : 1u;
#pragma unroll 1
    for (unsigned j = 0; j < ndst; ++j) {
        unsigned long long *dst = o.out[j] + w0;
        const unsigned head = (unsigned)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1u) & (n_words ? 1u : 0u);   // words before 16-byte alignment
        const unsigned pairs = (n_words - head) >> 1;
        const unsigned tail = n_words - head - 2 * pairs;
        if (threadIdx.x == 0 && head) {
            if (kMode == 2) multimem_st_u64(dst, s_words[0]);
            else dst[0] = s_words[0];
        }
        if (threadIdx.x == 1 && tail) {
            if (kMode == 2) multimem_st_u64(dst + n_words - 1, s_words[n_words - 1]);
            else dst[n_words - 1] = s_words[n_words - 1];
        }
        for (unsigned k = threadIdx.x; k < pairs; k += blockDim.x) {
            const unsigned long long a = s_words[head + 2 * k], b = s_words[head + 2 * k + 1];
            unsigned long long *q = dst + head + 2 * k;
            if (kMode == 2) {
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "r"((uint32_t)a),
                             "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)) : "memory");
            } else {
                *reinterpret_cast<uint4 *>(q) = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
            }
        }
    }
}

constexpr unsigned kGatherChunk = 1536;      // records staged in shared memory at a time (36 KB)

template <int kMode>
__global__ void __launch_bounds__(kGroupTiles)
gather_kernel(const DecodeParams p, const unsigned long long *group_base, const OutSet out)
{
    // One CTA per group of kGroupTiles tiles: scan the counts in shared memory, then one THREAD per record of the
    // group (binary search for its tile) so that all record loads of the group are in flight at once.  The fast
    // path of decode_kernel fills a tile's slots in no particular order: a record's place among the (at most
    // kSlotsPerTile) records of its tile is the number of them with a smaller offset.  The finished records of the
    // group are staged in shared memory in output order -- they form ONE contiguous byte range of the output -- and
    // leave with wide coalesced stores.
    __shared__ unsigned long long s_warp[33];
    __shared__ unsigned int s_excl[kGroupTiles + 1];
    __shared__ unsigned int s_ovf[kGroupTiles];
    __shared__ __align__(16) unsigned long long s_words[3 * kGatherChunk];
    const unsigned t0 = blockIdx.x * kGroupTiles;
    const unsigned t = t0 + threadIdx.x;
    const uint2 e = t < p.n_tiles ? p.tile_tab[t] : make_uint2(0u, 0u);
    unsigned long long total64;
    const unsigned long long excl = block_exclusive_scan(e.y, s_warp, &total64);
    if (total64 == 0) return;
    s_excl[threadIdx.x] = (unsigned)excl;
    s_ovf[threadIdx.x] = e.x;
    if (threadIdx.x == 0) s_excl[kGroupTiles] = (unsigned)total64;
    __syncthreads();
    const unsigned long long gbase = group_base[blockIdx.x];
    if (gbase >= p.cap) return;
    // records beyond the output's capacity are counted (the caller learns the total) but not written
    const unsigned total = (unsigned)min(total64, p.cap - gbase);
    const uint4 *scratch = reinterpret_cast<const uint4 *>(p.scratch);
    for (unsigned c0 = 0; c0 < total; c0 += kGatherChunk) {
        const unsigned c1 = min(total, c0 + kGatherChunk);
        // a record's place differs from its slot by less than kSlotsPerTile (unordered tiles hold at most that many)
        const unsigned r0 = c0 >= (unsigned)kSlotsPerTile ? c0 - kSlotsPerTile : 0u;
        const unsigned r1 = (unsigned)min(total64, (unsigned long long)c1 + kSlotsPerTile);
        for (unsigned r = r0 + threadIdx.x; r < r1; r += kGroupTiles) {
            // largest k with s_excl[k] <= r (tiles with zero frames share their successor's value)
            unsigned lo = 0, hi = kGroupTiles;
            while (hi - lo > 1) {
                const unsigned mid = (lo + hi) >> 1;
                if (s_excl[mid] <= r) lo = mid;
                else hi = mid;
            }
            const unsigned idx = r - s_excl[lo];
            const unsigned n = s_excl[lo + 1] - s_excl[lo];
            const unsigned tile = t0 + lo;
            unsigned long long src;
            unsigned rank = idx;
            if (n <= (unsigned)kSlotsPerTile || idx < (unsigned)kSlotsPerTile) {      // a fixed slot
                src = (unsigned long long)tile * kSlotsPerTile + idx;
            } else {                                                                   // ordered tile, overflow area
                const unsigned long long o = (unsigned long long)s_ovf[lo] + (idx - kSlotsPerTile);
                if (o >= p.ovf_cap) continue;
                src = (unsigned long long)p.n_tiles * kSlotsPerTile + o;
            }
            const uint4 q = scratch[2 * src];
            const uint32_t meta = reinterpret_cast<const uint32_t *>(scratch + 2 * src + 1)[0];
            if (n <= (unsigned)kSlotsPerTile && n > 1u) {
                rank = 0;
                const uint4 *first = scratch + 2 * (unsigned long long)tile * kSlotsPerTile;
                for (unsigned k = 0; k < n; ++k)
                    rank += (reinterpret_cast<const uint32_t *>(first + 2 * k + 1)[0] & 0xFFFFu) < (meta & 0xFFFFu) ? 1u : 0u;
            }
            const unsigned d = s_excl[lo] + rank;
            if (d < c0 || d >= c1) continue;
            // first sample of the tile: segments are p.seg_len apart, tiles kWarpTile apart inside one
            const unsigned seg = tile / p.tiles_per_seg;
            const unsigned long long off0 = p.base_offset + (unsigned long long)seg * p.seg_len +
                                            (unsigned long long)(tile - seg * p.tiles_per_seg) * kWarpTile;
            // scratch slot -> airgpu_frame (three little-endian u64 words): byte-swapped frame words, then
            // bytes 12, 13 | fixed_bit | reserved, then the absolute sample offset
            const uint32_t hi32 = __byte_perm(q.w, meta >> 16, 0x7423);
            unsigned long long *w = s_words + 3 * (d - c0);
            w[0] = (unsigned long long)__byte_perm(q.x, 0, 0x0123) | ((unsigned long long)__byte_perm(q.y, 0, 0x0123) << 32);
            w[1] = (unsigned long long)__byte_perm(q.z, 0, 0x0123) | ((unsigned long long)hi32 << 32);
            w[2] = off0 + (meta & 0xFFFFu);
        }
        __syncthreads();
        copy_out<kMode>(out, (gbase + c0) * 3ull, s_words, 3u * (c1 - c0));
        __syncthreads();
    }
}

// ---- N1: frame fields -------------------------------------------------------------
// One thread per frame: three 8-byte loads, integer bit twiddling, two 16-byte stores.
// HBM bound by construction (24 B in, 32 B out per frame).
__global__ void __launch_bounds__(256) fields_kernel(const unsigned long long *frames, unsigned long long n,
                                                     uint4 *out)
{
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned long long w0 = frames[3 * k], w1 = frames[3 * k + 1];
    unsigned char b[14];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (unsigned char)(w0 >> (8 * i));
#pragma unroll
    for (int i = 0; i < 6; ++i) b[8 + i] = (unsigned char)(w1 >> (8 * i));
    const unsigned char *me = b + 4;                       // packet[4..11]
    const uint32_t icao = ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    const uint32_t df = b[0] >> 3, ca = b[0] & 5u, tc = me[0] >> 3;
    uint32_t kind = 0, alt = 0, lat = 0, lon = 0, flags = 0;
    unsigned long long cs = 0;
    if (tc >= 1 && tc <= 4) {                              // msgs.rs:208-213, 171-187
        kind = 1;
        unsigned long long acc = 0;
#pragma unroll
        for (int i = 1; i < 7; ++i) acc = (acc << 8) | me[i];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint32_t v = (uint32_t)(acc >> (42 - 6 * c)) & 0x3Fu;
            // msgs.rs:164-169 CHAR_CONVERT: 1-26 letters, 32 '_', 48-57 digits, else '#'
            const uint32_t ch = (v >= 1 && v <= 26) ? ('A' + v - 1) : (v == 32) ? '_' : (v >= 48 && v <= 57) ? v : '#';
            cs |= (unsigned long long)ch << (8 * c);
        }
    } else if (tc >= 9 && tc <= 18) {                      // msgs.rs:121-125, 69-102
        kind = 2;
        int a = (int)((((uint32_t)me[1] & 0xFEu) >> 1) << 4) | (int)(((uint32_t)me[2] & 0xF0u) >> 4);
        a = a * ((me[1] & 1u) ? 25 : 100) - 1000;
        alt = (uint32_t)a;
        lat = (((uint32_t)me[2] & 3u) << 15) | ((uint32_t)me[3] << 7) | (((uint32_t)me[4] & 0xFEu) >> 1);
        lon = (((uint32_t)me[4] & 1u) << 16) | ((uint32_t)me[5] << 8) | me[6];
        flags = ((me[0] & 6u) >> 1) | ((me[0] & 1u) << 8) | (((me[2] & 8u) >> 3) << 16) | (((me[2] & 4u) >> 2) << 24);
    }
    out[2 * k] = make_uint4(icao, df | (ca << 8) | (tc << 16) | (kind << 24), alt, lat);
    out[2 * k + 1] = make_uint4(lon, flags, (uint32_t)cs, (uint32_t)(cs >> 32));
}

__global__ void levels_u8_kernel(uint16_t *out, uint32_t minus_one)
{
    unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;   // idx = I | Q << 8
    if (idx >= 65536u) return;
    // the decode kernel's own arithmetic, both output positions and both streams must agree
    uint32_t o0, o1, p0, p1;
    levels_u8_streams(idx | (idx << 16), (idx ^ 0x5A5Au) * 0x10001u, minus_one, o0, o1);
    levels_u8_streams((idx ^ 0xA5A5u) * 0x10001u, idx | (idx << 16), minus_one, p0, p1);
    const uint32_t v = o0 & 0xFFFFu;
    const bool same = (o1 & 0xFFFFu) == v && (p0 >> 16) == v && (p1 >> 16) == v && v == (levels_u8_pair(idx, minus_one) & 0xFFFFu);
    out[idx] = same ? (uint16_t)v : (uint16_t)0xFFFFu;
}

__global__ void levels_cs16_kernel(const uint32_t *iq, unsigned long long n, uint16_t *out, uint32_t minus_one)
{
    unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    // the decode kernel's own arithmetic: both halves of a packed word must agree with the single-sample form
    const uint32_t w = iq[idx], other = iq[(idx * 7 + 3) % n];
    const uint32_t one = 0u - minus_one;
    const uint32_t v = level_word_cs16(w, other, minus_one, one), x = level_word_cs16(other, w, minus_one, one);
    const uint32_t lvl = 0xFFFFu + neg_isqrt_cs16(w, minus_one, one);
    out[idx] = ((v & 0xFFFFu) == lvl && (x >> 16) == lvl && lvl <= 0xFFFFu) ? (uint16_t)lvl : (uint16_t)0u;
}

}  // namespace

cudaError_t launch_decode(int format, const DecodeParams &params, cudaStream_t stream)
{
    if (params.n_tiles == 0) return cudaSuccess;
    DecodeParams p = params;
    const bool single = p.tiles_per_seg >= p.n_tiles;     // one segment: no per-tile division
    // complete tiles of a single-segment launch: tile * 2048 + 2288 <= n_samples
    constexpr unsigned long long kSpan = kStream + kTileWords;
    p.full_tiles = 0;
    if (single && p.n_samples >= kSpan)
        p.full_tiles = (unsigned)std::min<unsigned long long>(p.n_tiles, (p.n_samples - kSpan) / kWarpTile + 1);
    // tiles per warp: amortise the warp start-up on long captures, keep every SM busy on short buffers
    static const int forced = [] {
        const char *e = std::getenv("AIRGPU_TILES_PER_WARP");
        return e ? std::atoi(e) : 0;
    }();
    p.tiles_per_warp = forced > 0 ? (unsigned)forced : (p.n_tiles >= 131072u ? 4u : (p.n_tiles >= 32768u ? 2u : 1u));
    const unsigned per_cta = kWarps * p.tiles_per_warp;
    const unsigned grid = (p.n_tiles + per_cta - 1) / per_cta;
    static const int stage = [] {
        const char *e = std::getenv("AIRGPU_STAGE");
        return e ? std::atoi(e) : AIRGPU_STAGE_DEFAULT;
    }();
    if (format == AIRGPU_FMT_U8) {
        if (single && stage && p.vec_ok && p.full_tiles) decode_kernel_staged_u8<<<grid, kThreads, 0, stream>>>(p);
        else if (single) decode_kernel<AIRGPU_FMT_U8, true><<<grid, kThreads, 0, stream>>>(p);
        else decode_kernel<AIRGPU_FMT_U8, false><<<grid, kThreads, 0, stream>>>(p);
    } else {
        if (single) decode_kernel<AIRGPU_FMT_CS16, true><<<grid, kThreads, 0, stream>>>(p);
        else decode_kernel<AIRGPU_FMT_CS16, false><<<grid, kThreads, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t launch_finalize(const DecodeParams &p, const OutSet &dst, unsigned long long *d_total, cudaStream_t stream)
{
    const unsigned n_groups = (p.n_tiles + kGroupTiles - 1) / kGroupTiles;
    group_scan_kernel<<<1, kScanThreads, 0, stream>>>(p.group_sum, n_groups, p.group_base, d_total,
                                                       p.counters + kCounterGate, dst);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (n_groups == 0) return cudaSuccess;
    if (dst.multicast) gather_kernel<2><<<n_groups, kGroupTiles, 0, stream>>>(p, p.group_base, dst);
    else if (dst.n > 1) gather_kernel<1><<<n_groups, kGroupTiles, 0, stream>>>(p, p.group_base, dst);
    else gather_kernel<0><<<n_groups, kGroupTiles, 0, stream>>>(p, p.group_base, dst);
    return cudaGetLastError();
}

// One-kernel barrier between the ranks of an exchange.  Thread q publishes this rank's epoch in rank q's flag
// array (flags[q][rank], peer-mapped) with release semantics at system scope -- everything this GPU stored
// before (the records of the gather kernels earlier in the stream) is visible to a peer that acquires it -- and
// then waits until rank q's epoch shows up in its own array (flags[rank][q]).  Every rank runs on its own GPU
// (never two ranks of one exchange on one device: they would wait for each other's kernel).
// epoch == 0: the kernel counts for itself in flags[rank][n_ranks] (local memory), so that the same launch can be
// replayed from a CUDA graph.  Barriers that may run CONCURRENTLY on one GPU (two streams) need separate flag arrays.
// A peer that does not show up within ~4 s (a crashed rank) is not waited for any longer: the kernel records the
// failure in flags[rank][n_ranks + 1] and returns, so that nothing spins until a watchdog kills the job.
__global__ void peer_barrier_kernel(const PeerFlags f)
{
    __shared__ unsigned long long s_epoch;
    const unsigned q = threadIdx.x;
    if (q == 0) {
        unsigned long long e = f.epoch;
        if (e == 0ull) {
            unsigned long long *mine = f.flags[f.rank] + f.n_ranks;
            e = *mine + 1ull;
            *mine = e;
        }
        s_epoch = e;
    }
    __syncthreads();
    if (q >= f.n_ranks) return;
    const unsigned long long epoch = s_epoch;
    __threadfence_system();
    unsigned long long *theirs = f.flags[q] + f.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    const unsigned long long *mine = f.flags[f.rank] + q;
    unsigned long long seen, t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
        if (seen >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) {           // nanoseconds
            f.flags[f.rank][f.n_ranks + 1] = epoch;      // sticky: "the barrier of this epoch timed out"
            break;
        }
    }
}

cudaError_t launch_peer_barrier(const PeerFlags &f, cudaStream_t stream)
{
    peer_barrier_kernel<<<1, 32, 0, stream>>>(f);
    return cudaGetLastError();
}

cudaError_t launch_decode_fields(const airgpu_frame *frames, unsigned long long n, airgpu_fields *out, cudaStream_t stream)
{
    static_assert(sizeof(airgpu_fields) == 32, "airgpu_fields layout");
    if (n == 0) return cudaSuccess;
    fields_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const unsigned long long *>(frames), n,
                                                                    reinterpret_cast<uint4 *>(out));
    return cudaGetLastError();
}

cudaError_t launch_levels_u8(uint16_t *out65536, cudaStream_t stream)
{
    levels_u8_kernel<<<256, 256, 0, stream>>>(out65536, 0xFFFFFFFFu);
    return cudaGetLastError();
}

cudaError_t launch_levels_cs16(const int16_t *iq, unsigned long long n, uint16_t *out, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    levels_cs16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint32_t *>(iq), n, out, 0xFFFFFFFFu);
    return cudaGetLastError();
}

}  // namespace airgpu
