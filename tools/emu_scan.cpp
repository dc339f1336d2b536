// emu_scan.cpp -- host check of airgpu_scan.cuh (no GPU): the packed two-stream preamble gate and
// its hit-bit layout against a direct evaluation of demod.rs:17-44 on random level arrays, plus a
// bank-conflict count of the shared-memory access patterns.   g++ -O2 -std=c++17 tools/emu_scan.cpp
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <set>

#include "../air_rs_b200/csrc/airgpu_scan.cuh"

using namespace airgpu;

static bool gate_direct(const std::vector<uint32_t> &L, int i)
{
    static const int hs[4] = {0, 2, 7, 9};
    static const int ls[12] = {1, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15};
    uint32_t hi = 0, lo = 0xFFFFFFFFu;
    for (int h : hs) hi = L[i + h] > hi ? L[i + h] : hi;
    for (int l : ls) lo = L[i + l] < lo ? L[i + l] : lo;
    return hi <= lo;   // inverted levels: every high is at least every low
}

static int conflicts(const std::vector<int> &chunk_of_lane)   // 16-byte chunk index per lane, LDS/STS.128
{
    int worst = 0;
    for (int phase = 0; phase < 4; ++phase) {      // a 128-bit access is served a quarter warp at a time
        int cnt[8] = {0};
        for (int l = 8 * phase; l < 8 * phase + 8; ++l) cnt[chunk_of_lane[l] & 7]++;
        for (int b = 0; b < 8; ++b) worst = cnt[b] > worst ? cnt[b] : worst;
    }
    return worst;
}

template <bool kBf16>
static int check_gate()
{
    long long checked = 0, passes = 0;
    for (int trial = 0; trial < 400; ++trial) {
        srand(1234 + trial);
        const int range = 2 + trial % 7;           // few distinct levels: many passes and many ties
        std::vector<uint32_t> L(2 * kStream + 240 + 64);
        // bf16-comparable levels stay <= 0x7F00; the generic compare must work on all 16 bits
        const uint32_t scale = trial % 3 == 0 ? (kBf16 ? 2000u : 4095u) : 1u;   // 2 * 8 * scale fits the level range
        for (auto &v : L) v = 2u * (uint32_t)(rand() % range) * scale + ((!kBf16 && scale == 1u && trial % 5 == 0) ? 0xFFE0u : 0u);
        std::vector<uint32_t> W(kTileWords + 64);
        for (int w = 0; w < kTileWords; ++w) W[w] = L[w] | (L[kStream + w] << 16);
        for (int lane = 0; lane < 32; ++lane) {
            uint32_t R[48];
            for (int d = 0; d < 48; ++d) R[d] = W[kLaneX * lane + d];
            R[47] = 0xDEADBEEFu;                   // must not matter
            uint32_t hits[2];
            gate_scan<kBf16>(R, hits, 0xFFFFFFFFu);
            std::set<int> got;
            for (int h = 0; h < 2; ++h)
                for (int b = 0; b < 32; ++b)
                    if (hits[h] >> b & 1) got.insert(hit_stream(b) * kStream + kLaneX * lane + hit_x(h, b));
            for (int s = 0; s < 2; ++s)
                for (int x = 0; x < kLaneX; ++x) {
                    const int i = s * kStream + kLaneX * lane + x;
                    const bool want = gate_direct(L, i);
                    if (want != (got.count(i) != 0)) {
                        printf("MISMATCH trial %d lane %d stream %d x %d: want %d\n", trial, lane, s, x, (int)want);
                        return 1;
                    }
                    checked++;
                    passes += want;
                }
            if (got.size() > 64) { printf("too many bits\n"); return 1; }
        }
    }
    printf("gate_scan<%s>: %lld offsets checked, %lld passes, all equal\n", kBf16 ? "bf16" : "u16", checked, passes);
    return 0;
}

int main()
{
    if (check_gate<true>() || check_gate<false>()) return 1;

    // every (h, b) maps to a distinct (stream, x)
    std::set<int> seen;
    for (int h = 0; h < 2; ++h)
        for (int b = 0; b < 32; ++b) seen.insert(hit_stream(b) * 64 + hit_x(h, b));
    printf("hit-bit map: %zu distinct of 64\n", seen.size());

    // bank conflicts
    int worst = 0;
    for (int m = 0; m < 5; ++m)
        for (int half = 0; half < 2; ++half) {
            std::vector<int> c(32);
            for (int l = 0; l < 32; ++l) c[l] = phys_chunk4(2 * (l + 32 * m) + half);
            const int w = conflicts(c);
            worst = w > worst ? w : worst;
        }
    printf("phase-1 stores: worst %d-way\n", worst);
    worst = 0;
    for (int k = 0; k < 12; ++k) {
        std::vector<int> c(32);
        for (int l = 0; l < 32; ++l) c[l] = phys_chunk4(8 * l + k);
        const int w = conflicts(c);
        worst = w > worst ? w : worst;
    }
    printf("phase-2 loads: worst %d-way\n", worst);
    // phys_word agrees with phys_chunk4
    for (int w = 0; w < kTileWords; ++w)
        if (phys_word(w) != 4 * phys_chunk4(w >> 2) + (w & 3)) { printf("phys mismatch at %d\n", w); return 1; }
    printf("phys_word == phys_chunk4 layout, padded words %d (max used %d)\n", kTileWordsPadded, phys_word(kTileWords - 1));

    // the scalar readers: DF test and slicer addressing against the level they are meant to read
    {
        srand(99);
        std::vector<uint32_t> L(2 * kStream + 240);
        for (auto &v : L) v = (uint32_t)(rand() & 0xFFFF);
        std::vector<uint16_t> S(2 * kTileWordsPadded, 0xDEAD);      // pads keep the sentinel
        for (int w = 0; w < kTileWords; ++w) {
            S[2 * phys_word(w)] = (uint16_t)L[w];
            S[2 * phys_word(w) + 1] = (uint16_t)L[kStream + w];
        }
        int max_index = 0;
        long long reads = 0;
        for (int i = 0; i < 2 * kStream; ++i) {
            const int xw = i & (kStream - 1), st = i >> 10;
            const int base = level_index(xw, st, 16), cross = df_cross(xw);
            for (int k = 0; k < 10; ++k) {
                const int idx = base + 2 * k + (k >= cross ? 8 : 0);
                if (S[idx] != (uint16_t)L[i + 16 + k]) { printf("DF address wrong: i %d k %d\n", i, k); return 1; }
                max_index = idx > max_index ? idx : max_index;
                reads++;
            }
            for (int lane = 0; lane < 32; ++lane) {
                const int wj = slicer_word(xw, lane);
                const int i0 = 2 * phys_word(wj) + st, i1 = i0 + slicer_step(wj);
                for (int r = 0; r < 4; ++r) {
                    const int k = 31 - lane + 32 * r;
                    if (k >= 112) continue;                          // round 3 is predicated to lanes 16..31
                    const int a = i0 + kSlicerRoundStride * r, b = i1 + kSlicerRoundStride * r;
                    if (S[a] != (uint16_t)L[i + 16 + 2 * k] || S[b] != (uint16_t)L[i + 17 + 2 * k]) {
                        printf("slicer address wrong: i %d bit %d\n", i, k);
                        return 1;
                    }
                    max_index = b > max_index ? b : max_index;
                    reads += 2;
                }
            }
        }
        printf("scalar readers: %lld level reads at the right address, max u16 index %d < %d\n", reads, max_index,
               2 * kTileWordsPadded);
        if (max_index >= 2 * kTileWordsPadded) return 1;
    }
    return seen.size() == 64 ? 0 : 1;
}
