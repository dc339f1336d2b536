// airgpu_kernels.cuh -- device-side interface of the ADS-B decode stage (sm_100a).
//
// One fused kernel per capture does the whole reference path
//   IQ -> magnitude -> preamble/DF gate -> bit slice -> CRC-24 (+1-bit repair) -> frame list
// (reference src/adsb.rs:92-122 and what it calls); two small kernels then put
// the per-tile frame lists into the reference's emission order.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/airgpu.h"

namespace airgpu {

// ---- tiling ---------------------------------------------------------------
// A warp is the unit of work: it owns kWarpTile consecutive candidate offsets and a
// private slice of shared memory, so the hot loop has no CTA-wide barrier at all.
constexpr int kWarpTile = 2048;                   // candidate offsets per warp: two streams of 1024 (airgpu_scan.cuh)
constexpr int kThreads = 128;                     // 4 independent warps per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kTile = kWarpTile;                  // ordering unit: one tile_tab entry per warp tile
constexpr int kSlotsPerTile = 8;                  // fixed scratch slots per tile; more frames than that go to the overflow area
constexpr int kSlotBytes = 32;                    // 4 big-endian frame words | offset in tile | fixed_bit << 16 | pad
constexpr int kFrameSamples = 240;                // 16 + 112 * 2, reference src/adsb.rs:98
constexpr int kGroupTiles = 256;                  // tiles per ordering group (one gather CTA)

struct DecodeParams {
    const void *iq;                 // interleaved IQ, device memory
    unsigned long long n_samples;   // complex samples in the capture
    unsigned long long seg_len;     // samples per independent segment (>= 1)
    unsigned int tiles_per_seg;
    unsigned int n_tiles;           // n_segments * tiles_per_seg
    unsigned int minus_one;         // 0xFFFFFFFF (see levels_u8_pair)
    unsigned int vec_ok;            // every warp slice starts 16-byte aligned (base aligned, seg_len % 8 == 0)
    unsigned int full_tiles;        // set by launch_decode: leading tiles that are complete (single segment only)
    unsigned int tiles_per_warp;    // set by launch_decode: consecutive rounds of tiles one warp decodes
    unsigned int force_ordered;     // tests: every tile takes the ordered (overflow) path
    unsigned long long base_offset; // added to every frame offset
    void *scratch;                  // n_tiles * kSlotsPerTile fixed slots of kSlotBytes, then ovf_cap overflow slots
    unsigned long long cap;         // capacity of the final output
    unsigned long long ovf_cap;     // capacity of the overflow area
    unsigned long long *counters;   // see the enum below
    unsigned long long *ovf_counter; // next free record of the overflow area (zeroed before the launch)
    uint2 *tile_tab;                // per tile: (overflow base, frame count)
    unsigned long long *group_sum;  // per group of kGroupTiles tiles: gate passes << 32 | frames (zeroed before the launch)
    unsigned long long *group_base; // per group: ordered position of its first frame
};

enum { kCounterGate = 1, kNumCounters = 4 };

// Where the ordered records and the frame count of a call go: one local array, or the same offset of several
// local / peer-mapped arrays, or one NVSwitch multicast address (airgpu_decode_device_peers).
struct OutSet {
    unsigned long long *out[AIRGPU_MAX_PEERS];     // airgpu_frame arrays, u64 view
    unsigned long long *count[AIRGPU_MAX_PEERS];   // where the count lands (may be nullptr)
    unsigned n;
    unsigned multicast;
};

struct PeerFlags {
    unsigned long long *flags[AIRGPU_MAX_PEERS];   // flags[q]: rank q's array of n_ranks epochs, as mapped on this device
    unsigned n_ranks, rank;
    unsigned long long epoch;
};

// Launchers (stream-ordered, no synchronisation inside).
cudaError_t launch_decode(int format, const DecodeParams &p, cudaStream_t stream);
// Ordered frames are appended to `out` at index *d_total (device counter, updated in place).
cudaError_t launch_finalize(const DecodeParams &p, const OutSet &dst, unsigned long long *d_total, cudaStream_t stream);
cudaError_t launch_peer_barrier(const PeerFlags &f, cudaStream_t stream);

// N1: per-frame field decode (packet.rs:25-49, msgs.rs:69-102, 171-201).
cudaError_t launch_decode_fields(const airgpu_frame *frames, unsigned long long n, airgpu_fields *out, cudaStream_t stream);

// Exhaustive self-check helper used by the tests: level (inverted magnitude proxy)
// the kernel computes for every U8 (I, Q) pair / for a list of CS16 samples.
cudaError_t launch_levels_u8(uint16_t *out65536, cudaStream_t stream);
cudaError_t launch_levels_cs16(const int16_t *iq, unsigned long long n, uint16_t *out, cudaStream_t stream);

}  // namespace airgpu
