/*
 * adsb_oracle.h -- CPU restatement of air_rs's ADS-B decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * --impl reference legs may build, load or call it.  The CUDA library
 * (libairgpu.so) never links or calls into this file.
 *
 * What it restates (all file:line relative to the upstream reference tree):
 *   src/utils.rs:46-52        get_magnitude
 *   src/adsb.rs:92-122        process_sdr_data_thread (loop bounds, order, no skip)
 *   src/adsb/demod.rs:17-57   check_for_adsb_packet
 *   src/adsb/demod.rs:65-82   extract_packet
 *   src/adsb/demod.rs:92-131  extract_manchester_relative
 *   src/adsb/demod.rs:180-201 decode_packet
 *   src/adsb/crc.rs:10-40     get_adsb_crc
 *   src/adsb/crc.rs:49-65     try_crc_recovery
 *
 * Pinning status: the CRC, gate and bad-CRC known-answer tests the reference
 * carries (demod.rs:250-278, 337-380) and its seven CRC-valid DF17 frames
 * (aircraft.rs:188-261, demod.rs:339-344) are checked in tests/test_oracle.py.
 * The reference ships NO IQ fixture and cannot be compiled here (no Rust
 * toolchain), so the end-to-end IQ -> frames behaviour is "parity unpinned":
 * it rests on this literal restatement agreeing with an independent numpy
 * restatement (oracle/oracle_np.py) on every test input.
 */
#ifndef ADSB_ORACLE_H
#define ADSB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 24-byte record the GPU stage emits (include/airgpu.h: airgpu_frame). */
typedef struct {
    uint8_t  bytes[14];   /* the Vec<u8> handed to AdsbPacket::new               */
    uint8_t  fixed_bit;   /* 0xFF: CRC matched; else data bit 0..87 that was flipped */
    uint8_t  reserved;    /* 0 */
    uint64_t offset;      /* base + i, i = preamble start (adsb.rs:98 loop index)  */
} oracle_frame;

#define ORACLE_FMT_CS16 0
#define ORACLE_FMT_U8   1

#define ORACLE_PREAMBLE_SAMPLES 16
#define ORACLE_FRAME_SAMPLES    240   /* 16 + 112*2, adsb.rs:98 */

/* ---- literal pieces (one C function per reference function) ------------- */

/* utils.rs:46-52: ((re as f64)^2 + (im as f64)^2).sqrt() as u32 */
void oracle_get_magnitude(const int16_t *iq, size_t n, uint32_t *mags);

/* SURVEY 8(d) definition of the U8 input mode: re = (2u-255)*128 */
void oracle_widen_u8(const uint8_t *iq, size_t n, int16_t *out);

/* demod.rs:17-57. Returns 1 and writes *high when the gate passes, else 0. */
int oracle_check_for_adsb_packet(const uint32_t buf[32], uint32_t *high);

/* demod.rs:92-131 (len must be a multiple of 16). Returns 1 (Some) / 0 (None). */
int oracle_extract_manchester_relative(const uint32_t *buf, size_t len, uint32_t high,
                                       uint16_t *symbols);

/* demod.rs:180-201. Always "Some" in the reference; returns number of bytes. */
size_t oracle_decode_packet(const uint16_t *symbols, size_t n, uint8_t *bytes);

/* crc.rs:10-40 */
uint32_t oracle_get_adsb_crc(const uint8_t *buf, size_t len);

/* crc.rs:49-65. buf has len bytes (14), flipped in place on success.
 * Returns 1 (Some) and the MSB-first bit position in *flipped, else 0. */
int oracle_try_crc_recovery(uint8_t *buf, size_t len, uint32_t calc_crc,
                            uint32_t packet_crc, int *flipped);

/* demod.rs:65-82. buf = 224 magnitudes. Returns 1 (Some) / 0 (None). */
int oracle_extract_packet(const uint32_t *buf, size_t len, uint32_t high,
                          uint8_t out[14], int *fixed_bit);

/* adsb.rs:96-116 for ONE buffer of magnitudes: every offset i in
 * [0, len-240) ascending; returns the number of frames the loop emits
 * (all of them are counted; at most cap are stored).  gate_passes, if not
 * NULL, receives the reference's num_processed counter (adsb.rs:105).
 * len < 240 panics in the reference; here it yields zero frames. */
size_t oracle_process_mags(const uint32_t *mags, size_t len, uint64_t base,
                           oracle_frame *out, size_t cap, uint64_t *gate_passes);

/* ---- whole-path entry points -------------------------------------------- */

/* Literal path: widen (U8) -> get_magnitude -> process_mags, for a capture cut
 * into independent segments of segment_samples (0 = one segment = CONTINUOUS).
 * Frame offsets are base + segment_start + i.  Single-threaded.  */
size_t oracle_decode_literal(const void *iq, size_t n_samples, int format,
                             size_t segment_samples, uint64_t base,
                             oracle_frame *out, size_t cap, uint64_t *gate_passes);

/* Fast path: integer isqrt into u16, branch-light gate, table CRC, 88-entry
 * syndrome lookup; n_threads > 1 splits each segment into contiguous candidate
 * ranges with a 239-sample overlap.  Must equal oracle_decode_literal on every
 * input (tests/test_oracle.py proves it on the test corpus). */
size_t oracle_decode_fast(const void *iq, size_t n_samples, int format,
                          size_t segment_samples, uint64_t base,
                          oracle_frame *out, size_t cap, uint64_t *gate_passes,
                          int n_threads);

/* Literal path run chunk-parallel (same overlap scheme as the fast path): the
 * reference's own arithmetic on all host cores, for bench.py --impl reference. */
size_t oracle_decode_literal_mt(const void *iq, size_t n_samples, int format,
                                size_t segment_samples, uint64_t base,
                                oracle_frame *out, size_t cap, uint64_t *gate_passes,
                                int n_threads);

/* ---- next row N1: what AdsbPacket::new derives from the 14 bytes ------------------ *
 * src/adsb/packet.rs:25-49, src/adsb/msgs.rs:69-102 (AircraftPosition::new),
 * :141-162 (to_6bit_chunks), :164-187 (CHAR_CONVERT, AircraftID::new).  Same 32-byte layout as
 * airgpu_fields.  Pinned by the reference's unit tests msgs.rs:225-321 (tests/test_fields.py). */
typedef struct {
    uint32_t icao;
    uint8_t  downlink_format, capability, msg_type, kind;   /* kind: 0 Uknown, 1 AircraftID, 2 AircraftPosition */
    int32_t  altitude;
    uint32_t cpr_latitude, cpr_longitude;
    uint8_t  surveillance_status, nic_supplement, cpr_time, cpr_odd;
    char     callsign[8];
} oracle_fields;

void oracle_packet_fields(const uint8_t packet[14], oracle_fields *out);
void oracle_frames_fields(const oracle_frame *frames, size_t n, oracle_fields *out);

/* 88 single-bit syndromes T[p] = crc(e_p) used by the fast path (for tests). */
void oracle_syndrome_table(uint32_t table[88]);

#ifdef __cplusplus
}
#endif
#endif
