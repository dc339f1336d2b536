/*
 * airgpu.h -- C ABI of the B200 ADS-B / Mode S decode stage.
 *
 * This library is a drop-in for ONE thread of jaxsonpd/air_rs: the body of
 *     fn process_sdr_data_thread(rx: Receiver<Vec<Complex<i16>>>, tx: Sender<AdsbPacket>)
 * (reference src/adsb.rs:92-122, spawned at src/adsb.rs:147), i.e.
 *     get_magnitude            src/utils.rs:46-52
 *     check_for_adsb_packet    src/adsb/demod.rs:17-57
 *     extract_packet           src/adsb/demod.rs:65-82  (slicer :92-131, :180-201)
 *     get_adsb_crc             src/adsb/crc.rs:10-40
 *     try_crc_recovery         src/adsb/crc.rs:49-65
 * Everything before it (SDR / playback threads) and after it
 * (AdsbPacket::new, aircraft tracking, TUI / web) stays in the Rust host.
 *
 * Conventions
 *   - plain C types only; no exceptions, no unwinding across the boundary;
 *   - every call returns AIRGPU_OK (0) or a negative airgpu_status;
 *     airgpu_last_error() gives the message for the calling thread;
 *   - there is NO CPU fallback: without a CUDA device airgpu_create fails;
 *   - a context is used by one thread at a time (the reference has exactly one
 *     decode thread, src/adsb.rs:147).
 *
 * Semantics (bit-exact with the reference's CPU decode)
 *   A capture of n_samples complex samples is cut into independent segments of
 *   segment_samples (0 = the whole capture is one segment).  Within a segment
 *   of length L every offset i in [0, L-240) is tested in ascending order
 *   (src/adsb.rs:98); a frame is emitted for every i that passes the gate and
 *   whose CRC matches or is repairable -- there is NO skip after a hit and NO
 *   de-duplication (src/adsb.rs:113 is a no-op).  L < 240, which panics in the
 *   reference, yields zero frames here (the one deliberate deviation).
 *   segment_samples = 20000 reproduces the reference's playback CHUNKING
 *   (src/adsb.rs:78); one segment per received buffer reproduces the SDR path.
 *   To reproduce a whole-file REPLAY pass only the samples the playback thread
 *   sends: it loops `while i < data.len() - 20000` (src/adsb.rs:77), i.e. it
 *   never sends the last chunk, complete or not -- n_samples =
 *   ((len - 1) / 20000) * 20000 (airgpu_playback_samples() below).
 */
#ifndef AIRGPU_H
#define AIRGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIRGPU_ABI_VERSION 2
#define AIRGPU_MAX_PEERS 8       /* destinations of a fused frame exchange: the GPUs of one NVSwitch box */

typedef enum airgpu_status {
    AIRGPU_OK = 0,
    AIRGPU_ERR_INVALID = -1,    /* bad argument                                   */
    AIRGPU_ERR_NO_DEVICE = -2,  /* no CUDA device / wrong architecture            */
    AIRGPU_ERR_CUDA = -3,       /* a CUDA call failed; see airgpu_last_error()    */
    AIRGPU_ERR_NOMEM = -4,      /* host or device allocation failed               */
    AIRGPU_ERR_OVERFLOW = -5,   /* more frames than the output capacity           */
    AIRGPU_ERR_BUSY = -6,       /* ring full: collect before submitting more      */
    AIRGPU_ERR_TICKET = -7      /* unknown or already collected ticket            */
} airgpu_status;

/* Sample formats.  CS16 is what the reference receives (Vec<Complex<i16>>,
 * src/adsb.rs:54-59: interleaved little-endian i16 I,Q -- 4 bytes/sample).
 * U8 is the RTL-SDR native format BASELINE.json names (interleaved unsigned
 * I,Q bytes -- 2 bytes/sample); it is DEFINED as CS16 with
 * re = (2*u - 255) * 128 and decoded with identical results. */
typedef enum airgpu_format {
    AIRGPU_FMT_CS16 = 0,
    AIRGPU_FMT_U8 = 1
} airgpu_format;

/* One decoded frame.  `bytes` is exactly the Vec<u8> the reference passes to
 * AdsbPacket::new (src/adsb.rs:107, src/adsb/packet.rs:25).  `offset` and
 * `fixed_bit` do not exist in the reference's output; they are derived from its
 * control flow (loop index at src/adsb.rs:98; flipped bit at src/adsb/crc.rs:52-55). */
typedef struct airgpu_frame {
    uint8_t  bytes[14];
    uint8_t  fixed_bit;   /* 0xFF = CRC matched as received; else repaired data bit 0..87 (MSB first) */
    uint8_t  reserved;    /* 0 */
    uint64_t offset;      /* base_offset + segment start + i (preamble start, in samples) */
} airgpu_frame;           /* 24 bytes */

typedef struct airgpu_config {
    uint32_t struct_size;        /* sizeof(airgpu_config), for ABI evolution            */
    int32_t  device;             /* CUDA device ordinal                                  */
    uint32_t format;             /* airgpu_format                                        */
    uint32_t ring_slots;         /* streaming ring depth (0 -> 4)                        */
    uint64_t max_buffer_samples; /* largest buffer airgpu_submit will see (0 -> 262144)  */
    uint64_t max_frames;         /* output capacity per submit / per decode call (0 -> auto) */
} airgpu_config;

typedef struct airgpu_ctx airgpu_ctx;

/* Counters of the last completed decode call on this context. */
typedef struct airgpu_stats {
    uint64_t n_samples;
    uint64_t n_frames;       /* frames produced (may exceed the capacity on overflow)   */
    uint64_t gate_passes;    /* the reference's num_processed counter, src/adsb.rs:105  */
    uint64_t n_tiles;
    float    kernel_ms;      /* device time of the decode kernels (CUDA events), 0 if not measured */
    float    h2d_ms;         /* host->device copy time of the last host-buffer decode   */
    float    decode_ms;      /* device time of the fused decode kernel alone: average over its launches since the
                                previous airgpu_get_stats (the newest 64 of them), CUDA events on the launch stream */
    float    decode_launches; /* how many launches that average covers                */
} airgpu_stats;

const char *airgpu_version(void);
const char *airgpu_last_error(void);
int airgpu_device_count(void);

/* Replaces the setup the reference does with expect() at src/adsb.rs:35-48 /
 * :126-147: returns an error code instead of panicking. */
int airgpu_create(const airgpu_config *cfg, airgpu_ctx **out);
void airgpu_destroy(airgpu_ctx *ctx);

/* ---- streaming: the loop body of process_sdr_data_thread ----------------- *
 * submit  == `while let Ok(buf) = rx.recv()` (src/adsb.rs:95): copies the buffer
 *            into a pinned ring slot, queues H2D on the copy stream and the
 *            decode on the compute stream, and returns at once.
 * collect == the frames that buffer yields, ascending offset, in the order the
 *            reference would tx.send() them (src/adsb.rs:107-111).  Tickets
 *            must be collected in submission order.  Each buffer is one
 *            independent segment (no state across buffers, as in the reference). */
int airgpu_submit(airgpu_ctx *ctx, const void *iq, size_t n_samples,
                  uint64_t base_offset, uint64_t *ticket);
int airgpu_collect(airgpu_ctx *ctx, uint64_t ticket, airgpu_frame *out, size_t cap,
                   size_t *n_frames);
/* The ring keeps room for the worst case on the device (a constant buffer yields a
 * frame at every offset: max_buffer_samples - 240 records), so nothing is ever dropped
 * there.  If `cap` is smaller than the buffer's frame count airgpu_collect returns
 * AIRGPU_ERR_OVERFLOW with *n_frames = that count and KEEPS the ticket: collect it
 * again with a larger array.  Only the records that exist cross PCIe (a small fixed
 * head with the count, the rest on demand), not max_frames of them. */

/* Samples of a capture of `len` samples that the reference's playback thread actually
 * sends in chunks of `chunk` (src/adsb.rs:75-89: the last chunk is never sent). */
size_t airgpu_playback_samples(size_t len, size_t chunk);

/* ---- one-shot decode of a capture held in HOST memory -------------------- *
 * Chunks the capture through the pinned ring (H2D overlapped with compute,
 * 240-sample overlap between chunks inside a segment) and returns the frames
 * in reference order.  segment_samples as described above; n_streams batches
 * of equal length are just segment_samples = samples per stream. */
int airgpu_decode(airgpu_ctx *ctx, const void *iq, size_t n_samples,
                  size_t segment_samples, uint64_t base_offset,
                  airgpu_frame *out, size_t cap, size_t *n_frames);

/* ---- decode of a capture already resident in DEVICE memory --------------- *
 * d_iq and d_out are device pointers on the context's device; d_out holds
 * `cap` records.  `stream` is a cudaStream_t (NULL = the context's compute
 * stream).  Asynchronous: the frame count lands in *d_count (device, 8 bytes);
 * with d_count == NULL it is kept by the library and airgpu_sync_count() waits
 * for the decode and returns it.
 * Shards of a long capture call this once per GPU with base_offset = first
 * sample of the shard and n_samples including the 240-sample right halo. */
int airgpu_decode_device(airgpu_ctx *ctx, const void *d_iq, size_t n_samples,
                         size_t segment_samples, uint64_t base_offset,
                         airgpu_frame *d_out, size_t cap, uint64_t *d_count,
                         void *stream);
int airgpu_sync_count(airgpu_ctx *ctx, uint64_t *n_frames);

int airgpu_get_stats(airgpu_ctx *ctx, airgpu_stats *out);

/* Pre-size the device workspace for captures of up to n_samples (cut into segments of
 * segment_samples, 0 = one) and outputs of up to `cap` frames: later calls within these
 * bounds neither allocate nor synchronise the device (a long-running streaming host
 * calls this once after airgpu_create). */
int airgpu_reserve(airgpu_ctx *ctx, size_t n_samples, size_t segment_samples, size_t cap);

/* CUDA-event timing of the decode kernels (airgpu_stats.kernel_ms / decode_ms) on or off.
 * Default on; off inside airgpu_graph_begin .. airgpu_graph_end automatically. */
int airgpu_set_timing(airgpu_ctx *ctx, int enabled);

/* ---- CUDA graphs: record a sequence of device-side calls once, replay it with one launch ---- *
 * Between begin and end every airgpu_decode_device / airgpu_decode_device_peers /
 * airgpu_decode_fields / airgpu_peer_barrier call on `stream` is captured instead of
 * executed (cudaStreamBeginCapture); the workspace must already be large enough
 * (airgpu_reserve or one eager call first), otherwise the call fails with
 * AIRGPU_ERR_INVALID instead of allocating.  A launch replays the recorded work on the
 * same buffers: a 20 000-sample buffer costs one launch instead of five. */
typedef struct airgpu_graph airgpu_graph;
int airgpu_graph_begin(airgpu_ctx *ctx, void *stream);
int airgpu_graph_end(airgpu_ctx *ctx, void *stream, airgpu_graph **out);
int airgpu_graph_launch(airgpu_graph *graph, void *stream);
/* A capture may span several streams (forked from the capturing one with events) and several contexts: tell every
 * OTHER context whose calls land in the capture, so that it neither records timing events nor allocates. */
int airgpu_set_capturing(airgpu_ctx *ctx, int capturing);
void airgpu_graph_destroy(airgpu_graph *graph);

/* ---- multi-GPU: the frame exchange fused into the ordering kernels ---------------- *
 * A long capture shards into contiguous candidate ranges (+240-sample halo), one per
 * GPU; the only exchange is the per-GPU ordered frame lists.  Instead of
 * decoding into a local array and copying it to the peers afterwards, the kernels that
 * put the records in order store each record -- and the frame count -- straight to
 * every destination: device-accessible addresses in this GPU's own memory, in peer
 * GPUs' memory (cudaDeviceEnablePeerAccess / CUDA IPC / symmetric memory), or ONE
 * NVSwitch multicast address (`multicast` = 1: multimem.st, the switch replicates).
 * All destinations receive the same `cap`-record array layout. */
typedef struct airgpu_peers {
    uint32_t      struct_size;              /* sizeof(airgpu_peers)                         */
    uint32_t      n_outs;                   /* 1 .. AIRGPU_MAX_PEERS                        */
    uint32_t      multicast;                /* 1: outs[0] / counts[0] are multicast addresses */
    uint32_t      reserved;
    airgpu_frame *outs[AIRGPU_MAX_PEERS];
    uint64_t     *counts[AIRGPU_MAX_PEERS]; /* may be NULL: no count written there         */
} airgpu_peers;
int airgpu_decode_device_peers(airgpu_ctx *ctx, const void *d_iq, size_t n_samples,
                               size_t segment_samples, uint64_t base_offset,
                               const airgpu_peers *dst, size_t cap, void *stream);
/* One-kernel barrier across the ranks of an exchange: flags[q] is rank q's array of
 * n_ranks epochs (zero-initialised) as mapped in this process; every rank calls it with
 * the same, increasing `epoch` (>= 1).  After it, the records every rank stored before its
 * own call are visible here.  epoch = 0: the kernel keeps the count itself in
 * flags[rank][n_ranks], so that one recorded launch can be replayed from a CUDA graph.
 * The arrays hold n_ranks + 2 words: a peer that does not arrive within ~4 s (a crashed
 * rank) is given up on and the epoch is recorded in flags[rank][n_ranks + 1] (non-zero =
 * a barrier timed out; the caller checks it).  Barriers that may run concurrently on one
 * GPU (two streams) need separate flag arrays.  One rank per GPU: two ranks of one
 * exchange on the same device would wait for each other's kernel. */
int airgpu_peer_barrier(airgpu_ctx *ctx, uint64_t *const *flags, uint32_t n_ranks, uint32_t rank,
                        uint64_t epoch, void *stream);

/* ---- multi-GPU from ONE host thread (what the Rust decode thread can call) --------- *
 * A group owns one context per listed device (a device may be listed more than once).
 * airgpu_group_decode shards a capture held in HOST memory over them (contiguous
 * candidate ranges + 240-sample halo), runs every shard's H2D copies and kernels
 * concurrently, and returns the concatenated frame list: rank order == ascending
 * offset == the reference's order, identical to airgpu_decode on one GPU. */
typedef struct airgpu_group airgpu_group;
int airgpu_group_create(const int *devices, uint32_t n_devices, uint32_t format, airgpu_group **out);
int airgpu_group_decode(airgpu_group *grp, const void *iq, size_t n_samples, uint64_t base_offset,
                        airgpu_frame *out, size_t cap, size_t *n_frames);
/* per-shard counters of the last airgpu_group_decode (stats[k] for devices[k]) */
int airgpu_group_stats(airgpu_group *grp, airgpu_stats *stats, uint32_t n_stats);
void airgpu_group_destroy(airgpu_group *grp);
/* Convenience: create a group, decode once, destroy it (SURVEY 8(b) "airgpu_decode_sharded"). */
int airgpu_decode_sharded(const int *devices, uint32_t n_devices, uint32_t format, const void *iq,
                          size_t n_samples, uint64_t base_offset, airgpu_frame *out, size_t cap,
                          size_t *n_frames);

/* ---- next row N1: frame field decode on the device ---------------------------- *
 * What AdsbPacket::new derives from the 14 bytes (src/adsb/packet.rs:25-49) and the
 * two message decoders it calls (src/adsb/msgs.rs:69-102 AircraftPosition::new,
 * :171-201 AircraftID::new), one record per frame, so that an exchange of decoded
 * records can replace an exchange of raw frames.  Quirks kept: capability = b0 & 5. */
typedef struct airgpu_fields {
    uint32_t icao;                 /* packet.rs:28                                        */
    uint8_t  downlink_format;      /* b0 >> 3                                             */
    uint8_t  capability;           /* b0 & 5 (sic, packet.rs:27)                          */
    uint8_t  msg_type;             /* b4 >> 3 (type code)                                 */
    uint8_t  kind;                 /* 0 Uknown, 1 AircraftID (TC 1-4), 2 AircraftPosition (TC 9-18) */
    int32_t  altitude;             /* feet, msgs.rs:71-75 (position only, else 0)         */
    uint32_t cpr_latitude;         /* 17 bits, msgs.rs:84-86                              */
    uint32_t cpr_longitude;        /* 17 bits, msgs.rs:87-89                              */
    uint8_t  surveillance_status;  /* msgs.rs:78                                          */
    uint8_t  nic_supplement;       /* msgs.rs:79                                          */
    uint8_t  cpr_time;             /* msgs.rs:80                                          */
    uint8_t  cpr_odd;              /* msgs.rs:81-82: 1 = CprFormat::Odd                   */
    char     callsign[8];          /* msgs.rs:164-187, not NUL terminated (ID only, else zeros) */
} airgpu_fields;                   /* 32 bytes */

/* d_frames / d_out are device pointers; stream is a cudaStream_t (NULL = compute stream). */
int airgpu_decode_fields(airgpu_ctx *ctx, const airgpu_frame *d_frames, size_t n_frames,
                         airgpu_fields *d_out, void *stream);
/* Convenience for host arrays (copies in, decodes, copies out, synchronises). */
int airgpu_decode_fields_host(airgpu_ctx *ctx, const airgpu_frame *frames, size_t n_frames,
                              airgpu_fields *out);

/* Page-locked host buffers ("pinned host ring buffers"): a receive thread that
 * fills buffers from airgpu_host_alloc lets airgpu_decode copy asynchronously
 * straight from them.  Replaces the plain Vec allocations at src/adsb.rs:60,64,78. */
int airgpu_host_alloc(size_t bytes, void **out);
int airgpu_host_free(void *p);

/* ---- diagnostics (device arithmetic only; used by the parity tests) ------- *
 * The per-sample "level" the kernel compares instead of the magnitude:
 *   U8  : out[I | Q << 8] for all 65536 byte pairs;
 *   CS16: out[k] = 65535 - floor(sqrt(re_k^2 + im_k^2)). */
int airgpu_dbg_levels_u8(airgpu_ctx *ctx, uint16_t *out65536);
int airgpu_dbg_levels_cs16(airgpu_ctx *ctx, const int16_t *iq, size_t n_samples, uint16_t *out);

#ifdef __cplusplus
}
#endif
#endif /* AIRGPU_H */
