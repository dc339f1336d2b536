// airgpu_api.cu -- the C ABI declared in include/airgpu.h.
//
// Host-side plumbing only: context, device workspace, the pinned ring used by
// airgpu_submit/collect and the chunk pipeline of airgpu_decode (H2D on a copy
// stream overlapped with the decode kernels on a compute stream).  All decode
// arithmetic lives in airgpu_kernels.cu; there is no CPU fallback anywhere.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "airgpu_kernels.cuh"

using namespace airgpu;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? AIRGPU_ERR_NOMEM : AIRGPU_ERR_CUDA,     \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline size_t bytes_per_sample(int fmt) { return fmt == AIRGPU_FMT_U8 ? 2 : 4; }

struct Slot {
    void *h_in = nullptr;          // pinned
    void *d_in = nullptr;
    airgpu_frame *h_out = nullptr; // pinned
    unsigned long long *h_count = nullptr;   // pinned
    cudaEvent_t copied = nullptr, done = nullptr;
    uint64_t ticket = 0;
    bool busy = false;
};

struct Geometry {
    unsigned long long seg_len = 0;
    unsigned tiles_per_seg = 0;
    unsigned n_tiles = 0;
};

// tiles for a capture of n samples cut into segments of seg (0 = one segment)
bool make_geometry(size_t n, size_t seg, Geometry &g)
{
    if (seg == 0 || seg > n) seg = n;
    g.seg_len = seg ? seg : 1;
    if (n == 0 || seg <= (size_t)kFrameSamples) {
        g.tiles_per_seg = 0;
        g.n_tiles = 0;
        return true;
    }
    unsigned long long n_seg = (n + seg - 1) / seg;
    unsigned long long tps = (seg - kFrameSamples + kTile - 1) / kTile;
    unsigned long long nt = n_seg * tps;
    if (nt > 0x7FFFFFFFull) return false;
    g.tiles_per_seg = (unsigned)tps;
    g.n_tiles = (unsigned)nt;
    return true;
}

}  // namespace

struct airgpu_ctx {
    int device = 0;
    int format = AIRGPU_FMT_CS16;
    size_t max_buffer_samples = 0;
    size_t max_frames = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evh0 = nullptr, evh1 = nullptr, ev_sync = nullptr, evk0 = nullptr, evk1 = nullptr;
    bool ev_valid = false, evh_valid = false, evk_valid = false, sync_valid = false;

    // workspace shared by every decode on the compute stream (stream-ordered reuse)
    void *scratch = nullptr;                   // kSlotBytes per slot
    size_t scratch_cap = 0;                    // slots
    uint2 *tile_tab = nullptr;
    unsigned long long *group_sum = nullptr;   // [overflow counter | per-group sums | bases], 1 + 2 * groups_cap
    size_t groups_cap = 0;
    size_t tiles_cap = 0;
    unsigned long long *counters = nullptr;   // kNumCounters + 1 (last = running frame total)
    unsigned long long *h_counters = nullptr; // pinned mirror
    airgpu_frame *out_dev = nullptr;           // device output for host-facing calls
    size_t out_cap = 0;

    // chunk pipeline for airgpu_decode
    void *chunk_dev[2] = {nullptr, nullptr};
    size_t chunk_bytes = 0;
    cudaEvent_t chunk_copied[2] = {nullptr, nullptr}, chunk_done[2] = {nullptr, nullptr};

    std::vector<Slot> slots;
    uint64_t next_ticket = 1, next_collect = 1;

    airgpu_stats stats{};
    unsigned force_ordered = 0;                // AIRGPU_FORCE_ORDERED=1 (tests): every tile takes the ordered path
};

namespace {

int ensure_tiles(airgpu_ctx *c, size_t n_tiles)
{
    if (n_tiles <= c->tiles_cap) return AIRGPU_OK;
    CU(cudaDeviceSynchronize());
    if (c->tile_tab) cudaFree(c->tile_tab);
    if (c->group_sum) cudaFree(c->group_sum);
    c->tile_tab = nullptr;
    c->group_sum = nullptr;
    c->tiles_cap = 0;
    size_t want = std::max<size_t>(n_tiles, 1024);
    size_t groups = (want + kGroupTiles - 1) / kGroupTiles;
    CU(cudaMalloc(&c->tile_tab, want * sizeof(uint2)));
    CU(cudaMalloc(&c->group_sum, (1 + 2 * groups) * sizeof(unsigned long long)));
    c->tiles_cap = want;
    c->groups_cap = groups;
    return AIRGPU_OK;
}

int ensure_scratch(airgpu_ctx *c, size_t cap)
{
    if (cap <= c->scratch_cap) return AIRGPU_OK;
    CU(cudaDeviceSynchronize());
    if (c->scratch) cudaFree(c->scratch);
    c->scratch = nullptr;
    c->scratch_cap = 0;
    CU(cudaMalloc(&c->scratch, cap * (size_t)kSlotBytes));
    c->scratch_cap = cap;
    return AIRGPU_OK;
}

int ensure_out(airgpu_ctx *c, size_t cap)
{
    if (cap <= c->out_cap) return AIRGPU_OK;
    CU(cudaDeviceSynchronize());
    if (c->out_dev) cudaFree(c->out_dev);
    c->out_dev = nullptr;
    c->out_cap = 0;
    CU(cudaMalloc(&c->out_dev, cap * sizeof(airgpu_frame)));
    c->out_cap = cap;
    return AIRGPU_OK;
}

// Queue decode + ordering for one device-resident piece.  `d_total` is a device
// counter the ordered frames are appended at (so consecutive pieces concatenate in
// order without the host knowing the counts); it ends up holding the running total.
// The caller zeroes *d_total and counters[kCounterGate] at the start of a call.
int enqueue_piece(airgpu_ctx *c, const void *d_iq, size_t n, size_t seg, uint64_t base,
                  airgpu_frame *d_out, size_t cap, unsigned long long *d_total, cudaStream_t stream)
{
    Geometry g;
    if (!make_geometry(n, seg, g)) return fail(AIRGPU_ERR_INVALID, "capture too large for one call (%zu samples)", n);
    int rc;
    if ((rc = ensure_tiles(c, g.n_tiles)) != AIRGPU_OK) return rc;
    // scratch = kSlotsPerTile fixed record slots per tile + an overflow area as large as the output
    const size_t ovf_cap = std::max<size_t>(cap, 1);
    if ((rc = ensure_scratch(c, (size_t)g.n_tiles * kSlotsPerTile + ovf_cap)) != AIRGPU_OK) return rc;

    // the overflow index and the per-group sums restart with every piece (one memset: they are adjacent)
    CU(cudaMemsetAsync(c->group_sum, 0, (1 + c->groups_cap) * sizeof(unsigned long long), stream));

    DecodeParams p{};
    p.iq = d_iq;
    p.n_samples = n;
    p.seg_len = g.seg_len;
    p.tiles_per_seg = g.tiles_per_seg;
    p.n_tiles = g.n_tiles;
    p.base_offset = base;
    p.minus_one = 0xFFFFFFFFu;
    p.vec_ok = ((reinterpret_cast<uintptr_t>(d_iq) & 15u) == 0 && (g.n_tiles == g.tiles_per_seg || g.seg_len % 8 == 0)) ? 1u : 0u;
    p.scratch = c->scratch;
    p.cap = cap;
    p.ovf_cap = ovf_cap;
    p.counters = c->counters;
    p.tile_tab = c->tile_tab;
    p.ovf_counter = c->group_sum;
    p.group_sum = c->group_sum + 1;
    p.group_base = c->group_sum + 1 + c->groups_cap;
    p.force_ordered = c->force_ordered;
    CU(cudaEventRecord(c->evk0, stream));
    CU(launch_decode(c->format, p, stream));
    CU(cudaEventRecord(c->evk1, stream));
    c->evk_valid = true;
    CU(launch_finalize(p, d_out, d_total, stream));
    c->stats.n_tiles += g.n_tiles;
    c->stats.n_samples += n;
    return AIRGPU_OK;
}

int begin_call(airgpu_ctx *c, unsigned long long *d_total, cudaStream_t stream)
{
    c->stats = airgpu_stats{};
    CU(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), stream));
    CU(cudaMemsetAsync(c->counters + kCounterGate, 0, sizeof(unsigned long long), stream));
    return AIRGPU_OK;
}

}  // namespace

extern "C" {

const char *airgpu_version(void) { return "airgpu 0.1.0 (sm_100a, abi 1)"; }
const char *airgpu_last_error(void) { return g_err; }

int airgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void airgpu_destroy(airgpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->compute) cudaStreamSynchronize(c->compute);
    if (c->copy) cudaStreamSynchronize(c->copy);
    for (Slot &s : c->slots) {
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.d_in) cudaFree(s.d_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.h_count) cudaFreeHost(s.h_count);
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.done) cudaEventDestroy(s.done);
    }
    for (int b = 0; b < 2; ++b) {
        if (c->chunk_dev[b]) cudaFree(c->chunk_dev[b]);
        if (c->chunk_copied[b]) cudaEventDestroy(c->chunk_copied[b]);
        if (c->chunk_done[b]) cudaEventDestroy(c->chunk_done[b]);
    }
    if (c->scratch) cudaFree(c->scratch);
    if (c->tile_tab) cudaFree(c->tile_tab);
    if (c->group_sum) cudaFree(c->group_sum);
    if (c->counters) cudaFree(c->counters);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->out_dev) cudaFree(c->out_dev);
    for (cudaEvent_t e : {c->ev0, c->ev1, c->evh0, c->evh1, c->ev_sync, c->evk0, c->evk1})
        if (e) cudaEventDestroy(e);
    if (c->compute) cudaStreamDestroy(c->compute);
    if (c->copy) cudaStreamDestroy(c->copy);
    delete c;
}

int airgpu_create(const airgpu_config *cfg, airgpu_ctx **out)
{
    if (!out) return fail(AIRGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!cfg || cfg->struct_size < sizeof(airgpu_config))
        return fail(AIRGPU_ERR_INVALID, "airgpu_config.struct_size mismatch");
    if (cfg->format != AIRGPU_FMT_CS16 && cfg->format != AIRGPU_FMT_U8)
        return fail(AIRGPU_ERR_INVALID, "unknown sample format %u", cfg->format);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(AIRGPU_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(AIRGPU_ERR_INVALID, "device %d out of range (0..%d)", cfg->device, ndev - 1);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(AIRGPU_ERR_NO_DEVICE, "device %d is sm_%d%d; this build targets sm_100a only",
                    cfg->device, prop.major, prop.minor);
    CU(cudaSetDevice(cfg->device));

    airgpu_ctx *c = new (std::nothrow) airgpu_ctx();
    if (!c) return fail(AIRGPU_ERR_NOMEM, "out of host memory");
    c->device = cfg->device;
    c->format = (int)cfg->format;
    c->max_buffer_samples = cfg->max_buffer_samples ? cfg->max_buffer_samples : 262144;
    c->max_frames = cfg->max_frames ? cfg->max_frames : 8192;
    if (const char *e = std::getenv("AIRGPU_FORCE_ORDERED")) c->force_ordered = std::atoi(e) ? 1u : 0u;
    unsigned n_slots = cfg->ring_slots ? cfg->ring_slots : 4;

#define CUX(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            int rc_ = fail(e_ == cudaErrorMemoryAllocation ? AIRGPU_ERR_NOMEM : AIRGPU_ERR_CUDA,  \
                           "%s failed: %s", #call, cudaGetErrorString(e_));                       \
            airgpu_destroy(c);                                                                    \
            return rc_;                                                                           \
        }                                                                                         \
    } while (0)

    CUX(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
    CUX(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    CUX(cudaEventCreate(&c->ev0));
    CUX(cudaEventCreate(&c->ev1));
    CUX(cudaEventCreate(&c->evh0));
    CUX(cudaEventCreate(&c->evh1));
    CUX(cudaEventCreate(&c->evk0));
    CUX(cudaEventCreate(&c->evk1));
    CUX(cudaEventCreateWithFlags(&c->ev_sync, cudaEventDisableTiming));
    CUX(cudaMalloc(&c->counters, (kNumCounters + 1) * sizeof(unsigned long long)));
    CUX(cudaMemset(c->counters, 0, (kNumCounters + 1) * sizeof(unsigned long long)));
    CUX(cudaHostAlloc(&c->h_counters, (kNumCounters + 1) * sizeof(unsigned long long), cudaHostAllocDefault));
    c->slots.resize(n_slots);
    const size_t in_bytes = c->max_buffer_samples * bytes_per_sample(c->format);
    for (Slot &s : c->slots) {
        CUX(cudaHostAlloc(&s.h_in, in_bytes, cudaHostAllocDefault));
        CUX(cudaMalloc(&s.d_in, in_bytes));
        CUX(cudaHostAlloc(&s.h_out, c->max_frames * sizeof(airgpu_frame), cudaHostAllocDefault));
        CUX(cudaHostAlloc(&s.h_count, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
        CUX(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
        CUX(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    for (int b = 0; b < 2; ++b) {
        CUX(cudaEventCreateWithFlags(&c->chunk_copied[b], cudaEventDisableTiming));
        CUX(cudaEventCreateWithFlags(&c->chunk_done[b], cudaEventDisableTiming));
    }
#undef CUX
    *out = c;
    return AIRGPU_OK;
}

int airgpu_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(AIRGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return AIRGPU_OK;
}

int airgpu_host_free(void *p)
{
    if (p) CU(cudaFreeHost(p));
    return AIRGPU_OK;
}

// ---------------------------------------------------------------------------
// device-resident decode
// ---------------------------------------------------------------------------
int airgpu_decode_device(airgpu_ctx *c, const void *d_iq, size_t n_samples, size_t segment_samples,
                         uint64_t base_offset, airgpu_frame *d_out, size_t cap, uint64_t *d_count,
                         void *stream)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    if ((n_samples && !d_iq) || (cap && !d_out)) return fail(AIRGPU_ERR_INVALID, "NULL device pointer");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->compute;
    unsigned long long *total = d_count ? (unsigned long long *)d_count : c->counters + kNumCounters;
    int rc = begin_call(c, total, s);
    if (rc != AIRGPU_OK) return rc;
    CU(cudaEventRecord(c->ev0, s));
    rc = enqueue_piece(c, d_iq, n_samples, segment_samples, base_offset, d_out, cap, total, s);
    if (rc != AIRGPU_OK) return rc;
    CU(cudaEventRecord(c->ev1, s));
    c->ev_valid = true;
    c->evh_valid = false;
    if (!d_count) {
        // mirror the counters for airgpu_sync_count (a caller that passes d_count reads it itself)
        CU(cudaMemcpyAsync(c->h_counters, c->counters, kNumCounters * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(c->h_counters + kNumCounters, total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU(cudaEventRecord(c->ev_sync, s));
        c->sync_valid = true;
    } else {
        c->sync_valid = false;
    }
    return AIRGPU_OK;
}

int airgpu_sync_count(airgpu_ctx *c, uint64_t *n_frames)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    if (!c->sync_valid)
        return fail(AIRGPU_ERR_INVALID, "airgpu_sync_count: the last airgpu_decode_device was given d_count; read the count there");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->ev_sync));
    if (n_frames) *n_frames = c->h_counters[kNumCounters];
    c->stats.n_frames = c->h_counters[kNumCounters];
    c->stats.gate_passes = c->h_counters[kCounterGate];
    return AIRGPU_OK;
}

int airgpu_get_stats(airgpu_ctx *c, airgpu_stats *out)
{
    if (!c || !out) return fail(AIRGPU_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    if (c->ev_valid) {
        CU(cudaEventSynchronize(c->ev1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->stats.kernel_ms = ms;
    }
    if (c->evk_valid) {
        float ms = 0.f;
        CU(cudaEventSynchronize(c->evk1));
        CU(cudaEventElapsedTime(&ms, c->evk0, c->evk1));
        c->stats.decode_ms = ms;
    }
    if (c->evh_valid) {
        float ms = 0.f;
        CU(cudaEventSynchronize(c->evh1));
        CU(cudaEventElapsedTime(&ms, c->evh0, c->evh1));
        c->stats.h2d_ms = ms;
    }
    *out = c->stats;
    return AIRGPU_OK;
}

// ---------------------------------------------------------------------------
// host-buffer decode: chunk pipeline
// ---------------------------------------------------------------------------
int airgpu_decode(airgpu_ctx *c, const void *iq, size_t n_samples, size_t segment_samples,
                  uint64_t base_offset, airgpu_frame *out, size_t cap, size_t *n_frames)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    if (n_frames) *n_frames = 0;
    if ((n_samples && !iq) || (cap && !out)) return fail(AIRGPU_ERR_INVALID, "NULL buffer");
    CU(cudaSetDevice(c->device));
    const size_t bps = bytes_per_sample(c->format);
    const size_t seg = (segment_samples == 0 || segment_samples > n_samples) ? n_samples : segment_samples;
    int rc;
    if ((rc = ensure_out(c, std::max<size_t>(cap, 1))) != AIRGPU_OK) return rc;

    // Pieces: at most kChunkSamples (+ overlap) samples are on the device at a time.
    // Short segments travel whole, several per piece; a long segment is cut into
    // contiguous candidate ranges, each carrying the 240 samples its last candidate
    // reads, and decoded as a segment of its own (same candidates, same frames).
    const size_t kChunkSamples = (size_t)32 << 20;   // 64 MiB of U8 / 128 MiB of CS16
    struct Piece {
        size_t first, n, seg;
        uint64_t base;
    };
    std::vector<Piece> pieces;
    size_t max_piece = 0;
    if (seg > (size_t)kFrameSamples) {
        if (seg <= kChunkSamples) {
            const size_t per = std::max<size_t>(1, kChunkSamples / seg);
            const size_t n_seg = (n_samples + seg - 1) / seg;
            for (size_t s0 = 0; s0 < n_seg; s0 += per) {
                const size_t first = s0 * seg;
                pieces.push_back({first, std::min(n_samples - first, per * seg), seg, base_offset + first});
            }
        } else {
            for (size_t s0 = 0; s0 < n_samples; s0 += seg) {
                const size_t len = std::min(seg, n_samples - s0);
                if (len <= (size_t)kFrameSamples) continue;
                const size_t cands = len - kFrameSamples;
                for (size_t a = 0; a < cands; a += kChunkSamples) {
                    const size_t cnum = std::min(kChunkSamples, cands - a);
                    pieces.push_back({s0 + a, cnum + kFrameSamples, cnum + kFrameSamples, base_offset + s0 + a});
                }
            }
        }
    }
    for (const Piece &pc : pieces) max_piece = std::max(max_piece, pc.n);

    const size_t need_bytes = max_piece * bps;
    if (need_bytes > c->chunk_bytes) {
        CU(cudaDeviceSynchronize());
        for (int b = 0; b < 2; ++b) {
            if (c->chunk_dev[b]) cudaFree(c->chunk_dev[b]);
            c->chunk_dev[b] = nullptr;
        }
        c->chunk_bytes = 0;
        for (int b = 0; b < 2; ++b) CU(cudaMalloc(&c->chunk_dev[b], need_bytes));
        c->chunk_bytes = need_bytes;
    }

    unsigned long long *total = c->counters + kNumCounters;
    if ((rc = begin_call(c, total, c->compute)) != AIRGPU_OK) return rc;
    CU(cudaEventRecord(c->ev0, c->compute));
    CU(cudaEventRecord(c->evh0, c->copy));
    for (size_t k = 0; k < pieces.size(); ++k) {
        const Piece &pc = pieces[k];
        const int b = (int)(k & 1);
        const char *src = static_cast<const char *>(iq) + pc.first * bps;
        if (k >= 2) CU(cudaStreamWaitEvent(c->copy, c->chunk_done[b], 0));   // device buffer b is free again
        // page-locked sources (airgpu_host_alloc) copy asynchronously; pageable ones are staged by the driver
        CU(cudaMemcpyAsync(c->chunk_dev[b], src, pc.n * bps, cudaMemcpyHostToDevice, c->copy));
        CU(cudaEventRecord(c->chunk_copied[b], c->copy));
        CU(cudaStreamWaitEvent(c->compute, c->chunk_copied[b], 0));
        rc = enqueue_piece(c, c->chunk_dev[b], pc.n, pc.seg, pc.base, c->out_dev, cap, total, c->compute);
        if (rc != AIRGPU_OK) return rc;
        CU(cudaEventRecord(c->chunk_done[b], c->compute));
    }
    CU(cudaEventRecord(c->evh1, c->copy));
    CU(cudaEventRecord(c->ev1, c->compute));
    c->ev_valid = true;
    c->evh_valid = true;

    CU(cudaMemcpyAsync(c->h_counters, c->counters, kNumCounters * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, c->compute));
    CU(cudaMemcpyAsync(c->h_counters + kNumCounters, total, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       c->compute));
    CU(cudaEventRecord(c->ev_sync, c->compute));
    CU(cudaStreamSynchronize(c->compute));
    const unsigned long long n = c->h_counters[kNumCounters];
    c->stats.n_frames = n;
    c->stats.gate_passes = c->h_counters[kCounterGate];
    if (n_frames) *n_frames = (size_t)n;
    const size_t ncopy = (size_t)std::min<unsigned long long>(n, cap);
    if (ncopy) CU(cudaMemcpy(out, c->out_dev, ncopy * sizeof(airgpu_frame), cudaMemcpyDeviceToHost));
    if (n > cap) return fail(AIRGPU_ERR_OVERFLOW, "%llu frames but capacity %zu", n, cap);
    return AIRGPU_OK;
}

// ---------------------------------------------------------------------------
// streaming ring
// ---------------------------------------------------------------------------
int airgpu_submit(airgpu_ctx *c, const void *iq, size_t n_samples, uint64_t base_offset, uint64_t *ticket)
{
    if (!c || !ticket) return fail(AIRGPU_ERR_INVALID, "NULL argument");
    if (n_samples && !iq) return fail(AIRGPU_ERR_INVALID, "iq is NULL");
    if (n_samples > c->max_buffer_samples)
        return fail(AIRGPU_ERR_INVALID, "buffer of %zu samples exceeds max_buffer_samples=%zu", n_samples,
                    c->max_buffer_samples);
    CU(cudaSetDevice(c->device));
    Slot &s = c->slots[(c->next_ticket - 1) % c->slots.size()];
    if (s.busy)
        return fail(AIRGPU_ERR_BUSY, "ring full (%zu slots): collect ticket %llu first", c->slots.size(),
                    (unsigned long long)c->next_collect);
    const size_t bytes = n_samples * bytes_per_sample(c->format);
    int rc;
    if ((rc = ensure_out(c, c->max_frames)) != AIRGPU_OK) return rc;
    if (bytes) memcpy(s.h_in, iq, bytes);
    if (bytes) CU(cudaMemcpyAsync(s.d_in, s.h_in, bytes, cudaMemcpyHostToDevice, c->copy));
    CU(cudaEventRecord(s.copied, c->copy));
    CU(cudaStreamWaitEvent(c->compute, s.copied, 0));
    unsigned long long *total = c->counters + kNumCounters;
    if ((rc = begin_call(c, total, c->compute)) != AIRGPU_OK) return rc;
    rc = enqueue_piece(c, s.d_in, n_samples, 0, base_offset, c->out_dev, c->max_frames, total, c->compute);
    if (rc != AIRGPU_OK) return rc;
    CU(cudaMemcpyAsync(s.h_count, total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->compute));
    CU(cudaMemcpyAsync(s.h_count + 1, c->counters + kCounterGate, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       c->compute));
    CU(cudaMemcpyAsync(s.h_out, c->out_dev, c->max_frames * sizeof(airgpu_frame), cudaMemcpyDeviceToHost, c->compute));
    CU(cudaEventRecord(s.done, c->compute));
    c->ev_valid = false;
    c->evh_valid = false;
    s.busy = true;
    s.ticket = c->next_ticket;
    *ticket = c->next_ticket++;
    return AIRGPU_OK;
}

int airgpu_collect(airgpu_ctx *c, uint64_t ticket, airgpu_frame *out, size_t cap, size_t *n_frames)
{
    if (!c || !n_frames) return fail(AIRGPU_ERR_INVALID, "NULL argument");
    *n_frames = 0;
    if (ticket != c->next_collect)
        return fail(AIRGPU_ERR_TICKET, "tickets are collected in submission order: expected %llu, got %llu",
                    (unsigned long long)c->next_collect, (unsigned long long)ticket);
    Slot &s = c->slots[(ticket - 1) % c->slots.size()];
    if (!s.busy || s.ticket != ticket)
        return fail(AIRGPU_ERR_TICKET, "ticket %llu was never submitted", (unsigned long long)ticket);
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(s.done));
    const unsigned long long n = s.h_count[0];
    c->stats.n_frames = n;
    c->stats.gate_passes = s.h_count[1];
    s.busy = false;
    c->next_collect++;
    *n_frames = (size_t)n;
    const size_t have = (size_t)std::min<unsigned long long>(n, c->max_frames);
    const size_t ncopy = std::min(have, cap);
    if (ncopy && !out) return fail(AIRGPU_ERR_INVALID, "out is NULL");
    if (ncopy) memcpy(out, s.h_out, ncopy * sizeof(airgpu_frame));
    if (n > ncopy)
        return fail(AIRGPU_ERR_OVERFLOW, "%llu frames but capacity %zu (max_frames=%zu)", n, cap, c->max_frames);
    return AIRGPU_OK;
}

// ---------------------------------------------------------------------------
// N1: frame field decode
// ---------------------------------------------------------------------------
int airgpu_decode_fields(airgpu_ctx *c, const airgpu_frame *d_frames, size_t n_frames, airgpu_fields *d_out,
                         void *stream)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    if (n_frames && (!d_frames || !d_out)) return fail(AIRGPU_ERR_INVALID, "NULL device pointer");
    CU(cudaSetDevice(c->device));
    CU(launch_decode_fields(d_frames, n_frames, d_out, stream ? (cudaStream_t)stream : c->compute));
    return AIRGPU_OK;
}

int airgpu_decode_fields_host(airgpu_ctx *c, const airgpu_frame *frames, size_t n_frames, airgpu_fields *out)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    if (n_frames == 0) return AIRGPU_OK;
    if (!frames || !out) return fail(AIRGPU_ERR_INVALID, "NULL buffer");
    CU(cudaSetDevice(c->device));
    airgpu_frame *d_in = nullptr;
    airgpu_fields *d_out = nullptr;
    CU(cudaMalloc(&d_in, n_frames * sizeof(airgpu_frame)));
    cudaError_t e = cudaMalloc(&d_out, n_frames * sizeof(airgpu_fields));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, frames, n_frames * sizeof(airgpu_frame), cudaMemcpyHostToDevice, c->compute);
    if (e == cudaSuccess) e = launch_decode_fields(d_in, n_frames, d_out, c->compute);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n_frames * sizeof(airgpu_fields), cudaMemcpyDeviceToHost, c->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute);
    cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    CU(e);
    return AIRGPU_OK;
}

// ---------------------------------------------------------------------------
// diagnostics used by the parity tests (device arithmetic only)
// ---------------------------------------------------------------------------
int airgpu_dbg_levels_u8(airgpu_ctx *c, uint16_t *out65536)
{
    if (!c || !out65536) return fail(AIRGPU_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    uint16_t *d = nullptr;
    CU(cudaMalloc(&d, 65536 * sizeof(uint16_t)));
    cudaError_t e = launch_levels_u8(d, c->compute);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out65536, d, 65536 * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute);
    cudaFree(d);
    CU(e);
    return AIRGPU_OK;
}

int airgpu_dbg_levels_cs16(airgpu_ctx *c, const int16_t *iq, size_t n_samples, uint16_t *out)
{
    if (!c || (n_samples && (!iq || !out))) return fail(AIRGPU_ERR_INVALID, "NULL argument");
    if (n_samples == 0) return AIRGPU_OK;
    CU(cudaSetDevice(c->device));
    int16_t *d_in = nullptr;
    uint16_t *d_out = nullptr;
    CU(cudaMalloc(&d_in, n_samples * 4));
    cudaError_t e = cudaMalloc(&d_out, n_samples * 2);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, iq, n_samples * 4, cudaMemcpyHostToDevice, c->compute);
    if (e == cudaSuccess) e = launch_levels_cs16(d_in, n_samples, d_out, c->compute);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n_samples * 2, cudaMemcpyDeviceToHost, c->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->compute);
    cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    CU(e);
    return AIRGPU_OK;
}

}  // extern "C"
