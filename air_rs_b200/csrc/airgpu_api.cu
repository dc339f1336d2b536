// airgpu_api.cu -- the C ABI declared in include/airgpu.h.
//
// Host-side plumbing only: context, device workspace, the pinned ring used by
// airgpu_submit/collect, the chunk pipeline of airgpu_decode (H2D on a copy stream
// overlapped with the decode kernels on a compute stream), CUDA-graph capture, the
// multi-destination (peer / multicast) form of the ordering kernels and the
// single-thread multi-GPU group.  All decode arithmetic lives in airgpu_kernels.cu;
// there is no CPU fallback anywhere.
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

constexpr size_t kHeadFrames = 256;     // records of a ring buffer that always travel with the count (6 KB)

struct Slot {
    void *h_in = nullptr;          // pinned
    void *d_in = nullptr;
    airgpu_frame *d_out = nullptr; // device: room for the worst case (a frame at every offset)
    airgpu_frame *h_head = nullptr;          // pinned, kHeadFrames records
    unsigned long long *h_count = nullptr;   // pinned: frames, gate passes
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
    size_t ring_cap = 0;                       // records per ring slot on the device
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evh0 = nullptr, evh1 = nullptr, ev_sync = nullptr;
    bool ev_valid = false, evh_valid = false, sync_valid = false;
    // event pairs around the decode-kernel launches since the last airgpu_get_stats (a ring: the newest kKernelEvents)
    static constexpr size_t kKernelEvents = 64;
    cudaEvent_t kev0[kKernelEvents] = {}, kev1[kKernelEvents] = {};
    size_t kev_n = 0;
    bool timing = true;                        // record the CUDA events behind airgpu_stats
    bool capturing = false;                    // between airgpu_graph_begin and airgpu_graph_end

    // workspace shared by every decode on the compute stream (stream-ordered reuse)
    void *scratch = nullptr;                   // kSlotBytes per slot
    size_t scratch_cap = 0;                    // slots
    uint2 *tile_tab = nullptr;
    // one allocation: [frame total | counters (kNumCounters) | overflow counter | per-group sums | per-group bases];
    // what a call has to zero is contiguous, so a single-piece call issues ONE memset
    unsigned long long *ws = nullptr;
    size_t groups_cap = 0;
    size_t tiles_cap = 0;
    unsigned long long *h_counters = nullptr; // pinned mirror: [total | counters]
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

    unsigned long long *total() const { return ws; }
    unsigned long long *counters() const { return ws + 1; }
    unsigned long long *ovf_counter() const { return ws + 1 + kNumCounters; }
    unsigned long long *group_sum() const { return ws + 2 + kNumCounters; }
    unsigned long long *group_base() const { return ws + 2 + kNumCounters + groups_cap; }
};

struct airgpu_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int device = 0;
};

namespace {

int ensure_tiles(airgpu_ctx *c, size_t n_tiles)
{
    if (n_tiles <= c->tiles_cap) return AIRGPU_OK;
    if (c->capturing) return fail(AIRGPU_ERR_INVALID, "workspace too small inside a graph capture: call airgpu_reserve first");
    CU(cudaDeviceSynchronize());
    if (c->tile_tab) cudaFree(c->tile_tab);
    if (c->ws) cudaFree(c->ws);
    c->tile_tab = nullptr;
    c->ws = nullptr;
    c->tiles_cap = 0;
    size_t want = std::max<size_t>(n_tiles, 1024);
    size_t groups = (want + kGroupTiles - 1) / kGroupTiles;
    CU(cudaMalloc(&c->tile_tab, want * sizeof(uint2)));
    CU(cudaMalloc(&c->ws, (2 + kNumCounters + 2 * groups) * sizeof(unsigned long long)));
    CU(cudaMemset(c->ws, 0, (2 + kNumCounters + 2 * groups) * sizeof(unsigned long long)));
    c->tiles_cap = want;
    c->groups_cap = groups;
    return AIRGPU_OK;
}

int ensure_scratch(airgpu_ctx *c, size_t cap)
{
    if (cap <= c->scratch_cap) return AIRGPU_OK;
    if (c->capturing) return fail(AIRGPU_ERR_INVALID, "scratch too small inside a graph capture: call airgpu_reserve first");
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
    if (c->capturing) return fail(AIRGPU_ERR_INVALID, "output buffer too small inside a graph capture");
    CU(cudaDeviceSynchronize());
    if (c->out_dev) cudaFree(c->out_dev);
    c->out_dev = nullptr;
    c->out_cap = 0;
    CU(cudaMalloc(&c->out_dev, cap * sizeof(airgpu_frame)));
    c->out_cap = cap;
    return AIRGPU_OK;
}

// scratch = kSlotsPerTile fixed slots per tile + an overflow area.  The overflow area only ever holds the frames
// of tiles with more than kSlotsPerTile of them, so it never needs more than the candidates of the call, nor
// more than the output can take.
size_t overflow_cap(size_t n_samples, size_t cap) { return std::max<size_t>(std::min(cap, n_samples), 1); }

OutSet single_out(airgpu_frame *d_out)
{
    OutSet o{};
    o.out[0] = reinterpret_cast<unsigned long long *>(d_out);
    o.n = 1;
    return o;
}

// Queue decode + ordering for one device-resident piece.  `d_total` is a device counter the ordered frames are
// appended at (so consecutive pieces concatenate in order without the host knowing the counts); it ends up holding
// the running total.  `first` zeroes it and the gate counter together with the per-piece sums (one memset when
// d_total is the context's own counter: pass nullptr for that -- the workspace may move when it grows).
int enqueue_piece(airgpu_ctx *c, const void *d_iq, size_t n, size_t seg, uint64_t base, const OutSet &dst, size_t cap,
                  unsigned long long *d_total, bool first, cudaStream_t stream)
{
    Geometry g;
    if (!make_geometry(n, seg, g)) return fail(AIRGPU_ERR_INVALID, "capture too large for one call (%zu samples)", n);
    int rc;
    if ((rc = ensure_tiles(c, g.n_tiles)) != AIRGPU_OK) return rc;
    const size_t ovf_cap = overflow_cap(n, cap);
    if ((rc = ensure_scratch(c, (size_t)g.n_tiles * kSlotsPerTile + ovf_cap)) != AIRGPU_OK) return rc;
    if (!d_total) d_total = c->total();

    const size_t n_groups = ((size_t)g.n_tiles + kGroupTiles - 1) / kGroupTiles;
    if (first) {
        c->stats = airgpu_stats{};
        // [total | counters | overflow counter | sums of this piece's groups]
        CU(cudaMemsetAsync(c->ws, 0, (2 + kNumCounters + n_groups) * sizeof(unsigned long long), stream));
        if (d_total != c->total()) CU(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), stream));
    } else {
        CU(cudaMemsetAsync(c->ovf_counter(), 0, (1 + n_groups) * sizeof(unsigned long long), stream));
    }

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
    p.counters = c->counters();
    p.tile_tab = c->tile_tab;
    p.ovf_counter = c->ovf_counter();
    p.group_sum = c->group_sum();
    p.group_base = c->group_base();
    p.force_ordered = c->force_ordered;
    const bool timed = c->timing && !c->capturing;
    const size_t kslot = c->kev_n % airgpu_ctx::kKernelEvents;
    if (timed) CU(cudaEventRecord(c->kev0[kslot], stream));
    CU(launch_decode(c->format, p, stream));
    if (timed) {
        CU(cudaEventRecord(c->kev1[kslot], stream));
        c->kev_n++;
    }
    CU(launch_finalize(p, dst, d_total, stream));
    c->stats.n_tiles += g.n_tiles;
    c->stats.n_samples += n;
    return AIRGPU_OK;
}

// [total | counters] -> pinned mirror, one copy
int mirror_counters(airgpu_ctx *c, cudaStream_t s)
{
    CU(cudaMemcpyAsync(c->h_counters, c->ws, (1 + kNumCounters) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    return AIRGPU_OK;
}

void destroy_ctx(airgpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->compute) cudaStreamSynchronize(c->compute);
    if (c->copy) cudaStreamSynchronize(c->copy);
    for (Slot &s : c->slots) {
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_head) cudaFreeHost(s.h_head);
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
    if (c->ws) cudaFree(c->ws);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->out_dev) cudaFree(c->out_dev);
    for (cudaEvent_t e : {c->ev0, c->ev1, c->evh0, c->evh1, c->ev_sync})
        if (e) cudaEventDestroy(e);
    for (size_t k = 0; k < airgpu_ctx::kKernelEvents; ++k) {
        if (c->kev0[k]) cudaEventDestroy(c->kev0[k]);
        if (c->kev1[k]) cudaEventDestroy(c->kev1[k]);
    }
    if (c->compute) cudaStreamDestroy(c->compute);
    if (c->copy) cudaStreamDestroy(c->copy);
    delete c;
}

// ---- host-buffer decode, split so that a group can queue every device before it waits for any ----
struct HostPiece {
    size_t first, n, seg;
    uint64_t base;
};

// Pieces: at most kChunkSamples (+ overlap) samples are on the device at a time.  Short segments travel whole,
// several per piece; a long segment is cut into contiguous candidate ranges, each carrying the 240 samples its
// last candidate reads, and decoded as a segment of its own (same candidates, same frames).
std::vector<HostPiece> plan_pieces(size_t n_samples, size_t seg, uint64_t base_offset)
{
    const size_t kChunkSamples = (size_t)32 << 20;   // 64 MiB of U8 / 128 MiB of CS16
    std::vector<HostPiece> pieces;
    if (seg <= (size_t)kFrameSamples) return pieces;
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
    return pieces;
}

// everything up to (not including) the wait: H2D chunk pipeline + kernels + the counter mirror
int decode_host_enqueue(airgpu_ctx *c, const void *iq, size_t n_samples, size_t segment_samples, uint64_t base_offset, size_t cap)
{
    CU(cudaSetDevice(c->device));
    const size_t bps = bytes_per_sample(c->format);
    const size_t seg = (segment_samples == 0 || segment_samples > n_samples) ? n_samples : segment_samples;
    int rc;
    if ((rc = ensure_out(c, std::max<size_t>(cap, 1))) != AIRGPU_OK) return rc;
    if ((rc = ensure_tiles(c, 1)) != AIRGPU_OK) return rc;
    const std::vector<HostPiece> pieces = plan_pieces(n_samples, seg, base_offset);
    size_t max_piece = 0;
    for (const HostPiece &pc : pieces) max_piece = std::max(max_piece, pc.n);
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
    if (pieces.empty()) {
        c->stats = airgpu_stats{};
        CU(cudaMemsetAsync(c->ws, 0, (1 + kNumCounters) * sizeof(unsigned long long), c->compute));
    }
    if (c->timing) {
        CU(cudaEventRecord(c->ev0, c->compute));
        CU(cudaEventRecord(c->evh0, c->copy));
    }
    const OutSet dst = single_out(c->out_dev);
    for (size_t k = 0; k < pieces.size(); ++k) {
        const HostPiece &pc = pieces[k];
        const int b = (int)(k & 1);
        const char *src = static_cast<const char *>(iq) + pc.first * bps;
        if (k >= 2) CU(cudaStreamWaitEvent(c->copy, c->chunk_done[b], 0));   // device buffer b is free again
        // page-locked sources (airgpu_host_alloc) copy asynchronously; pageable ones are staged by the driver
        CU(cudaMemcpyAsync(c->chunk_dev[b], src, pc.n * bps, cudaMemcpyHostToDevice, c->copy));
        CU(cudaEventRecord(c->chunk_copied[b], c->copy));
        CU(cudaStreamWaitEvent(c->compute, c->chunk_copied[b], 0));
        rc = enqueue_piece(c, c->chunk_dev[b], pc.n, pc.seg, pc.base, dst, cap, nullptr, k == 0, c->compute);
        if (rc != AIRGPU_OK) return rc;
        CU(cudaEventRecord(c->chunk_done[b], c->compute));
    }
    if (c->timing) {
        CU(cudaEventRecord(c->evh1, c->copy));
        CU(cudaEventRecord(c->ev1, c->compute));
    }
    c->ev_valid = c->timing;
    c->evh_valid = c->timing;
    c->sync_valid = false;
    return mirror_counters(c, c->compute);
}

// wait for the count, then queue the copy of exactly the records that exist (the caller synchronises)
int decode_host_finish(airgpu_ctx *c, airgpu_frame *out, size_t cap, size_t *n_frames)
{
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->compute));
    const unsigned long long n = c->h_counters[0];
    c->stats.n_frames = n;
    c->stats.gate_passes = c->h_counters[1 + kCounterGate];
    if (n_frames) *n_frames = (size_t)n;
    const size_t ncopy = (size_t)std::min<unsigned long long>(n, cap);
    if (ncopy) CU(cudaMemcpyAsync(out, c->out_dev, ncopy * sizeof(airgpu_frame), cudaMemcpyDeviceToHost, c->compute));
    return AIRGPU_OK;
}

}  // namespace

extern "C" {

const char *airgpu_version(void) { return "airgpu 0.2.0 (sm_100a, abi 2)"; }
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

size_t airgpu_playback_samples(size_t len, size_t chunk)
{
    if (chunk == 0 || len == 0) return 0;
    return ((len - 1) / chunk) * chunk;      // src/adsb.rs:77: while i < data.len() - chunk
}

void airgpu_destroy(airgpu_ctx *c) { destroy_ctx(c); }

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
    // a constant buffer yields a frame at EVERY offset (ties pass the gate, crc(0) = 0): the device side of the ring
    // always has room for that, so that nothing the reference would send is ever dropped (src/adsb.rs:98-111)
    c->ring_cap = std::max<size_t>(c->max_frames,
                                   c->max_buffer_samples > (size_t)kFrameSamples ? c->max_buffer_samples - kFrameSamples : 1);
    if (const char *e = std::getenv("AIRGPU_FORCE_ORDERED")) c->force_ordered = std::atoi(e) ? 1u : 0u;
    unsigned n_slots = cfg->ring_slots ? cfg->ring_slots : 4;

#define CUX(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            int rc_ = fail(e_ == cudaErrorMemoryAllocation ? AIRGPU_ERR_NOMEM : AIRGPU_ERR_CUDA,  \
                           "%s failed: %s", #call, cudaGetErrorString(e_));                       \
            destroy_ctx(c);                                                                       \
            return rc_;                                                                           \
        }                                                                                         \
    } while (0)

    CUX(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
    CUX(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    CUX(cudaEventCreate(&c->ev0));
    CUX(cudaEventCreate(&c->ev1));
    CUX(cudaEventCreate(&c->evh0));
    CUX(cudaEventCreate(&c->evh1));
    for (size_t k = 0; k < airgpu_ctx::kKernelEvents; ++k) {
        CUX(cudaEventCreate(&c->kev0[k]));
        CUX(cudaEventCreate(&c->kev1[k]));
    }
    CUX(cudaEventCreateWithFlags(&c->ev_sync, cudaEventDisableTiming));
    CUX(cudaHostAlloc(&c->h_counters, (kNumCounters + 1) * sizeof(unsigned long long), cudaHostAllocDefault));
    memset(c->h_counters, 0, (kNumCounters + 1) * sizeof(unsigned long long));
    c->slots.resize(n_slots);
    const size_t in_bytes = c->max_buffer_samples * bytes_per_sample(c->format);
    for (Slot &s : c->slots) {
        CUX(cudaHostAlloc(&s.h_in, in_bytes, cudaHostAllocDefault));
        CUX(cudaMalloc(&s.d_in, in_bytes));
        CUX(cudaMalloc(&s.d_out, c->ring_cap * sizeof(airgpu_frame)));
        CUX(cudaHostAlloc(&s.h_head, kHeadFrames * sizeof(airgpu_frame), cudaHostAllocDefault));
        CUX(cudaHostAlloc(&s.h_count, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
        CUX(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
        CUX(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    for (int b = 0; b < 2; ++b) {
        CUX(cudaEventCreateWithFlags(&c->chunk_copied[b], cudaEventDisableTiming));
        CUX(cudaEventCreateWithFlags(&c->chunk_done[b], cudaEventDisableTiming));
    }
#undef CUX
    // the ring's own needs, so that a streaming host never allocates after this point
    int rc = airgpu_reserve(c, c->max_buffer_samples, 0, c->ring_cap);
    if (rc != AIRGPU_OK) {
        destroy_ctx(c);
        return rc;
    }
    *out = c;
    return AIRGPU_OK;
}

int airgpu_reserve(airgpu_ctx *c, size_t n_samples, size_t segment_samples, size_t cap)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    Geometry g;
    if (!make_geometry(n_samples, segment_samples, g)) return fail(AIRGPU_ERR_INVALID, "capture too large (%zu samples)", n_samples);
    int rc;
    if ((rc = ensure_tiles(c, g.n_tiles)) != AIRGPU_OK) return rc;
    return ensure_scratch(c, (size_t)g.n_tiles * kSlotsPerTile + overflow_cap(n_samples, cap));
}

int airgpu_set_timing(airgpu_ctx *c, int enabled)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    c->timing = enabled != 0;
    if (!c->timing) {
        c->ev_valid = c->evh_valid = false;
        c->kev_n = 0;
    }
    return AIRGPU_OK;
}

int airgpu_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(AIRGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    // portable: page-locked for every device of the process (airgpu_group_decode copies one buffer to several GPUs)
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
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
static int decode_device_common(airgpu_ctx *c, const void *d_iq, size_t n_samples, size_t segment_samples, uint64_t base_offset,
                                const OutSet &dst, size_t cap, unsigned long long *total, bool mirror, cudaStream_t s)
{
    const bool timed = c->timing && !c->capturing;
    if (timed) CU(cudaEventRecord(c->ev0, s));
    int rc = enqueue_piece(c, d_iq, n_samples, segment_samples, base_offset, dst, cap, total, true, s);
    if (rc != AIRGPU_OK) return rc;
    if (timed) CU(cudaEventRecord(c->ev1, s));
    c->ev_valid = timed;
    c->evh_valid = false;
    if (mirror && !c->capturing) {
        // mirror the counters for airgpu_sync_count (a caller that passes d_count reads it itself)
        if ((rc = mirror_counters(c, s)) != AIRGPU_OK) return rc;
        CU(cudaEventRecord(c->ev_sync, s));
        c->sync_valid = true;
    } else {
        c->sync_valid = false;
    }
    return AIRGPU_OK;
}

int airgpu_decode_device(airgpu_ctx *c, const void *d_iq, size_t n_samples, size_t segment_samples,
                         uint64_t base_offset, airgpu_frame *d_out, size_t cap, uint64_t *d_count,
                         void *stream)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    if ((n_samples && !d_iq) || (cap && !d_out)) return fail(AIRGPU_ERR_INVALID, "NULL device pointer");
    CU(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->compute;
    unsigned long long *total = (unsigned long long *)d_count;       // nullptr: the context's own counter
    return decode_device_common(c, d_iq, n_samples, segment_samples, base_offset, single_out(d_out), cap, total, d_count == nullptr, s);
}

int airgpu_decode_device_peers(airgpu_ctx *c, const void *d_iq, size_t n_samples, size_t segment_samples,
                               uint64_t base_offset, const airgpu_peers *dst, size_t cap, void *stream)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    if (!dst || dst->struct_size < sizeof(airgpu_peers)) return fail(AIRGPU_ERR_INVALID, "airgpu_peers.struct_size mismatch");
    if (dst->n_outs < 1 || dst->n_outs > AIRGPU_MAX_PEERS) return fail(AIRGPU_ERR_INVALID, "n_outs must be 1..%d", AIRGPU_MAX_PEERS);
    if (dst->multicast && dst->n_outs != 1) return fail(AIRGPU_ERR_INVALID, "a multicast exchange has exactly one destination address");
    if (n_samples && !d_iq) return fail(AIRGPU_ERR_INVALID, "NULL device pointer");
    OutSet o{};
    o.n = dst->n_outs;
    o.multicast = dst->multicast ? 1u : 0u;
    for (unsigned j = 0; j < dst->n_outs; ++j) {
        if (cap && !dst->outs[j]) return fail(AIRGPU_ERR_INVALID, "outs[%u] is NULL", j);
        o.out[j] = reinterpret_cast<unsigned long long *>(dst->outs[j]);
        o.count[j] = reinterpret_cast<unsigned long long *>(dst->counts[j]);
    }
    CU(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->compute;
    // the running total lives in the context; the count reaches the destinations through the scan kernel
    return decode_device_common(c, d_iq, n_samples, segment_samples, base_offset, o, cap, nullptr, true, s);
}

int airgpu_peer_barrier(airgpu_ctx *c, uint64_t *const *flags, uint32_t n_ranks, uint32_t rank, uint64_t epoch, void *stream)
{
    if (!c || !flags) return fail(AIRGPU_ERR_INVALID, "NULL argument");
    if (n_ranks < 1 || n_ranks > AIRGPU_MAX_PEERS || rank >= n_ranks) return fail(AIRGPU_ERR_INVALID, "bad rank %u of %u", rank, n_ranks);
    PeerFlags f{};
    for (unsigned q = 0; q < n_ranks; ++q) {
        if (!flags[q]) return fail(AIRGPU_ERR_INVALID, "flags[%u] is NULL", q);
        f.flags[q] = reinterpret_cast<unsigned long long *>(flags[q]);
    }
    f.n_ranks = n_ranks;
    f.rank = rank;
    f.epoch = epoch;
    CU(cudaSetDevice(c->device));
    CU(launch_peer_barrier(f, stream ? (cudaStream_t)stream : c->compute));
    return AIRGPU_OK;
}

int airgpu_sync_count(airgpu_ctx *c, uint64_t *n_frames)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    if (!c->sync_valid)
        return fail(AIRGPU_ERR_INVALID, "airgpu_sync_count: the last airgpu_decode_device was given d_count; read the count there");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->ev_sync));
    if (n_frames) *n_frames = c->h_counters[0];
    c->stats.n_frames = c->h_counters[0];
    c->stats.gate_passes = c->h_counters[1 + kCounterGate];
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
    if (c->kev_n) {
        // average over the decode-kernel launches since the previous call (the newest kKernelEvents of them)
        const size_t have = std::min(c->kev_n, airgpu_ctx::kKernelEvents);
        double sum = 0.0;
        for (size_t k = 0; k < have; ++k) {
            const size_t slot = (c->kev_n - 1 - k) % airgpu_ctx::kKernelEvents;
            float ms = 0.f;
            CU(cudaEventSynchronize(c->kev1[slot]));
            CU(cudaEventElapsedTime(&ms, c->kev0[slot], c->kev1[slot]));
            sum += ms;
        }
        c->stats.decode_ms = (float)(sum / (double)have);
        c->stats.decode_launches = (float)have;
        c->kev_n = 0;
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
// CUDA graphs
// ---------------------------------------------------------------------------
int airgpu_graph_begin(airgpu_ctx *c, void *stream)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    if (!stream) return fail(AIRGPU_ERR_INVALID, "graph capture needs an explicit stream");
    if (c->capturing) return fail(AIRGPU_ERR_INVALID, "a capture is already in progress on this context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    return AIRGPU_OK;
}

int airgpu_graph_end(airgpu_ctx *c, void *stream, airgpu_graph **out)
{
    if (!c || !out) return fail(AIRGPU_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (!c->capturing) return fail(AIRGPU_ERR_INVALID, "no capture in progress");
    c->capturing = false;
    CU(cudaSetDevice(c->device));
    cudaGraph_t g = nullptr;
    CU(cudaStreamEndCapture((cudaStream_t)stream, &g));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(g);
        CU(e);
    }
    airgpu_graph *h = new (std::nothrow) airgpu_graph();
    if (!h) {
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(g);
        return fail(AIRGPU_ERR_NOMEM, "out of host memory");
    }
    h->graph = g;
    h->exec = exec;
    h->device = c->device;
    *out = h;
    return AIRGPU_OK;
}

int airgpu_set_capturing(airgpu_ctx *c, int capturing)
{
    if (!c) return fail(AIRGPU_ERR_INVALID, "ctx is NULL");
    c->capturing = capturing != 0;
    return AIRGPU_OK;
}

int airgpu_graph_launch(airgpu_graph *g, void *stream)
{
    if (!g || !stream) return fail(AIRGPU_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(g->device));
    CU(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
    return AIRGPU_OK;
}

void airgpu_graph_destroy(airgpu_graph *g)
{
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
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
    int rc = decode_host_enqueue(c, iq, n_samples, segment_samples, base_offset, cap);
    if (rc != AIRGPU_OK) return rc;
    size_t n = 0;
    if ((rc = decode_host_finish(c, out, cap, &n)) != AIRGPU_OK) return rc;
    CU(cudaStreamSynchronize(c->compute));
    if (n_frames) *n_frames = n;
    if (n > cap) return fail(AIRGPU_ERR_OVERFLOW, "%zu frames but capacity %zu", n, cap);
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
    if (bytes) memcpy(s.h_in, iq, bytes);
    if (bytes) CU(cudaMemcpyAsync(s.d_in, s.h_in, bytes, cudaMemcpyHostToDevice, c->copy));
    CU(cudaEventRecord(s.copied, c->copy));
    CU(cudaStreamWaitEvent(c->compute, s.copied, 0));
    // one memset + three kernels + small copies per buffer: the count with the gate counter, and a fixed head of
    // records; a buffer with more frames than the head has the rest fetched by airgpu_collect
    int rc = enqueue_piece(c, s.d_in, n_samples, 0, base_offset, single_out(s.d_out), c->ring_cap, nullptr, true, c->compute);
    if (rc != AIRGPU_OK) return rc;
    CU(cudaMemcpyAsync(s.h_count, c->total(), sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->compute));
    CU(cudaMemcpyAsync(s.h_count + 1, c->counters() + kCounterGate, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->compute));
    CU(cudaMemcpyAsync(s.h_head, s.d_out, std::min(kHeadFrames, c->ring_cap) * sizeof(airgpu_frame), cudaMemcpyDeviceToHost, c->compute));
    CU(cudaEventRecord(s.done, c->compute));
    c->ev_valid = false;
    c->evh_valid = false;
    c->sync_valid = false;
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
    *n_frames = (size_t)n;
    if (n > cap)      // the ticket stays collectable: nothing is dropped, the caller comes back with room for n
        return fail(AIRGPU_ERR_OVERFLOW, "%llu frames but capacity %zu: collect ticket %llu again with a larger array", n, cap,
                    (unsigned long long)ticket);
    if (n && !out) return fail(AIRGPU_ERR_INVALID, "out is NULL");
    const size_t head = (size_t)std::min<unsigned long long>(n, kHeadFrames);
    if (head) memcpy(out, s.h_head, head * sizeof(airgpu_frame));
    if (n > head) {
        // rare: a buffer with more than kHeadFrames frames; its device array is untouched until the slot is reused
        CU(cudaMemcpyAsync(out + head, s.d_out + head, (size_t)(n - head) * sizeof(airgpu_frame), cudaMemcpyDeviceToHost, c->compute));
        CU(cudaStreamSynchronize(c->compute));
    }
    s.busy = false;
    c->next_collect++;
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

// ---------------------------------------------------------------------------
// multi-GPU from one host thread
// ---------------------------------------------------------------------------
struct airgpu_group {
    std::vector<airgpu_ctx *> ctx;
    std::vector<airgpu_stats> stats;
    int format = AIRGPU_FMT_CS16;
};

namespace {

// Candidate-range boundaries b[0..world]: shard r owns candidates [b[r], b[r+1]) and reads samples
// [b[r], b[r+1] + 240).  Boundaries sit on tile multiples so that every shard keeps 16-byte aligned loads.
std::vector<size_t> shard_bounds(size_t n_samples, unsigned world)
{
    const size_t align = 16384;
    const size_t cands = n_samples > (size_t)kFrameSamples ? n_samples - kFrameSamples : 0;
    std::vector<size_t> b(world + 1);
    for (unsigned r = 0; r < world; ++r) {
        const unsigned long long cut = (unsigned long long)cands * r / world;
        b[r] = std::min<size_t>(cands, (size_t)(cut / align * align));
    }
    b[world] = cands;
    return b;
}

}  // namespace

extern "C" {

void airgpu_group_destroy(airgpu_group *g)
{
    if (!g) return;
    for (airgpu_ctx *c : g->ctx) destroy_ctx(c);
    delete g;
}

int airgpu_group_create(const int *devices, uint32_t n_devices, uint32_t format, airgpu_group **out)
{
    if (!out) return fail(AIRGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return fail(AIRGPU_ERR_INVALID, "1..64 devices, got %u", n_devices);
    airgpu_group *g = new (std::nothrow) airgpu_group();
    if (!g) return fail(AIRGPU_ERR_NOMEM, "out of host memory");
    g->format = (int)format;
    for (uint32_t k = 0; k < n_devices; ++k) {
        airgpu_config cfg{};
        cfg.struct_size = sizeof cfg;
        cfg.device = devices[k];
        cfg.format = format;
        cfg.ring_slots = 1;                  // the group only uses the one-shot path
        cfg.max_buffer_samples = 1024;
        cfg.max_frames = 64;
        airgpu_ctx *c = nullptr;
        const int rc = airgpu_create(&cfg, &c);
        if (rc != AIRGPU_OK) {
            airgpu_group_destroy(g);
            return rc;
        }
        g->ctx.push_back(c);
    }
    g->stats.resize(n_devices);
    *out = g;
    return AIRGPU_OK;
}

int airgpu_group_decode(airgpu_group *g, const void *iq, size_t n_samples, uint64_t base_offset, airgpu_frame *out, size_t cap,
                        size_t *n_frames)
{
    if (!g) return fail(AIRGPU_ERR_INVALID, "group is NULL");
    if (n_frames) *n_frames = 0;
    if ((n_samples && !iq) || (cap && !out)) return fail(AIRGPU_ERR_INVALID, "NULL buffer");
    const unsigned world = (unsigned)g->ctx.size();
    const size_t bps = bytes_per_sample(g->format);
    const std::vector<size_t> b = shard_bounds(n_samples, world);
    // queue every shard (H2D chunk pipeline + kernels) on its device before waiting for any of them
    for (unsigned r = 0; r < world; ++r) {
        const size_t first = b[r], n = b[r + 1] > b[r] ? b[r + 1] - b[r] + kFrameSamples : 0;
        const int rc = decode_host_enqueue(g->ctx[r], static_cast<const char *>(iq) + first * bps, n, 0, base_offset + first, cap);
        if (rc != AIRGPU_OK) return rc;
    }
    // counts in rank order: rank order == ascending offset == the reference's order (adsb.rs:98)
    size_t total = 0;
    for (unsigned r = 0; r < world; ++r) {
        size_t n = 0;
        const size_t room = cap > total ? cap - total : 0;
        const int rc = decode_host_finish(g->ctx[r], out ? out + std::min(total, cap) : nullptr, room, &n);
        if (rc != AIRGPU_OK) return rc;
        total += n;
    }
    for (unsigned r = 0; r < world; ++r) {
        CU(cudaSetDevice(g->ctx[r]->device));
        CU(cudaStreamSynchronize(g->ctx[r]->compute));
        airgpu_get_stats(g->ctx[r], &g->stats[r]);
    }
    if (n_frames) *n_frames = total;
    if (total > cap) return fail(AIRGPU_ERR_OVERFLOW, "%zu frames but capacity %zu", total, cap);
    return AIRGPU_OK;
}

int airgpu_group_stats(airgpu_group *g, airgpu_stats *stats, uint32_t n_stats)
{
    if (!g || !stats) return fail(AIRGPU_ERR_INVALID, "NULL argument");
    for (uint32_t k = 0; k < n_stats && k < g->stats.size(); ++k) stats[k] = g->stats[k];
    return AIRGPU_OK;
}

int airgpu_decode_sharded(const int *devices, uint32_t n_devices, uint32_t format, const void *iq, size_t n_samples,
                          uint64_t base_offset, airgpu_frame *out, size_t cap, size_t *n_frames)
{
    airgpu_group *g = nullptr;
    int rc = airgpu_group_create(devices, n_devices, format, &g);
    if (rc != AIRGPU_OK) return rc;
    rc = airgpu_group_decode(g, iq, n_samples, base_offset, out, cap, n_frames);
    airgpu_group_destroy(g);
    return rc;
}

}  // extern "C"
