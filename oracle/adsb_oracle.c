/*
 * adsb_oracle.c -- CPU restatement of air_rs's ADS-B decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see adsb_oracle.h).  Plain C11 + pthreads.
 *
 * Two implementations of the same function live here on purpose:
 *   - the LITERAL one follows the reference statement by statement (f64 sqrt,
 *     32-word window copy, early-exit double loops, bit-vector long division,
 *     brute-force 112-flip repair).  It is the ground truth.
 *   - the FAST one is derived independently (integer isqrt, min/max gate,
 *     table CRC, syndrome lookup) and is only trusted because the tests show
 *     it emits the identical frame list.  It exists so that full-size captures
 *     can be checked in seconds and as the "optimised CPU" baseline.
 * Parity is "unpinned" end to end (no IQ fixture upstream) -- see the header.
 */
#define _GNU_SOURCE
#include "adsb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

/* ========================================================================= */
/* Literal restatement                                                       */
/* ========================================================================= */

/* utils.rs:46-52 */
void oracle_get_magnitude(const int16_t *iq, size_t n, uint32_t *mags)
{
    for (size_t k = 0; k < n; ++k) {
        double re = (double)iq[2 * k];
        double im = (double)iq[2 * k + 1];
        /* powi(2) is re*re; `as u32` truncates toward zero (value is >= 0). */
        mags[k] = (uint32_t)sqrt(re * re + im * im);
    }
}

/* U8 widening, centred on 127.5, exact integers in +-32640. */
void oracle_widen_u8(const uint8_t *iq, size_t n, int16_t *out)
{
    for (size_t k = 0; k < 2 * n; ++k)
        out[k] = (int16_t)((2 * (int)iq[k] - 255) * 128);
}

/* demod.rs:17-57 */
int oracle_check_for_adsb_packet(const uint32_t buf[32], uint32_t *high)
{
    static const int pre_lows[12] = {1, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 15};
    static const int pre_highs[4] = {0, 2, 7, 9};
    uint32_t min = UINT32_MAX;

    for (int h = 0; h < 4; ++h) {                      /* demod.rs:27-36 */
        for (int l = 0; l < 12; ++l)
            if (buf[pre_highs[h]] < buf[pre_lows[l]])
                return 0;
        if (buf[pre_highs[h]] < min)
            min = buf[pre_highs[h]];
    }

    static const int df_lows[5] = {1, 2, 4, 6, 9};      /* demod.rs:45-46 */
    static const int df_highs[5] = {0, 3, 5, 7, 8};
    for (int h = 0; h < 5; ++h)                         /* demod.rs:48-54 */
        for (int l = 0; l < 5; ++l)
            if (buf[df_highs[h] + 16] < buf[df_lows[l] + 16])
                return 0;

    if (high)
        *high = (uint32_t)((float)min * 0.9f);          /* demod.rs:56 (dead value) */
    return 1;
}

/* demod.rs:92-131 */
int oracle_extract_manchester_relative(const uint32_t *buf, size_t len, uint32_t high,
                                       uint16_t *symbols)
{
    (void)high;                                         /* `_high` is unused upstream */
    int errors = 0;
    size_t n_out = 0;
    for (size_t block_start = 0; block_start < len; block_start += 16) {
        uint16_t symbol = 0;
        for (int bit = 0; bit < 8; ++bit) {
            size_t i = block_start + (size_t)bit * 2;
            int first, second;
            if (buf[i] > buf[i + 1]) { first = 1; second = 0; }
            else                     { first = 0; second = 1; }
            if (first != second) {
                symbol |= (uint16_t)(first << (14 - bit * 2));
                symbol |= (uint16_t)(second << (15 - bit * 2));
            } else {                                    /* unreachable upstream too */
                errors += 1;
                if (errors > 2)
                    return 0;
            }
        }
        symbols[n_out++] = symbol;
        errors = 0;
    }
    return 1;
}

/* demod.rs:180-201 */
size_t oracle_decode_packet(const uint16_t *symbols, size_t n, uint8_t *bytes)
{
    for (size_t s = 0; s < n; ++s) {
        uint16_t encoded = symbols[s];
        uint8_t byte = 0;
        for (int i = 0; i < 8; ++i) {
            int hi = (encoded >> (15 - i * 2)) & 1;
            int lo = (encoded >> (14 - i * 2)) & 1;
            if (hi == 0 && lo == 1)
                byte |= (uint8_t)(1 << (7 - i));
            /* (1,0) and the invalid pairs leave the bit clear */
        }
        bytes[s] = byte;
    }
    return n;
}

/* crc.rs:10-40 -- message * x^24 mod 0x1FFF409 by long division on a bit vector */
uint32_t oracle_get_adsb_crc(const uint8_t *buf, size_t len)
{
    const uint32_t GENERATOR = 0x1FFF409u;
    enum { GENERATOR_LEN = 24 };
    size_t nbits = len * 8 + GENERATOR_LEN;
    unsigned char stack_bits[14 * 8 + GENERATOR_LEN];
    unsigned char *bits = nbits <= sizeof stack_bits ? stack_bits : malloc(nbits);

    size_t w = 0;
    for (size_t b = 0; b < len; ++b)
        for (int i = 7; i >= 0; --i)
            bits[w++] = (unsigned char)((buf[b] >> i) & 1);
    for (int i = 0; i < GENERATOR_LEN; ++i)
        bits[w++] = 0;

    for (size_t i = 0; i < nbits - GENERATOR_LEN; ++i)
        if (bits[i])
            for (int j = 0; j <= GENERATOR_LEN; ++j)
                bits[i + (size_t)j] ^= (unsigned char)((GENERATOR >> (GENERATOR_LEN - j)) & 1);

    uint32_t remainder = 0;
    for (int i = 0; i < GENERATOR_LEN; ++i)
        if (bits[nbits - GENERATOR_LEN + (size_t)i])
            remainder |= 1u << (GENERATOR_LEN - 1 - i);

    if (bits != stack_bits)
        free(bits);
    return remainder;
}

/* crc.rs:49-65 */
int oracle_try_crc_recovery(uint8_t *buf, size_t len, uint32_t calc_crc,
                            uint32_t packet_crc, int *flipped)
{
    (void)calc_crc;                                     /* `_calc_crc` unused upstream */
    uint8_t augmented[64];
    if (len > sizeof augmented)
        return 0;
    for (size_t num = 0; num < len; ++num) {
        memcpy(augmented, buf, len);
        for (int i = 0; i < 8; ++i) {
            augmented[num] = (uint8_t)(buf[num] ^ (1u << (7 - i)));
            uint32_t crc = oracle_get_adsb_crc(augmented, len - 3);
            if (crc == packet_crc) {
                memcpy(buf, augmented, len);
                if (flipped)
                    *flipped = (int)(num * 8 + (size_t)i);
                return 1;
            }
        }
    }
    return 0;
}

/* demod.rs:65-82 */
int oracle_extract_packet(const uint32_t *buf, size_t len, uint32_t high,
                          uint8_t out[14], int *fixed_bit)
{
    uint16_t symbols[14];
    uint8_t packet[14];
    if (len != 224)
        return 0;
    if (!oracle_extract_manchester_relative(buf, len, (uint32_t)((double)high * 0.9), symbols))
        return 0;
    size_t plen = oracle_decode_packet(symbols, 14, packet);

    uint32_t calced_crc = oracle_get_adsb_crc(packet, plen - 3);
    uint32_t packet_crc = ((uint32_t)packet[plen - 1] << 0) |
                          ((uint32_t)packet[plen - 2] << 8) |
                          ((uint32_t)packet[plen - 3] << 16);
    int flipped = 0xFF;
    if (calced_crc != packet_crc) {
        if (!oracle_try_crc_recovery(packet, plen, calced_crc, packet_crc, &flipped))
            return 0;
    }
    memcpy(out, packet, 14);
    if (fixed_bit)
        *fixed_bit = flipped;
    return 1;
}

/* adsb.rs:96-116, one buffer */
size_t oracle_process_mags(const uint32_t *mags, size_t len, uint64_t base,
                           oracle_frame *out, size_t cap, uint64_t *gate_passes)
{
    size_t emitted = 0;
    uint64_t processed = 0;
    if (len >= ORACLE_FRAME_SAMPLES) {                 /* upstream panics below 240 */
        for (size_t i = 0; i < len - ORACLE_FRAME_SAMPLES; ++i) {
            uint32_t check_mags[32];
            memcpy(check_mags, mags + i, sizeof check_mags);   /* adsb.rs:99-101 */
            uint32_t high;
            if (oracle_check_for_adsb_packet(check_mags, &high)) {
                processed += 1;
                uint8_t pkt[14];
                int fixed;
                if (oracle_extract_packet(mags + i + 16, 224, high, pkt, &fixed)) {
                    if (emitted < cap) {
                        oracle_frame *f = &out[emitted];
                        memcpy(f->bytes, pkt, 14);
                        f->fixed_bit = (uint8_t)fixed;
                        f->reserved = 0;
                        f->offset = base + i;
                    }
                    emitted += 1;
                    /* adsb.rs:113 `_i += 240` does not affect a Rust range loop */
                }
            }
        }
    }
    if (gate_passes)
        *gate_passes = processed;
    return emitted;
}

/* ========================================================================= */
/* Fast restatement (independent derivation)                                 */
/* ========================================================================= */

static uint32_t isqrt_u32(uint32_t n)                  /* bit-by-bit, no floating point */
{
    uint32_t res = 0, bit = 1u << 30;
    while (bit > n) bit >>= 2;
    while (bit) {
        if (n >= res + bit) { n -= res + bit; res = (res >> 1) + bit; }
        else                  res >>= 1;
        bit >>= 2;
    }
    return res;
}

static uint32_t g_crc_table[256];
static uint32_t g_syndrome[88];
static uint16_t g_u8_lut[128 * 128];                   /* |2u-255|>>1 pairs -> magnitude */
static pthread_once_t g_tables_once = PTHREAD_ONCE_INIT;

static inline uint32_t crc24_table(const uint8_t *p, size_t n)
{
    uint32_t crc = 0;
    for (size_t i = 0; i < n; ++i)
        crc = ((crc << 8) ^ g_crc_table[((crc >> 16) ^ p[i]) & 0xFF]) & 0xFFFFFFu;
    return crc;
}

static void build_tables(void)
{
    for (uint32_t b = 0; b < 256; ++b) {
        uint32_t r = b << 16;
        for (int k = 0; k < 8; ++k)
            r = (r & 0x800000u) ? ((r << 1) ^ 0xFFF409u) & 0xFFFFFFu : (r << 1) & 0xFFFFFFu;
        g_crc_table[b] = r;
    }
    for (int p = 0; p < 88; ++p) {
        uint8_t e[11] = {0};
        e[p >> 3] = (uint8_t)(0x80u >> (p & 7));
        g_syndrome[p] = crc24_table(e, 11);
    }
    for (uint32_t a = 0; a < 128; ++a)
        for (uint32_t b = 0; b < 128; ++b) {
            uint32_t re = (2 * a + 1) * 128, im = (2 * b + 1) * 128;
            g_u8_lut[a * 128 + b] = (uint16_t)isqrt_u32(re * re + im * im);
        }
}

void oracle_syndrome_table(uint32_t table[88])
{
    pthread_once(&g_tables_once, build_tables);
    memcpy(table, g_syndrome, sizeof g_syndrome);
}

static inline uint32_t fold_u8(uint32_t u) { return u >= 128 ? u - 128 : 127 - u; }

static void fast_mags(const void *iq, int format, size_t first, size_t count, uint16_t *m)
{
    if (format == ORACLE_FMT_U8) {
        const uint8_t *p = (const uint8_t *)iq + 2 * first;
        for (size_t k = 0; k < count; ++k)
            m[k] = g_u8_lut[fold_u8(p[2 * k]) * 128 + fold_u8(p[2 * k + 1])];
    } else {
        const int16_t *p = (const int16_t *)iq + 2 * first;
        for (size_t k = 0; k < count; ++k) {
            int32_t re = p[2 * k], im = p[2 * k + 1];
            uint32_t n = (uint32_t)(re * re) + (uint32_t)(im * im);
            uint32_t r = (uint32_t)sqrtf((float)n);
            while ((uint64_t)r * r > n) --r;
            while ((uint64_t)(r + 1) * (r + 1) <= n) ++r;
            m[k] = (uint16_t)r;
        }
    }
}

static inline uint16_t min16(uint16_t a, uint16_t b) { return a < b ? a : b; }
static inline uint16_t max16(uint16_t a, uint16_t b) { return a > b ? a : b; }

typedef struct {
    oracle_frame *v;
    size_t n, cap;
    uint64_t gate_passes;
} frame_vec;

static void vec_push(frame_vec *fv, const uint8_t pkt[14], int fixed, uint64_t off)
{
    if (fv->n == fv->cap) {
        fv->cap = fv->cap ? fv->cap * 2 : 64;
        fv->v = realloc(fv->v, fv->cap * sizeof *fv->v);
    }
    oracle_frame *f = &fv->v[fv->n++];
    memcpy(f->bytes, pkt, 14);
    f->fixed_bit = (uint8_t)fixed;
    f->reserved = 0;
    f->offset = off;
}

/* candidates [0, cands) over m[0 .. cands+239]; frame offset = off0 + i */
static void fast_scan(const uint16_t *mags, size_t cands, uint64_t off0, frame_vec *fv)
{
    for (size_t i = 0; i < cands; ++i) {
        const uint16_t *m = mags + i;
        uint16_t hmin = min16(min16(m[0], m[2]), min16(m[7], m[9]));
        uint16_t l = max16(m[1], m[3]);
        if (hmin < l) continue;
        l = max16(max16(max16(m[4], m[5]), max16(m[6], m[8])), l);
        l = max16(max16(max16(m[10], m[11]), max16(m[12], m[13])), max16(max16(m[14], m[15]), l));
        if (hmin < l) continue;
        uint16_t dh = min16(min16(min16(m[16], m[19]), min16(m[21], m[23])), m[24]);
        uint16_t dl = max16(max16(max16(m[17], m[18]), max16(m[20], m[22])), m[25]);
        if (dh < dl) continue;
        fv->gate_passes += 1;

        uint8_t pkt[14];
        const uint16_t *d = m + 16;
        for (int b = 0; b < 14; ++b) {
            unsigned v = 0;
            for (int k = 0; k < 8; ++k)
                v = (v << 1) | (unsigned)(d[16 * b + 2 * k] > d[16 * b + 2 * k + 1]);
            pkt[b] = (uint8_t)v;
        }
        uint32_t rx = ((uint32_t)pkt[11] << 16) | ((uint32_t)pkt[12] << 8) | pkt[13];
        uint32_t syn = crc24_table(pkt, 11) ^ rx;
        int fixed = 0xFF;
        if (syn) {
            int p;
            for (p = 0; p < 88; ++p)
                if (g_syndrome[p] == syn) break;
            if (p == 88) continue;
            pkt[p >> 3] ^= (uint8_t)(0x80u >> (p & 7));
            fixed = p;
        }
        vec_push(fv, pkt, fixed, off0 + i);
    }
}

/* ========================================================================= */
/* Work splitting shared by the fast and the threaded literal paths          */
/* ========================================================================= */

typedef struct {
    size_t first;       /* first sample (absolute index in the capture)          */
    size_t cands;       /* candidate offsets [first, first+cands)                 */
    uint64_t off0;      /* frame offset of candidate 0                             */
    frame_vec out;
} work_item;

typedef struct {
    const void *iq;
    int format;
    int literal;
    work_item *items;
    size_t n_items;
    atomic_size_t next;
} work_ctx;

static void run_item(const work_ctx *c, work_item *it)
{
    size_t span = it->cands + ORACLE_FRAME_SAMPLES;   /* samples the loop bound needs */
    if (c->literal) {
        uint32_t *mags = malloc(span * sizeof *mags);
        if (c->format == ORACLE_FMT_U8) {
            int16_t *wide = malloc(span * 2 * sizeof *wide);
            oracle_widen_u8((const uint8_t *)c->iq + 2 * it->first, span, wide);
            oracle_get_magnitude(wide, span, mags);
            free(wide);
        } else {
            oracle_get_magnitude((const int16_t *)c->iq + 2 * it->first, span, mags);
        }
        /* count first, then store: process_mags needs a sized buffer */
        size_t cap = 1024;
        for (;;) {
            oracle_frame *buf = malloc(cap * sizeof *buf);
            uint64_t gp = 0;
            size_t n = oracle_process_mags(mags, span, it->off0, buf, cap, &gp);
            if (n <= cap) {
                it->out.v = buf; it->out.n = n; it->out.cap = cap; it->out.gate_passes = gp;
                break;
            }
            free(buf);
            cap = n;
        }
        free(mags);
    } else {
        uint16_t *mags = malloc(span * sizeof *mags);
        fast_mags(c->iq, c->format, it->first, span, mags);
        fast_scan(mags, it->cands, it->off0, &it->out);
        free(mags);
    }
}

static void *worker(void *arg)
{
    work_ctx *c = arg;
    for (;;) {
        size_t k = atomic_fetch_add(&c->next, 1);
        if (k >= c->n_items) break;
        run_item(c, &c->items[k]);
    }
    return NULL;
}

static size_t decode_split(const void *iq, size_t n_samples, int format,
                           size_t segment_samples, uint64_t base,
                           oracle_frame *out, size_t cap, uint64_t *gate_passes,
                           int n_threads, int literal)
{
    pthread_once(&g_tables_once, build_tables);
    if (segment_samples == 0 || segment_samples > n_samples)
        segment_samples = n_samples;
    if (n_threads < 1) n_threads = 1;
    if (gate_passes) *gate_passes = 0;
    if (n_samples == 0) return 0;

    size_t n_seg = (n_samples + segment_samples - 1) / segment_samples;
    /* piece size: whole segments when single-threaded; otherwise ~16 pieces per thread */
    size_t piece = (size_t)-1;
    if (n_threads > 1) {
        piece = n_samples / ((size_t)n_threads * 16) + 1;
        if (piece < 4096) piece = 4096;
    }

    size_t n_items = 0, cap_items = 0;
    work_item *items = NULL;
    for (size_t s = 0; s < n_seg; ++s) {
        size_t seg0 = s * segment_samples;
        size_t len = n_samples - seg0 < segment_samples ? n_samples - seg0 : segment_samples;
        if (len <= ORACLE_FRAME_SAMPLES) continue;       /* no candidates (upstream: panic/none) */
        size_t cands = len - ORACLE_FRAME_SAMPLES;
        for (size_t a = 0; a < cands; ) {
            size_t c = cands - a < piece ? cands - a : piece;
            if (n_items == cap_items) {
                cap_items = cap_items ? cap_items * 2 : 64;
                items = realloc(items, cap_items * sizeof *items);
            }
            work_item *it = &items[n_items++];
            memset(it, 0, sizeof *it);
            it->first = seg0 + a;
            it->cands = c;
            it->off0 = base + seg0 + a;
            a += c;
        }
    }

    work_ctx ctx = {iq, format, literal, items, n_items, 0};
    if (n_threads == 1 || n_items <= 1) {
        for (size_t k = 0; k < n_items; ++k) run_item(&ctx, &items[k]);
    } else {
        pthread_t *th = malloc((size_t)n_threads * sizeof *th);
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, worker, &ctx);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        free(th);
    }

    size_t total = 0;
    uint64_t gp = 0;
    for (size_t k = 0; k < n_items; ++k) {
        for (size_t j = 0; j < items[k].out.n; ++j) {
            if (total < cap) out[total] = items[k].out.v[j];
            total += 1;
        }
        gp += items[k].out.gate_passes;
        free(items[k].out.v);
    }
    free(items);
    if (gate_passes) *gate_passes = gp;
    return total;
}

size_t oracle_decode_literal(const void *iq, size_t n_samples, int format,
                             size_t segment_samples, uint64_t base,
                             oracle_frame *out, size_t cap, uint64_t *gate_passes)
{
    return decode_split(iq, n_samples, format, segment_samples, base, out, cap, gate_passes, 1, 1);
}

size_t oracle_decode_literal_mt(const void *iq, size_t n_samples, int format,
                                size_t segment_samples, uint64_t base,
                                oracle_frame *out, size_t cap, uint64_t *gate_passes,
                                int n_threads)
{
    return decode_split(iq, n_samples, format, segment_samples, base, out, cap, gate_passes,
                        n_threads, 1);
}

size_t oracle_decode_fast(const void *iq, size_t n_samples, int format,
                          size_t segment_samples, uint64_t base,
                          oracle_frame *out, size_t cap, uint64_t *gate_passes,
                          int n_threads)
{
    return decode_split(iq, n_samples, format, segment_samples, base, out, cap, gate_passes,
                        n_threads, 0);
}

/* ========================================================================= */
/* N1: AdsbPacket::new field derivation (literal restatement)                */
/* ========================================================================= */

/* msgs.rs:141-162 */
static size_t to_6bit_chunks(const uint8_t *input, size_t n, uint8_t *out)
{
    uint32_t acc = 0;
    int bits = 0;
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        acc = (acc << 8) | input[i];
        bits += 8;
        while (bits >= 6) {
            bits -= 6;
            out[k++] = (uint8_t)((acc >> bits) & 0x3F);
        }
    }
    if (bits > 0)
        out[k++] = (uint8_t)((acc << (6 - bits)) & 0x3F);
    return k;
}

/* msgs.rs:164-169 */
static const char CHAR_CONVERT[65] =
    "#ABCDEFGHIJKLMNOPQRSTUVWXYZ#####_###############0123456789######";

void oracle_packet_fields(const uint8_t packet[14], oracle_fields *o)
{
    memset(o, 0, sizeof *o);
    o->downlink_format = packet[0] >> 3;                                  /* packet.rs:26 */
    o->capability = packet[0] & 5;                                        /* packet.rs:27 (sic) */
    o->icao = ((uint32_t)packet[1] << 16) | ((uint32_t)packet[2] << 8) | packet[3];
    o->msg_type = packet[4] >> 3;                                         /* packet.rs:29 */
    const uint8_t *msg = packet + 4;                                      /* packet[4..4+7] */
    if (1 <= o->msg_type && o->msg_type <= 4) {                           /* msgs.rs:208-213 */
        uint8_t six[16];
        size_t n = to_6bit_chunks(msg + 1, 6, six);                       /* msgs.rs:173 */
        o->kind = 1;
        for (size_t k = 0; k < n && k < 8; ++k)
            o->callsign[k] = six[k] < 64 ? CHAR_CONVERT[six[k]] : '?';
    } else if (9 <= o->msg_type && o->msg_type <= 18) {                   /* msgs.rs:121-125 */
        o->kind = 2;
        int alt_mode_25 = (msg[1] & (1 << 0)) == 1;                       /* msgs.rs:70 */
        int32_t altitude = ((int32_t)((msg[1] & 0xFE) >> 1) << 4) | ((int32_t)(msg[2] & 0xF0) >> 4);
        altitude *= alt_mode_25 ? 25 : 100;
        altitude -= 1000;
        o->altitude = altitude;
        o->surveillance_status = (msg[0] & 0x06) >> 1;
        o->nic_supplement = msg[0] & 0x01;
        o->cpr_time = (msg[2] & 0x08) >> 3;
        o->cpr_odd = ((msg[2] & 0x04) >> 2) == 1;
        o->cpr_latitude = ((uint32_t)(msg[2] & 0x03) << 15) | ((uint32_t)msg[3] << 7) | (((uint32_t)msg[4] & 0xFE) >> 1);
        o->cpr_longitude = ((uint32_t)(msg[4] & 0x01) << 16) | ((uint32_t)msg[5] << 8) | (uint32_t)msg[6];
    }
}

void oracle_frames_fields(const oracle_frame *frames, size_t n, oracle_fields *out)
{
    for (size_t k = 0; k < n; ++k)
        oracle_packet_fields(frames[k].bytes, &out[k]);
}
