/*
 * airgpu_synth.h -- device-side synthetic capture generator (workload only).
 *
 * Not part of the reference-facing boundary: the reference has no generator
 * (its author's capture is git-ignored, reference .gitignore:4).  This is the
 * device twin of air_rs_b200/synth.py, exported from the same
 * shared library so benchmarks can fill HBM without crossing PCIe.
 */
#ifndef AIRGPU_SYNTH_H
#define AIRGPU_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct airgpu_synth_table airgpu_synth_table;

const char *airgpu_synth_last_error(void);

/* Upload a table of injected transmissions, sorted by `start`.
 * payload: n_frames x 14 bytes, MSB-first bits; nbits: 112 or 56. */
int airgpu_synth_table_create(int device, const int64_t *start, const int32_t *nbits,
                              const uint8_t *payload, const int32_t *amp_i, const int32_t *amp_q,
                              const uint8_t *smear, size_t n_frames, airgpu_synth_table **out);
void airgpu_synth_table_destroy(airgpu_synth_table *t);

/* Render samples [j0, j0+n) as interleaved u8 (format 1) or i16 (format 0) into
 * device memory.  noise_gain = round(sigma * 65536 / sqrt(43690)); period > 0
 * repeats the frame schedule every `period` samples.  stream: cudaStream_t. */
int airgpu_synth_render(airgpu_synth_table *t, uint64_t seed, uint64_t j0, uint64_t n,
                        uint32_t format, int32_t noise_gain, uint64_t period, void *d_out,
                        void *stream);

#ifdef __cplusplus
}
#endif
#endif
