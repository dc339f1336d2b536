/* c_consumer.c -- a plain C99 program against include/airgpu.h and libairgpu.so: the boundary is usable from C
 * (and therefore from Rust's extern "C", cgo, JNI ...) without any C++ or CUDA type.  Built and run by
 * tests/test_abi.py::test_c_consumer_links_and_runs.  Without a GPU it checks the error path (no CPU fallback);
 * with one it decodes a constant buffer (a frame at every offset: ties pass the gate, crc(0) = 0). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "airgpu.h"

int main(void)
{
    printf("%s, abi %d, %d device(s)\n", airgpu_version(), AIRGPU_ABI_VERSION, airgpu_device_count());
    if (airgpu_playback_samples(100000, 20000) != 80000 || airgpu_playback_samples(20000, 20000) != 0) return 2;
    airgpu_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.format = AIRGPU_FMT_CS16;
    airgpu_ctx *ctx = NULL;
    int rc = airgpu_create(&cfg, &ctx);
    if (airgpu_device_count() == 0) {
        if (rc != AIRGPU_ERR_NO_DEVICE || ctx != NULL) return 3;
        printf("no device: %s\n", airgpu_last_error());
        return 0;
    }
    if (rc != AIRGPU_OK) {
        printf("create failed: %s\n", airgpu_last_error());
        return 4;
    }
    enum { N = 1000 };
    int16_t *iq = (int16_t *)calloc(2 * N, sizeof *iq);
    airgpu_frame *out = (airgpu_frame *)calloc(N, sizeof *out);
    size_t n = 0;
    rc = airgpu_decode(ctx, iq, N, 0, 5, out, N, &n);
    printf("decode rc %d, %zu frames, first offset %llu\n", rc, n, n ? (unsigned long long)out[0].offset : 0ull);
    rc = (rc == AIRGPU_OK && n == N - 240 && out[0].offset == 5 && out[n - 1].offset == 5 + N - 241 && out[0].fixed_bit == 0xFF) ? 0 : 5;
    airgpu_destroy(ctx);
    free(iq);
    free(out);
    return rc;
}
