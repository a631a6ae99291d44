/* A host in plain C that reads a Curvine file into GPU memory, CRC-verified, through nothing but include/curvine_b200*.h and
 * libcurvine_b200.so -- no CUDA headers, no C++.  It is what a cgo / JNI / Rust-FFI binding of the reference would call
 * (INTEGRATION.md); tests/test_host.py compiles and links it with gcc, the GPU suite runs it.
 *
 *   c_host <cluster.toml> <path> [device-read bytes per call]
 * prints: bytes read, sum of the per-block CRCs, mismatching blocks, and the first 16 bytes (copied back through a pinned buffer). */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>

#include "curvine_b200.h"
#include "curvine_b200_kernels.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int64_t rc_ = (call);                                                              \
        if (rc_ != 0) {                                                                    \
            fprintf(stderr, "%s failed: %" PRId64 " (%s)\n", #call, rc_, cv_last_error()); \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <cluster.toml> <path> [bytes per device read]\n", argv[0]);
        return 2;
    }
    const int64_t step = argc > 3 ? atoll(argv[3]) : (int64_t)64 << 20;
    cv_fs* fs = NULL;
    cv_reader* r = NULL;
    int64_t len = 0;
    CHECK(cv_fs_new(argv[1], &fs));
    CHECK(cv_open(fs, argv[2], &r, &len));

    void *d_dst = NULL, *h_head = NULL;
    cv_stream_t stream = NULL;
    CHECK(cvh_device_alloc((size_t)len + 1, &d_dst));
    CHECK(cvh_pinned_alloc(16, &h_head));
    CHECK(cvh_stream_create(&stream));

    int64_t total = 0;
    for (;;) { /* sequential device reads, each ordered on `stream`; 0 bytes = end of file (not an error) */
        int64_t got = 0;
        CHECK(cv_read_device(r, (uint8_t*)d_dst + total, step < len - total ? step : len - total, stream, &got));
        if (got == 0) break;
        total += got;
    }
    uint64_t sum_crc = 0, n_verified = 0;
    uint32_t n_bad = 0;
    CHECK(cv_verify(r, &sum_crc, &n_bad, &n_verified)); /* blocks until the CRCs are back */
    const size_t head = total < 16 ? (size_t)total : 16;
    CHECK(cvh_d2h_async(h_head, d_dst, head, stream, NULL));
    CHECK(cvh_stream_synchronize(stream));

    printf("bytes %" PRId64 " of %" PRId64 " sum_crc %" PRIu64 " verified %" PRIu64 " bad %u head", total, len, sum_crc, n_verified, n_bad);
    for (size_t i = 0; i < head; i++) printf(" %02x", ((const uint8_t*)h_head)[i]);
    printf("\n");

    CHECK(cv_close_reader(r));
    CHECK(cv_fs_close(fs));
    CHECK(cvh_stream_destroy(stream));
    CHECK(cvh_pinned_free(h_head));
    CHECK(cvh_device_free(d_dst));
    return n_bad ? 3 : 0;
}
