/* Plain-C consumer of include/pips_b200.h: proves the header is C (not only C++), that the library links
 * from C without torch, and that argument validation answers before any device is touched.
 *
 *   gcc -std=c99 -Wall -Wextra -Werror -Iinclude examples/abi_check.c -Lpips_b200/lib -lpips_b200 \
 *       -Wl,-rpath,$PWD/pips_b200/lib -o /tmp/abi_check && /tmp/abi_check
 *
 * (tests/test_abi_cpu.py builds and runs it; no GPU needed: every call below must fail in validation.) */
#include <stdio.h>
#include <string.h>

#include "pips_b200.h"

static int expect_error(const char* what, int rc, const char* needle) {
    const char* msg = pips_last_error();
    if (rc == 0 || !msg || !strstr(msg, needle)) {
        printf("FAIL %s: rc=%d msg=%s\n", what, rc, msg ? msg : "(null)");
        return 1;
    }
    printf("ok   %s -> \"%s\"\n", what, msg);
    return 0;
}

int main(void) {
    int bad = 0;
    if (pips_abi_version() != PIPS_B200_ABI_VERSION) {
        printf("FAIL abi version %d != header %d\n", pips_abi_version(), PIPS_B200_ABI_VERSION);
        return 1;
    }
    printf("ok   abi version %d; sizeof(pips_weights)=%zu sizeof(pips_workspace)=%zu sizeof(pips_problem)=%zu\n",
           pips_abi_version(), sizeof(pips_weights), sizeof(pips_workspace), sizeof(pips_problem));

    /* a problem description as a C caller would fill it; pointers stay NULL -> validation must refuse */
    pips_problem p;
    pips_weights w;
    pips_workspace ws;
    memset(&p, 0, sizeof p);
    memset(&w, 0, sizeof w);
    memset(&ws, 0, sizeof ws);
    p.B = 1; p.S = PIPS_S; p.N = 16; p.H = 48; p.W = 64;
    p.feat_dtype = PIPS_FEAT_F32; p.precision = PIPS_PREC_BF16X3; p.stride = 8.0f;
    bad += expect_error("pips_refine_iter(null buffers)", pips_refine_iter(&p, &w, &ws, NULL, NULL), "pips_");
    bad += expect_error("pips_gemm_tc(K % 64)", pips_gemm_tc(NULL, NULL, 512, 128, NULL, NULL, 512, 256, 128, 256, 100, NULL,
                                                             PIPS_EPI_BIAS, NULL, 0, NULL, NULL, 0, NULL), "multiple of 64");
    bad += expect_error("pips_pyramid_build(null)", pips_pyramid_build(NULL, 8, 48, 64, NULL, NULL, NULL), "pips_pyramid");
    bad += expect_error("pips_heatmap(null)", pips_heatmap(NULL, 0, 1, 8, 4, 16, 16, NULL, NULL, NULL, 1, NULL, NULL, 0, NULL),
                        "pips_heatmap");
    bad += expect_error("pips_peer_barrier(no timeout)", pips_peer_barrier((int* const*)&p, 0, 1, 1, 0, NULL), "timeout");
    if (pips_heatmap_scratch_floats(32, 3, 48, 64) != (size_t)32 * 3 * (48 * 64 + 24 * 32 + 12 * 16 + 6 * 8)) {
        printf("FAIL pips_heatmap_scratch_floats\n");
        ++bad;
    }
    printf(bad ? "FAILED (%d)\n" : "all checks passed%.0d\n", bad);
    return bad ? 1 : 0;
}
