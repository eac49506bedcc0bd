/* Plain-C client of libvqb200 (no Python, no torch): proves that include/vqb200.h is a self-contained
 * C ABI.  Without a GPU it checks the size queries and error paths; with a GPU (argv[1] == "gpu") it
 * runs codebook prepare -> search -> gather/loss on a tiny problem and compares with a CPU loop that
 * follows vqgan_ldm_baseline/models/quantizer.py:68-98.
 *
 *   gcc -std=c99 -I include tests/c_abi/c_abi_smoke.c -o c_abi_smoke \
 *       -L vq_gan_b200/lib -lvqb200 -L /usr/local/cuda/lib64 -lcudart -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vqb200.h"

/* the five CUDA runtime calls this client needs, declared by hand to stay plain C99 */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);
extern int cudaGetDeviceCount(int* n);

#define CHECK(x)                                                                        \
    do {                                                                                \
        int rc_ = (x);                                                                  \
        if (rc_ != 0) {                                                                 \
            fprintf(stderr, "%s failed: %d (%s)\n", #x, rc_, vqb_last_error());         \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

static float frand(unsigned* s) {
    *s = *s * 1664525u + 1013904223u;
    return ((float)(*s >> 8) / 8388608.0f) - 1.0f;
}

int main(int argc, char** argv) {
    if (vqb_version() < 100) return 2;
    if (vqb_codebook_pack_bytes(16384, 4) == 0 || vqb_codebook_pack_bytes(0, 4) != 0) return 3;
    if (vqb_search_workspace_bytes(4, 256, 1024, 16384, 0) == 0) return 4;
    if (vqb_index_bytes(128) != 1 || vqb_index_bytes(65536) != 2 || vqb_index_bytes(65537) != 4) return 5;
    if (vqb_search_f32(NULL, -1, 4, 1, NULL, 4, NULL, NULL, NULL, NULL, 0, 0, NULL, NULL) == 0) return 6;
    if (strlen(vqb_last_error()) == 0) return 7;
    printf("c_abi_smoke: host checks OK (version %d)\n", vqb_version());
    if (argc < 2 || strcmp(argv[1], "gpu") != 0) return 0;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != 0 || ndev == 0) {
        fprintf(stderr, "no CUDA device\n");
        return 8;
    }
    enum { B = 2, D = 4, HW = 96, K = 300, N = B * HW };
    static float z[B * D * HW], E[K * D], zq[B * D * HW], loss[2];
    static int64_t idx[N];
    unsigned seed = 12345;
    for (int i = 0; i < B * D * HW; ++i) z[i] = frand(&seed);
    for (int i = 0; i < K * D; ++i) E[i] = frand(&seed);

    void *dz, *dE, *dpack, *didx, *dzq, *dloss, *dws, *dpart;
    const size_t pack_bytes = vqb_codebook_pack_bytes(K, D);
    const size_t ws_bytes = vqb_search_workspace_bytes(B, D, HW, K, 0);
    const size_t part_bytes = vqb_tail_partials_bytes(N);
    CHECK(cudaMalloc(&dz, sizeof z));
    CHECK(cudaMalloc(&dE, sizeof E));
    CHECK(cudaMalloc(&dpack, pack_bytes));
    CHECK(cudaMalloc(&didx, sizeof idx));
    CHECK(cudaMalloc(&dzq, sizeof zq));
    CHECK(cudaMalloc(&dloss, sizeof loss));
    CHECK(cudaMalloc(&dws, ws_bytes ? ws_bytes : 16));
    CHECK(cudaMalloc(&dpart, part_bytes));
    CHECK(cudaMemcpy(dz, z, sizeof z, 1));
    CHECK(cudaMemcpy(dE, E, sizeof E, 1));
    CHECK(vqb_codebook_prepare_f32((const float*)dE, K, D, dpack, pack_bytes, NULL));
    CHECK(vqb_search_f32((const float*)dz, B, D, HW, (const float*)dE, K, dpack, (int64_t*)didx, NULL, dws, ws_bytes,
                         0, NULL, NULL));
    CHECK(vqb_gather_loss_st_f32((const float*)dz, (const float*)dE, (const int64_t*)didx, B, D, HW, K, 0.25f,
                                 (float*)dzq, (float*)dloss, dpart, part_bytes, NULL, NULL));
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaMemcpy(idx, didx, sizeof idx, 2));
    CHECK(cudaMemcpy(zq, dzq, sizeof zq, 2));
    CHECK(cudaMemcpy(loss, dloss, sizeof loss, 2));

    /* CPU restatement of quantizer.py:68-98 in double precision */
    int bad = 0;
    double sq = 0.0;
    for (int b = 0; b < B; ++b)
        for (int hw = 0; hw < HW; ++hw) {
            int best = 0;
            double bd = 1e300;
            for (int k = 0; k < K; ++k) {
                double d = 0.0;
                for (int c = 0; c < D; ++c) {
                    const double t = (double)z[(b * D + c) * HW + hw] - (double)E[k * D + c];
                    d += t * t;
                }
                if (d < bd) {
                    bd = d;
                    best = k;
                }
            }
            const int64_t got = idx[b * HW + hw];
            if (got != best) {
                double dg = 0.0;
                for (int c = 0; c < D; ++c) {
                    const double t = (double)z[(b * D + c) * HW + hw] - (double)E[got * D + c];
                    dg += t * t;
                }
                if (fabs(dg - bd) > 1e-6) ++bad; /* only float32-level ties may differ */
            }
            for (int c = 0; c < D; ++c) {
                const float zv = z[(b * D + c) * HW + hw], ev = E[got * D + c];
                const float want = zv + (ev - zv); /* straight-through value, quantizer.py:98 */
                if (zq[(b * D + c) * HW + hw] != want) ++bad;
                sq += ((double)ev - zv) * ((double)ev - zv);
            }
        }
    const double mse = sq / (B * D * HW);
    if (fabs(loss[0] - mse) > 1e-6 * mse || fabs(loss[1] - 1.25 * mse) > 1e-6 * mse) ++bad;
    printf("c_abi_smoke: gpu run, mismatches %d, mse %.8f (device %.8f), vq_loss %.8f\n", bad, mse, loss[0], loss[1]);
    cudaFree(dz); cudaFree(dE); cudaFree(dpack); cudaFree(didx); cudaFree(dzq); cudaFree(dloss); cudaFree(dws); cudaFree(dpart);
    return bad ? 9 : 0;
}
