/* A consumer of liblsmb200.so written in plain C: no Python, no torch.  tests/test_gpu_c_abi.py compiles it with gcc
 * and compares its output with the Python path on the same inputs.
 *   usage: c_abi_consumer <in.bin> <out.bin>
 *   in.bin : int32 B, N, nnz, nin, n_out; float pcm[B][16000]; int32 w_rowptr[N+1], w_col[nnz], w_q[nnz];
 *            int32 in_rowptr[N+1], in_col[nin]; double in_val[nin]; double leak[N]; int32 out_idx[n_out]; double theta
 *   out.bin: uint8 spikes[B][128][400]; double features[B][5*n_out]                                              */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "lsm_b200.h"

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != LSM_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ctx ? lsm_last_error(ctx) : "");  \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

static void *rd(FILE *f, size_t bytes)
{
    void *p = malloc(bytes ? bytes : 1);
    if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char **argv)
{
    lsm_ctx *ctx = NULL;
    if (argc != 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t hdr[5];
    if (fread(hdr, 4, 5, f) != 5) return 2;
    const int B = hdr[0], N = hdr[1], nnz = hdr[2], nin = hdr[3], n_out = hdr[4];
    float *pcm = rd(f, (size_t)B * 16000 * 4);
    int32_t *w_rowptr = rd(f, (size_t)(N + 1) * 4), *w_col = rd(f, (size_t)nnz * 4), *w_q = rd(f, (size_t)nnz * 4);
    int32_t *in_rowptr = rd(f, (size_t)(N + 1) * 4), *in_col = rd(f, (size_t)nin * 4);
    double *in_val = rd(f, (size_t)nin * 8), *leak = rd(f, (size_t)N * 8);
    int32_t *out_idx = rd(f, (size_t)n_out * 4);
    double *theta = rd(f, 8);
    fclose(f);

    CHECK(lsm_ctx_create(&ctx, 0));

    /* stage 1: create_dataset.py:148-158 */
    lsm_frontend_params fp = {0};
    fp.kind = LSM_FILTERBANK_GAMMATONE; fp.channels = 128; fp.n_samples = 16000; fp.nwin = 400; fp.hop = 160;
    fp.n_bins = 100; fp.n_thresholds = 4; fp.redundancy = 1;
    const double thr[4] = {0.95, 0.90, 0.80, 0.70};             /* sorted(SPIKE_THRESHOLDS, reverse=True) */
    for (int k = 0; k < 4; ++k) { fp.thresholds_desc[k] = thr[k]; fp.lower_bounds[k] = thr[k] - 0.1; }
    double *coefs = malloc(sizeof(double) * 128 * 10);
    int32_t zi0[100];
    double zf[100];
    CHECK(lsm_gammatone_design(16000.0, 128, 50.0, coefs));
    CHECK(lsm_zoom_table(98, 100, zi0, zf));
    lsm_frontend *fe = NULL;
    CHECK(lsm_frontend_create(ctx, &fp, coefs, zi0, zf, &fe));

    /* stage 2+3: extract_lsm_features.py:188 and :79-87 */
    lsm_reservoir_params rp = {N, 128, 400, 2, 24, n_out, *theta};
    lsm_reservoir *res = NULL;
    CHECK(lsm_reservoir_create(ctx, &rp, w_rowptr, w_col, w_q, in_rowptr, in_col, in_val, leak, out_idx, &res));

    uint8_t *spikes = malloc((size_t)B * 128 * 400);
    double *features = malloc(sizeof(double) * (size_t)B * 5 * n_out);
    const uint32_t original = LSM_F_SPIKE_COUNTS | LSM_F_SPIKE_VARIANCES | LSM_F_MEAN_SPIKE_TIMES | LSM_F_MEAN_ISI | LSM_F_ISI_VARIANCES;
    CHECK(lsm_pipeline_run_host(ctx, fe, res, pcm, B, original, 1, features, spikes));

    /* error behaviour: bad arguments are refused with a message, nothing crashes */
    if (lsm_pipeline_run_host(ctx, fe, res, NULL, B, original, 1, features, NULL) != LSM_ERR_INVALID) return 3;
    lsm_frontend_params bad = fp;
    bad.channels = 4096;
    lsm_frontend *fe2 = NULL;
    if (lsm_frontend_create(ctx, &bad, coefs, zi0, zf, &fe2) == LSM_OK) return 3;

    FILE *g = fopen(argv[2], "wb");
    fwrite(coefs, sizeof(double), 128 * 10, g);
    fwrite(spikes, 1, (size_t)B * 128 * 400, g);
    fwrite(features, sizeof(double), (size_t)B * 5 * n_out, g);
    fclose(g);
    printf("launches %lld fused %d\n", (long long)lsm_launch_count(ctx), lsm_pipeline_is_fused(fe, res));
    lsm_reservoir_destroy(res);
    lsm_frontend_destroy(fe);
    lsm_ctx_destroy(ctx);
    return 0;
}
