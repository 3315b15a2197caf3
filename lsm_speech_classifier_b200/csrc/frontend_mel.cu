// K1m — mel front end (placeholder until the STFT kernel lands; fails loudly, never falls back).
#include "lsm_common.cuh"

int lsm_mel_create(lsm_ctx *ctx, lsm_frontend *, const float *)
{
    LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "mel front end not built yet");
}
void lsm_mel_destroy(lsm_frontend *) {}
int lsm_launch_mel(lsm_ctx *ctx, lsm_frontend *, const float *, int, uint8_t *, double *, cudaStream_t)
{
    LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "mel front end not built yet");
}
