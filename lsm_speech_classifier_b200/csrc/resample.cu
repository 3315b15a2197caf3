// Polyphase resampler - the sample-rate conversion of the ingest step.
//
// Replaces the resampling inside librosa.load(filepath, sr=16000, mono=True) at /root/reference/create_dataset.py:26 for
// files that are not already at 16 kHz.  librosa's default resampler (soxr_hq) is a closed recipe; this is its documented
// alternative res_type="polyphase" = scipy.signal.resample_poly(y, up, down) with the default Kaiser(5.0) window: an FIR
// low-pass at 1/max(up, down), evaluated only at the output instants (scipy.signal.upfirdn).  The taps come from the caller
// (designed on the host with scipy.signal.firwin exactly as scipy does, in the input's dtype, transposed and flipped per
// phase as scipy's _pad_h lays them out); the kernel repeats upfirdn's arithmetic - for each output sample one float32
// multiply and one float32 add per tap, in ascending input order - so the result equals scipy's bit for bit
// (oracle/pyref.py resample_poly_ref, tests/test_oracle_frontend.py, tests/test_gpu_ingest.py).
#include "lsm_common.cuh"

namespace {

// out[b][o], o in [0, n_out): y = o + n_pre_remove; phase t = (y * down) % up; newest input x_idx = (y * down) / up;
// out = sum_{j < hpp} x[x_idx - hpp + 1 + j] * htf[t * hpp + j] over the inputs that exist (zero padding outside)
__global__ void __launch_bounds__(256) resample_poly_kernel(const float *__restrict__ x, int n_in, int up, int down,
                                                            const float *__restrict__ htf, int hpp, int n_pre_remove, int n_out,
                                                            float *__restrict__ out)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    const float *xb = x + (size_t)blockIdx.y * n_in;
    const long long pos = (long long)(o + n_pre_remove) * down;
    const int t = (int)(pos % up);
    const int xi = (int)(pos / up);
    const float *h = htf + (size_t)t * hpp;
    int k = xi - hpp + 1, j = 0;
    if (k < 0) { j = -k; k = 0; }
    float acc = 0.0f;
    for (; j < hpp && k < n_in; ++j, ++k) acc = __fadd_rn(acc, __fmul_rn(__ldg(xb + k), __ldg(h + j)));
    out[(size_t)blockIdx.y * n_out + o] = acc;
}

}  // namespace

extern "C" int lsm_resample_poly(lsm_ctx *ctx, const float *d_in, int32_t B, int32_t n_in, int32_t up, int32_t down,
                                 const float *h_taps, int32_t taps_per_phase, int32_t n_pre_remove, int32_t n_out, float *d_out)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (B < 0 || n_in <= 0 || up < 1 || down < 1 || taps_per_phase < 1 || n_pre_remove < 0 || n_out < 0 || !h_taps ||
        (B > 0 && n_out > 0 && (!d_in || !d_out)))
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_resample_poly: bad argument");
    if (B == 0 || n_out == 0) return LSM_OK;
    if (B > 65535) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_resample_poly: at most 65535 signals per call");
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    void *d_taps;
    int rc;
    const size_t tap_bytes = sizeof(float) * (size_t)up * taps_per_phase;
    if ((rc = lsm_stage_device(ctx, 5, tap_bytes, &d_taps)) != LSM_OK) return rc;
    LSM_CUDA(ctx, cudaMemcpyAsync(d_taps, h_taps, tap_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const dim3 grid((n_out + 255) / 256, B);
    resample_poly_kernel<<<grid, 256, 0, ctx->stream>>>(d_in, n_in, up, down, (const float *)d_taps, taps_per_phase, n_pre_remove, n_out, d_out);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    LSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // h_taps (host) was the source of an asynchronous copy
    return LSM_OK;
}
