// Standardisation of the feature matrix on the device: sklearn.preprocessing.StandardScaler as the reference applies
// it at /root/reference/extract_lsm_features.py:199-201 (fit on train, transform train and test), SURVEY.md 8(f) rank 1.
//
// Bit-exact with scikit-learn's dense float64 path (sklearn/utils/extmath.py _incremental_mean_and_var, first batch):
//   sum_j   = X[0][j] + X[1][j] + ...            (numpy axis-0 reduction: rows added in order)
//   mean_j  = sum_j / n
//   corr_j  = sum_i (X[i][j] - mean_j);  sq_j = sum_i (X[i][j] - mean_j)^2      (same order)
//   var_j   = (sq_j - corr_j*corr_j / n) / n
//   constant feature  <=>  var_j <= n*eps*var_j + (n*mean_j*eps)^2   ->  scale_j = 1, else sqrt(var_j)
//   out     = (X - mean) / scale
// One thread per column (coalesced across the row), rows strictly in order; explicit *_rn operations.
#include <float.h>

#include "lsm_common.cuh"

namespace {

__global__ void __launch_bounds__(128) scaler_fit_kernel(const double *__restrict__ X, int n, int F, double *__restrict__ mean,
                                                         double *__restrict__ var, double *__restrict__ scale)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= F) return;
    const double dn = (double)n;
    double s = 0.0;
    int i = 0;
    for (; i + 8 <= n; i += 8) {               // loads issued together, adds strictly in row order
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(X + (size_t)(i + u) * F + j);
#pragma unroll
        for (int u = 0; u < 8; ++u) s = __dadd_rn(s, v[u]);
    }
    for (; i < n; ++i) s = __dadd_rn(s, __ldg(X + (size_t)i * F + j));
    const double m = __ddiv_rn(s, dn);
    double corr = 0.0, sq = 0.0;
    i = 0;
    for (; i + 8 <= n; i += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(X + (size_t)(i + u) * F + j);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double t = __dsub_rn(v[u], m);
            corr = __dadd_rn(corr, t);
            sq = __dadd_rn(sq, __dmul_rn(t, t));
        }
    }
    for (; i < n; ++i) {
        const double t = __dsub_rn(__ldg(X + (size_t)i * F + j), m);
        corr = __dadd_rn(corr, t);
        sq = __dadd_rn(sq, __dmul_rn(t, t));
    }
    const double unnorm = __dsub_rn(sq, __ddiv_rn(__dmul_rn(corr, corr), dn));
    const double v = __ddiv_rn(unnorm, dn);
    // _is_constant_feature: var <= n*eps*var + (n*mean*eps)^2
    const double eps = DBL_EPSILON;
    const double nme = __dmul_rn(__dmul_rn(dn, m), eps);
    const double bound = __dadd_rn(__dmul_rn(__dmul_rn(dn, eps), v), __dmul_rn(nme, nme));
    mean[j] = m;
    var[j] = v;
    scale[j] = (v <= bound) ? 1.0 : __dsqrt_rn(v);
}

__global__ void __launch_bounds__(256) scaler_transform_kernel(const double *__restrict__ X, long long total, int F,
                                                               const double *__restrict__ mean, const double *__restrict__ scale,
                                                               double *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int j = (int)(idx % F);
        out[idx] = __ddiv_rn(__dsub_rn(X[idx], __ldg(mean + j)), __ldg(scale + j));
    }
}

}  // namespace

extern "C" int lsm_standardize_fit(lsm_ctx *ctx, const double *d_X, int32_t n, int32_t F, double *d_mean, double *d_var,
                                   double *d_scale)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!d_X || n <= 0 || F <= 0 || !d_mean || !d_var || !d_scale) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_standardize_fit: bad argument");
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    scaler_fit_kernel<<<(F + 127) / 128, 128, 0, ctx->stream>>>(d_X, n, F, d_mean, d_var, d_scale);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return LSM_OK;
}

extern "C" int lsm_standardize_transform(lsm_ctx *ctx, const double *d_X, int32_t n, int32_t F, const double *d_mean,
                                         const double *d_scale, double *d_out)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (n < 0 || F <= 0 || !d_mean || !d_scale || (n > 0 && (!d_X || !d_out))) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_standardize_transform: bad argument");
    if (n == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    scaler_transform_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_X, (long long)n * F, F, d_mean, d_scale, d_out);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return LSM_OK;
}
