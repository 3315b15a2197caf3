// Multinomial logistic regression readout on the device - downstream of the path (second half):
// /root/reference/train_classifier.py:36-47  LogisticRegression(random_state=42, max_iter=1000).fit / .predict.
//
// Same objective as scikit-learn's lbfgs solver (sklearn/linear_model/_logistic.py + _linear_loss.py, multinomial, fit_intercept):
//     f(W, b) = (1/n) * sum_i [ logsumexp(z_i) - z_i[y_i] ] + (1 / (2 C n)) * ||W||_F^2,      z_i = W x_i + b
// (all K classes parametrised, intercept not penalised), minimised by L-BFGS (history 10, backtracking line search) until
// max|grad| <= tol, the relative decrease falls under 64 eps, or max_iter - the stopping rules scipy's L-BFGS-B is called with.
// The optimum of this strictly convex problem is unique, so the fitted model agrees with scikit-learn's to solver tolerance; the bar
// (tests/test_gpu_readout.py) is the reference's own metric: test accuracy within 0.5 points, plus >= 99 % identical predictions.
//
// Device work per evaluation: two passes over X (float64[n][F], the standardised feature matrix already in HBM):
//   logreg_forward_kernel   z = W x + b for 4 rows per warp, softmax, loss partials, residual R = (p - onehot) / n
//   logreg_grad_kernel      partial G[chunk][k][j] = sum_{i in chunk} R[i][k] X[i][j]   (thread = feature column, coalesced rows)
//   logreg_reduce_kernel    G = sum_chunk partial (fixed order) + W / (C n); intercept gradient
// Both passes are HBM-bound (154 MB per pass at n = 9600, F = 2000); the K <= 16 logits ride in registers.
// More than 16 classes (Speech Commands has 35; up to 64 here): the same two passes per block of 16 classes - logits into a
// Z[n][Kp] matrix, a row-wise softmax kernel, the gradient block by block - so the register budget stays that of 16 classes.
#include <float.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "lsm_common.cuh"

namespace {

constexpr int kMaxK = 16;      // classes whose logits ride in registers (the reference has 12)
constexpr int kMaxKWide = 64;  // classes served in blocks of 16
constexpr int kRowsPerWarp = 4;
constexpr int kGradChunks = 64;

// Wt: [F][K] feature-major copy of the coefficients (a lane reads the K weights of its feature contiguously)
template <int K>
__global__ void __launch_bounds__(256) logreg_forward_kernel(const double *__restrict__ X, const int32_t *__restrict__ y, int n, int F,
                                                             const double *__restrict__ Wt, const double *__restrict__ b,
                                                             double *__restrict__ R, double *__restrict__ loss_part,
                                                             int32_t *__restrict__ pred)
{
    __shared__ double s_loss[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_total = gridDim.x * 8;
    double loss = 0.0;
    for (int r0 = (blockIdx.x * 8 + warp) * kRowsPerWarp; r0 < n; r0 += warps_total * kRowsPerWarp) {
        double acc[kRowsPerWarp][K];
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
            for (int k = 0; k < K; ++k) acc[r][k] = 0.0;
        for (int j = lane; j < F; j += 32) {
            double w[K];
#pragma unroll
            for (int k = 0; k < K; ++k) w[k] = __ldg(Wt + (size_t)j * K + k);
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r) {
                const int i = min(r0 + r, n - 1);
                const double x = __ldcs(X + (size_t)i * F + j);
#pragma unroll
                for (int k = 0; k < K; ++k) acc[r][k] = fma(x, w[k], acc[r][k]);
            }
        }
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[r][k] += __shfl_xor_sync(0xffffffffu, acc[r][k], o);
        if (lane < kRowsPerWarp && r0 + lane < n) {
            const int i = r0 + lane;
            double z[K];
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r)
                if (r == lane)
#pragma unroll
                    for (int k = 0; k < K; ++k) z[k] = acc[r][k] + __ldg(b + k);
            double zmax = z[0];
            int arg = 0;
#pragma unroll
            for (int k = 1; k < K; ++k) if (z[k] > zmax) { zmax = z[k]; arg = k; }
            if (pred) pred[i] = arg;
            if (R) {
                double se = 0.0, e[K];
#pragma unroll
                for (int k = 0; k < K; ++k) { e[k] = exp(z[k] - zmax); se += e[k]; }
                const int yi = __ldg(y + i);
                const double inv = 1.0 / se, invn = 1.0 / (double)n;
                double zy = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (k == yi) zy = z[k];
                    R[(size_t)i * K + k] = (e[k] * inv - (k == yi ? 1.0 : 0.0)) * invn;
                }
                loss += (log(se) + zmax - zy) * invn;
            }
        }
    }
    if (loss_part) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
        if (lane == 0) s_loss[warp] = loss;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += s_loss[w];
            loss_part[blockIdx.x] = t;
        }
    }
}

// thread = feature column j, CTA = 128 columns x one chunk of rows; R of the chunk staged in shared memory
template <int K>
__global__ void __launch_bounds__(128) logreg_grad_kernel(const double *__restrict__ X, const double *__restrict__ R, int n, int F,
                                                          int rows_per_chunk, double *__restrict__ part)
{
    extern __shared__ double s_R[];     // [rows][K]
    const int j = blockIdx.x * 128 + threadIdx.x;
    const int i0 = blockIdx.y * rows_per_chunk;
    const int rows = max(0, min(rows_per_chunk, n - i0));
    for (int t = threadIdx.x; t < rows * K; t += 128) s_R[t] = R[(size_t)i0 * K + t];
    __syncthreads();
    double g[K];
#pragma unroll
    for (int k = 0; k < K; ++k) g[k] = 0.0;
    if (j < F) {
        int r = 0;
        for (; r + 4 <= rows; r += 4) {
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = __ldcs(X + (size_t)(i0 + r + u) * F + j);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < K; ++k) g[k] = fma(x[u], s_R[(r + u) * K + k], g[k]);
        }
        for (; r < rows; ++r) {
            const double x = __ldcs(X + (size_t)(i0 + r) * F + j);
#pragma unroll
            for (int k = 0; k < K; ++k) g[k] = fma(x, s_R[r * K + k], g[k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) part[((size_t)blockIdx.y * K + k) * F + j] = g[k];
    }
}

// G[k][j] = sum_chunks part + l2 * W[k][j];  gb[k] = sum_i R[i][k]
template <int K>
__global__ void __launch_bounds__(128) logreg_reduce_kernel(const double *__restrict__ part, int chunks, int F, const double *__restrict__ Wt,
                                                            double l2, const double *__restrict__ R, int n, double *__restrict__ G,
                                                            double *__restrict__ gb)
{
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j < F) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int c = 0; c < chunks; ++c) s += part[((size_t)c * K + k) * F + j];
            G[(size_t)k * F + j] = s + l2 * Wt[(size_t)j * K + k];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < K) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += R[(size_t)i * K + threadIdx.x];
        gb[threadIdx.x] = s;
    }
}


// ---- more than 16 classes: blocks of 16 (Wt, Z, R are padded to Kp = a multiple of 16 columns; padding columns are zero)
__global__ void __launch_bounds__(256) logreg_logits16_kernel(const double *__restrict__ X, int n, int F, const double *__restrict__ Wt,
                                                              int Kp, int k0, const double *__restrict__ b, double *__restrict__ Z)
{
    constexpr int K = 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_total = gridDim.x * 8;
    for (int r0 = (blockIdx.x * 8 + warp) * kRowsPerWarp; r0 < n; r0 += warps_total * kRowsPerWarp) {
        double acc[kRowsPerWarp][K];
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
            for (int k = 0; k < K; ++k) acc[r][k] = 0.0;
        for (int j = lane; j < F; j += 32) {
            double w[K];
#pragma unroll
            for (int k = 0; k < K; ++k) w[k] = __ldg(Wt + (size_t)j * Kp + k0 + k);
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r) {
                const int i = min(r0 + r, n - 1);
                const double x = __ldcs(X + (size_t)i * F + j);
#pragma unroll
                for (int k = 0; k < K; ++k) acc[r][k] = fma(x, w[k], acc[r][k]);
            }
        }
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[r][k] += __shfl_xor_sync(0xffffffffu, acc[r][k], o);
        if (lane < kRowsPerWarp && r0 + lane < n) {
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r)
                if (r == lane)
#pragma unroll
                    for (int k = 0; k < K; ++k) Z[(size_t)(r0 + r) * Kp + k0 + k] = acc[r][k] + __ldg(b + k0 + k);
        }
    }
}

// one thread per row: softmax over the K real classes, residual, loss partial per block (summed in a fixed order), prediction
__global__ void __launch_bounds__(256) logreg_softmax_kernel(const double *__restrict__ Z, const int32_t *__restrict__ y, int n, int K, int Kp,
                                                             double *__restrict__ R, double *__restrict__ loss_part, int32_t *__restrict__ pred)
{
    __shared__ double s_loss[256];
    const int i = blockIdx.x * 256 + threadIdx.x;
    double loss = 0.0;
    if (i < n) {
        const double *z = Z + (size_t)i * Kp;
        double zmax = z[0];
        int arg = 0;
        for (int k = 1; k < K; ++k) if (z[k] > zmax) { zmax = z[k]; arg = k; }
        if (pred) pred[i] = arg;
        if (R) {
            double se = 0.0;
            for (int k = 0; k < K; ++k) se += exp(z[k] - zmax);
            const int yi = __ldg(y + i);
            const double inv = 1.0 / se, invn = 1.0 / (double)n;
            for (int k = 0; k < Kp; ++k)
                R[(size_t)i * Kp + k] = k < K ? (exp(z[k] - zmax) * inv - (k == yi ? 1.0 : 0.0)) * invn : 0.0;
            loss = (log(se) + zmax - z[yi]) * invn;
        }
    }
    if (loss_part) {
        s_loss[threadIdx.x] = loss;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 256; ++w) t += s_loss[w];
            loss_part[blockIdx.x] = t;
        }
    }
}

__global__ void __launch_bounds__(128) logreg_grad16_kernel(const double *__restrict__ X, const double *__restrict__ R, int n, int F,
                                                            int rows_per_chunk, int Kp, int k0, double *__restrict__ part)
{
    constexpr int K = 16;
    extern __shared__ double s_R[];     // [rows][16]
    const int j = blockIdx.x * 128 + threadIdx.x;
    const int i0 = blockIdx.y * rows_per_chunk;
    const int rows = max(0, min(rows_per_chunk, n - i0));
    for (int t = threadIdx.x; t < rows * K; t += 128) s_R[t] = R[(size_t)(i0 + t / K) * Kp + k0 + (t % K)];
    __syncthreads();
    double g[K];
#pragma unroll
    for (int k = 0; k < K; ++k) g[k] = 0.0;
    if (j < F) {
        for (int r = 0; r < rows; ++r) {
            const double x = __ldcs(X + (size_t)(i0 + r) * F + j);
#pragma unroll
            for (int k = 0; k < K; ++k) g[k] = fma(x, s_R[r * K + k], g[k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) part[((size_t)blockIdx.y * Kp + k0 + k) * F + j] = g[k];
    }
}

__global__ void __launch_bounds__(128) logreg_reduce_wide_kernel(const double *__restrict__ part, int chunks, int F, const double *__restrict__ Wt,
                                                                 int K, int Kp, double l2, const double *__restrict__ R, int n,
                                                                 double *__restrict__ G, double *__restrict__ gb)
{
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j < F) {
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int c = 0; c < chunks; ++c) s += part[((size_t)c * Kp + k) * F + j];
            G[(size_t)k * F + j] = s + l2 * Wt[(size_t)j * Kp + k];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < K) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += R[(size_t)i * Kp + threadIdx.x];
        gb[threadIdx.x] = s;
    }
}

struct LogregWork {
    lsm_ctx *ctx;
    const double *d_X;
    const int32_t *d_y;
    int n, F, K;
    double l2;
    double *d_Wt, *d_b, *d_R, *d_part, *d_G, *d_gb, *d_loss;
    int fwd_blocks, rows_per_chunk;
    int Kp = 0, chunks = kGradChunks, sm_blocks = 0;      // wide path: padded class count, row chunks, softmax blocks
    double *d_Z = nullptr;
    std::vector<double> h_Wt, h_loss, h_G;
};

int eval_wide(LogregWork &w, const std::vector<double> &theta, double *f, std::vector<double> &grad)
{
    lsm_ctx *ctx = w.ctx;
    const int F = w.F, n = w.n, K = w.K, Kp = w.Kp;
    std::fill(w.h_Wt.begin(), w.h_Wt.end(), 0.0);
    for (int k = 0; k < K; ++k)
        for (int j = 0; j < F; ++j) w.h_Wt[(size_t)j * Kp + k] = theta[(size_t)k * F + j];
    std::vector<double> hb(Kp, 0.0);
    for (int k = 0; k < K; ++k) hb[k] = theta[(size_t)K * F + k];
    cudaStream_t st = ctx->stream;
    LSM_CUDA(ctx, cudaMemcpyAsync(w.d_Wt, w.h_Wt.data(), sizeof(double) * F * Kp, cudaMemcpyHostToDevice, st));
    LSM_CUDA(ctx, cudaMemcpyAsync(w.d_b, hb.data(), sizeof(double) * Kp, cudaMemcpyHostToDevice, st));
    for (int k0 = 0; k0 < Kp; k0 += 16)
        logreg_logits16_kernel<<<w.fwd_blocks, 256, 0, st>>>(w.d_X, n, F, w.d_Wt, Kp, k0, w.d_b, w.d_Z);
    logreg_softmax_kernel<<<w.sm_blocks, 256, 0, st>>>(w.d_Z, w.d_y, n, K, Kp, w.d_R, w.d_loss, nullptr);
    dim3 gg((F + 127) / 128, w.chunks);
    for (int k0 = 0; k0 < Kp; k0 += 16)
        logreg_grad16_kernel<<<gg, 128, sizeof(double) * w.rows_per_chunk * 16, st>>>(w.d_X, w.d_R, n, F, w.rows_per_chunk, Kp, k0, w.d_part);
    logreg_reduce_wide_kernel<<<(F + 127) / 128, 128, 0, st>>>(w.d_part, w.chunks, F, w.d_Wt, K, Kp, w.l2, w.d_R, n, w.d_G, w.d_gb);
    ctx->launches += 2 + 2 * (Kp / 16);
    LSM_CUDA(ctx, cudaGetLastError());
    LSM_CUDA(ctx, cudaMemcpyAsync(grad.data(), w.d_G, sizeof(double) * K * F, cudaMemcpyDeviceToHost, st));
    LSM_CUDA(ctx, cudaMemcpyAsync(grad.data() + (size_t)K * F, w.d_gb, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
    LSM_CUDA(ctx, cudaMemcpyAsync(w.h_loss.data(), w.d_loss, sizeof(double) * w.sm_blocks, cudaMemcpyDeviceToHost, st));
    LSM_CUDA(ctx, cudaStreamSynchronize(st));
    double loss = 0.0;
    for (int b = 0; b < w.sm_blocks; ++b) loss += w.h_loss[b];          // fixed order: deterministic
    double reg = 0.0;
    for (size_t t = 0; t < (size_t)K * F; ++t) reg += theta[t] * theta[t];
    *f = loss + 0.5 * w.l2 * reg;
    return LSM_OK;
}

template <int K>
int eval_t(LogregWork &w, const std::vector<double> &theta, double *f, std::vector<double> &grad)
{
    lsm_ctx *ctx = w.ctx;
    const int F = w.F, n = w.n;
    // theta = [W (K x F) | b (K)]  ->  feature-major Wt
    for (int k = 0; k < K; ++k)
        for (int j = 0; j < F; ++j) w.h_Wt[(size_t)j * K + k] = theta[(size_t)k * F + j];
    cudaStream_t st = ctx->stream;
    LSM_CUDA(ctx, cudaMemcpyAsync(w.d_Wt, w.h_Wt.data(), sizeof(double) * F * K, cudaMemcpyHostToDevice, st));
    LSM_CUDA(ctx, cudaMemcpyAsync(w.d_b, theta.data() + (size_t)K * F, sizeof(double) * K, cudaMemcpyHostToDevice, st));
    logreg_forward_kernel<K><<<w.fwd_blocks, 256, 0, st>>>(w.d_X, w.d_y, n, F, w.d_Wt, w.d_b, w.d_R, w.d_loss, nullptr);
    dim3 gg((F + 127) / 128, kGradChunks);
    logreg_grad_kernel<K><<<gg, 128, sizeof(double) * w.rows_per_chunk * K, st>>>(w.d_X, w.d_R, n, F, w.rows_per_chunk, w.d_part);
    logreg_reduce_kernel<K><<<(F + 127) / 128, 128, 0, st>>>(w.d_part, kGradChunks, F, w.d_Wt, w.l2, w.d_R, n, w.d_G, w.d_gb);
    ctx->launches += 3;
    LSM_CUDA(ctx, cudaGetLastError());
    LSM_CUDA(ctx, cudaMemcpyAsync(grad.data(), w.d_G, sizeof(double) * K * F, cudaMemcpyDeviceToHost, st));
    LSM_CUDA(ctx, cudaMemcpyAsync(grad.data() + (size_t)K * F, w.d_gb, sizeof(double) * K, cudaMemcpyDeviceToHost, st));
    LSM_CUDA(ctx, cudaMemcpyAsync(w.h_loss.data(), w.d_loss, sizeof(double) * w.fwd_blocks, cudaMemcpyDeviceToHost, st));
    LSM_CUDA(ctx, cudaStreamSynchronize(st));
    double loss = 0.0;
    for (int b = 0; b < w.fwd_blocks; ++b) loss += w.h_loss[b];          // fixed order: deterministic
    double reg = 0.0;
    for (size_t t = 0; t < (size_t)K * F; ++t) reg += theta[t] * theta[t];
    *f = loss + 0.5 * w.l2 * reg;
    return LSM_OK;
}

int eval(LogregWork &w, const std::vector<double> &theta, double *f, std::vector<double> &grad)
{
    if (w.Kp) return eval_wide(w, theta, f, grad);
    switch (w.K) {
#define LSM_CASE(KK) case KK: return eval_t<KK>(w, theta, f, grad);
        LSM_CASE(2) LSM_CASE(3) LSM_CASE(4) LSM_CASE(5) LSM_CASE(6) LSM_CASE(7) LSM_CASE(8) LSM_CASE(9) LSM_CASE(10)
        LSM_CASE(11) LSM_CASE(12) LSM_CASE(13) LSM_CASE(14) LSM_CASE(15) LSM_CASE(16)
#undef LSM_CASE
    }
    return LSM_ERR_UNSUPPORTED;
}

double dot(const std::vector<double> &a, const std::vector<double> &b)
{
    double s = 0.0;
    for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
}

}  // namespace

extern "C" int lsm_logreg_fit(lsm_ctx *ctx, const double *d_X, const int32_t *d_y, int32_t n, int32_t F, int32_t n_classes,
                              double C_reg, int32_t max_iter, double tol, double *h_coef, double *h_intercept, int32_t *h_n_iter)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!d_X || !d_y || n <= 0 || F <= 0 || !h_coef || !h_intercept || !(C_reg > 0.0) || max_iter < 0)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_logreg_fit: bad argument");
    if (n_classes < 3 || n_classes > kMaxKWide) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "lsm_logreg_fit: 3..%d classes supported (multinomial objective), got %d", kMaxKWide, n_classes);
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int K = n_classes;
    LogregWork w;
    w.ctx = ctx; w.d_X = d_X; w.d_y = d_y; w.n = n; w.F = F; w.K = K;
    w.l2 = 1.0 / (C_reg * (double)n);
    w.fwd_blocks = std::min((n + 8 * kRowsPerWarp - 1) / (8 * kRowsPerWarp), ctx->sm_count * 4);
    w.rows_per_chunk = (n + kGradChunks - 1) / kGradChunks;
    const size_t nP = (size_t)K * F + K;
    void *buf;
    int rc;
    if (K > kMaxK) {
        // blocks of 16 classes; row chunks sized so that a chunk's residual block fits in 48 KB of shared memory
        w.Kp = (K + 15) / 16 * 16;
        w.chunks = std::max(kGradChunks, (n + 383) / 384);
        w.rows_per_chunk = (n + w.chunks - 1) / w.chunks;
        w.sm_blocks = (n + 255) / 256;
        const size_t Kp = (size_t)w.Kp;
        const size_t bytes = sizeof(double) * ((size_t)F * Kp + Kp + 2 * (size_t)n * Kp + (size_t)w.chunks * Kp * F + (size_t)K * F + K + w.sm_blocks);
        if ((rc = lsm_stage_device(ctx, 7, bytes, &buf)) != LSM_OK) return rc;
        double *p = (double *)buf;
        w.d_Wt = p; p += (size_t)F * Kp;
        w.d_b = p; p += Kp;
        w.d_Z = p; p += (size_t)n * Kp;
        w.d_R = p; p += (size_t)n * Kp;
        w.d_part = p; p += (size_t)w.chunks * Kp * F;
        w.d_G = p; p += (size_t)K * F;
        w.d_gb = p; p += K;
        w.d_loss = p;
        w.h_Wt.resize((size_t)F * Kp); w.h_loss.resize(w.sm_blocks);
    } else {
    if (sizeof(double) * w.rows_per_chunk * K > 48 * 1024) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "lsm_logreg_fit: n too large for the residual tile");
    const size_t bytes = sizeof(double) * ((size_t)F * K + K + (size_t)n * K + (size_t)kGradChunks * K * F + (size_t)K * F + K + w.fwd_blocks);
    if ((rc = lsm_stage_device(ctx, 7, bytes, &buf)) != LSM_OK) return rc;
    double *p = (double *)buf;
    w.d_Wt = p; p += (size_t)F * K;
    w.d_b = p; p += K;
    w.d_R = p; p += (size_t)n * K;
    w.d_part = p; p += (size_t)kGradChunks * K * F;
    w.d_G = p; p += (size_t)K * F;
    w.d_gb = p; p += K;
    w.d_loss = p;
    w.h_Wt.resize((size_t)F * K); w.h_loss.resize(w.fwd_blocks);
    }

    // ---- L-BFGS (Nocedal & Wright alg. 7.4/7.5, history 10)
    const int M = 10;
    std::vector<double> x(nP, 0.0), g(nP), xn(nP), gn(nP), d(nP), q(nP);
    std::vector<std::vector<double>> S, Y;
    std::vector<double> rho;
    double f, fn;
    if ((rc = eval(w, x, &f, g)) != LSM_OK) return rc;
    int it = 0;
    auto gmax = [&](const std::vector<double> &v) { double m = 0; for (double t : v) m = std::max(m, fabs(t)); return m; };
    // evaluations are cheap here, so run to a tenth of the requested gradient tolerance: never worse than a solver that stops at tol
    const double gtol = 0.1 * tol;
    while (it < max_iter && gmax(g) > gtol) {
        // two-loop recursion
        q = g;
        std::vector<double> al(S.size());
        for (int i = (int)S.size() - 1; i >= 0; --i) {
            al[i] = rho[i] * dot(S[i], q);
            for (size_t t = 0; t < nP; ++t) q[t] -= al[i] * Y[i][t];
        }
        double gamma = 1.0;
        if (!S.empty()) gamma = dot(S.back(), Y.back()) / dot(Y.back(), Y.back());
        for (size_t t = 0; t < nP; ++t) q[t] *= gamma;
        for (size_t i = 0; i < S.size(); ++i) {
            const double be = rho[i] * dot(Y[i], q);
            for (size_t t = 0; t < nP; ++t) q[t] += (al[i] - be) * S[i][t];
        }
        for (size_t t = 0; t < nP; ++t) d[t] = -q[t];
        double dg0 = dot(d, g);
        if (!(dg0 < 0.0)) {          // not a descent direction (numerical): restart with steepest descent
            S.clear(); Y.clear(); rho.clear();
            for (size_t t = 0; t < nP; ++t) d[t] = -g[t];
            dg0 = dot(d, g);
        }
        // Backtracking line search on the sufficient-decrease condition.  The objective is strictly convex (L2 term), so
        // y.s > 0 holds for every accepted step and the curvature condition is not needed to keep the L-BFGS update valid.
        const double c1 = 1e-4;
        double a = S.empty() ? std::min(1.0, 1.0 / std::max(gmax(g), 1e-300)) : 1.0;
        bool ok = false;
        for (int ls = 0; ls < 50; ++ls) {
            for (size_t t = 0; t < nP; ++t) xn[t] = x[t] + a * d[t];
            if ((rc = eval(w, xn, &fn, gn)) != LSM_OK) return rc;
            if (fn <= f + c1 * a * dg0) { ok = true; break; }
            a *= 0.5;
        }
        if (!ok) break;                          // no progress possible at working precision
        // accept xn
        std::vector<double> s(nP), yv(nP);
        for (size_t t = 0; t < nP; ++t) { s[t] = xn[t] - x[t]; yv[t] = gn[t] - g[t]; }
        const double sy = dot(s, yv);
        const double f_old = f;
        x.swap(xn); g.swap(gn); f = fn;
        ++it;
        if (sy > 1e-10 * dot(yv, yv)) {
            if ((int)S.size() == M) { S.erase(S.begin()); Y.erase(Y.begin()); rho.erase(rho.begin()); }
            S.push_back(std::move(s)); Y.push_back(std::move(yv)); rho.push_back(1.0 / sy);
        }
        if ((f_old - f) <= 64.0 * DBL_EPSILON * std::max(std::max(fabs(f_old), fabs(f)), 1.0)) break;    // scipy's ftol
    }
    for (int k = 0; k < K; ++k) {
        for (int j = 0; j < F; ++j) h_coef[(size_t)k * F + j] = x[(size_t)k * F + j];
        h_intercept[k] = x[(size_t)K * F + k];
    }
    if (h_n_iter) *h_n_iter = it;
    return LSM_OK;
}

extern "C" int lsm_logreg_predict(lsm_ctx *ctx, const double *d_X, int32_t n, int32_t F, int32_t n_classes, const double *h_coef,
                                  const double *h_intercept, int32_t *d_pred)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (n < 0 || F <= 0 || !h_coef || !h_intercept || (n > 0 && (!d_X || !d_pred))) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_logreg_predict: bad argument");
    if (n_classes < 2 || n_classes > kMaxKWide) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "lsm_logreg_predict: 2..%d classes supported, got %d", kMaxKWide, n_classes);
    if (n == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int K = n_classes;
    void *buf;
    int rc;
    if (K > kMaxK) {
        const int Kp = (K + 15) / 16 * 16;
        if ((rc = lsm_stage_device(ctx, 7, sizeof(double) * ((size_t)F * Kp + Kp + (size_t)n * Kp), &buf)) != LSM_OK) return rc;
        std::vector<double> wt((size_t)F * Kp, 0.0), hb(Kp, 0.0);
        for (int k = 0; k < K; ++k) {
            for (int j = 0; j < F; ++j) wt[(size_t)j * Kp + k] = h_coef[(size_t)k * F + j];
            hb[k] = h_intercept[k];
        }
        double *d_Wt = (double *)buf, *d_b = d_Wt + (size_t)F * Kp, *d_Z = d_b + Kp;
        cudaStream_t st = ctx->stream;
        LSM_CUDA(ctx, cudaMemcpyAsync(d_Wt, wt.data(), sizeof(double) * F * Kp, cudaMemcpyHostToDevice, st));
        LSM_CUDA(ctx, cudaMemcpyAsync(d_b, hb.data(), sizeof(double) * Kp, cudaMemcpyHostToDevice, st));
        const int blocks = std::min((n + 8 * kRowsPerWarp - 1) / (8 * kRowsPerWarp), ctx->sm_count * 4);
        for (int k0 = 0; k0 < Kp; k0 += 16) logreg_logits16_kernel<<<blocks, 256, 0, st>>>(d_X, n, F, d_Wt, Kp, k0, d_b, d_Z);
        logreg_softmax_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_Z, nullptr, n, K, Kp, nullptr, nullptr, d_pred);
        ctx->launches += 1 + Kp / 16;
        LSM_CUDA(ctx, cudaGetLastError());
        LSM_CUDA(ctx, cudaStreamSynchronize(st));
        return LSM_OK;
    }
    if ((rc = lsm_stage_device(ctx, 7, sizeof(double) * ((size_t)F * K + K), &buf)) != LSM_OK) return rc;
    std::vector<double> wt((size_t)F * K);
    for (int k = 0; k < K; ++k)
        for (int j = 0; j < F; ++j) wt[(size_t)j * K + k] = h_coef[(size_t)k * F + j];
    double *d_Wt = (double *)buf, *d_b = d_Wt + (size_t)F * K;
    cudaStream_t st = ctx->stream;
    LSM_CUDA(ctx, cudaMemcpyAsync(d_Wt, wt.data(), sizeof(double) * F * K, cudaMemcpyHostToDevice, st));
    LSM_CUDA(ctx, cudaMemcpyAsync(d_b, h_intercept, sizeof(double) * K, cudaMemcpyHostToDevice, st));
    const int blocks = std::min((n + 8 * kRowsPerWarp - 1) / (8 * kRowsPerWarp), ctx->sm_count * 4);
    switch (K) {
#define LSM_CASE(KK) case KK: logreg_forward_kernel<KK><<<blocks, 256, 0, st>>>(d_X, nullptr, n, F, d_Wt, d_b, nullptr, nullptr, d_pred); break;
        LSM_CASE(2) LSM_CASE(3) LSM_CASE(4) LSM_CASE(5) LSM_CASE(6) LSM_CASE(7) LSM_CASE(8) LSM_CASE(9) LSM_CASE(10)
        LSM_CASE(11) LSM_CASE(12) LSM_CASE(13) LSM_CASE(14) LSM_CASE(15) LSM_CASE(16)
#undef LSM_CASE
    }
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    LSM_CUDA(ctx, cudaStreamSynchronize(st));      // wt (host) was the source of an async copy
    return LSM_OK;
}
