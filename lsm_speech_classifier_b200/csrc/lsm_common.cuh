// Internal declarations shared by the sm_100a translation units of liblsmb200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "lsm_b200.h"
#include "sqrt_rn.cuh"

struct lsm_copy_pool;
void lsm_copy_pool_free(lsm_copy_pool *p);

struct lsm_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;   // created by the ctx
    cudaStream_t stream = nullptr;       // where work goes (own_stream or the caller's)
    cudaStream_t copy_stream[2] = {nullptr, nullptr};  // pipeline_run_host: H2D / D2H legs
    cudaEvent_t ev[12] = {};             // [0,6): chunked host pipeline; 8: fork, 9-10: join of the two launch lanes
    int64_t launches = 0;
    cudaEvent_t ev_stage_full[4] = {}, ev_stage_free[4] = {};   // copy-engine feed: staging ring hand-over
    int stage_busy[4] = {};
    unsigned stage_next = 0;
    struct lsm_copy_pool *copy_pool = nullptr;   // host threads that move pageable caller buffers through the pinned ring (api.cu)
    int host_feed = 0;                   // lsm_ctx_set_host_feed: 0 = the kernel reads pinned host PCM itself, 1 = copy engine first
    char err[512] = {0};
    // staging owned by the ctx for the *_host entry points (grown on demand)
    void *d_stage[16] = {};
    size_t d_stage_bytes[16] = {};
    void *h_pin[4] = {};
    size_t h_pin_bytes[4] = {};
};

struct lsm_frontend {
    lsm_frontend_params p;
    int ncols = 0;                 // spectrogram columns before zoom (98 gammatone / 101 mel)
    double *d_coefs = nullptr;     // gammatone: [C][10] design rows
    int32_t *d_zoom_i0 = nullptr;  // [n_bins]
    double *d_zoom_f = nullptr;    // [n_bins]
    double *d_scratch = nullptr;   // per-CTA [ncols][C] dB plane (stays in L2)
    int grid = 0;
    int *d_counters = nullptr;     // [64] dynamic work counters, one per in-flight launch; [64] = utterances filtered twice
    int mode = 1;                  // gammatone: 0 = exact filter only, 1 = speculative filter + exact re-execution of near-ties
    double spec_delta = 0.0;       // extra dB margin of the near-tie test on top of the derived bound (lsm_frontend_set_mode)
    double bound_scale = 1.0;      // multiplier of the derived bound (lsm_frontend_set_bound_scale; diagnostics only)
    double h_kappa[256] = {};      // per channel: bound on |amplitude(speculative) - amplitude(exact)| per unit max|x| (lsm_gammatone_error_bound)
    double *d_kappa = nullptr;     // [C] the same on the device
    float *d_xmax = nullptr;       // [2][energy_cap] max |sample| per utterance, written by the filter warps (lanes arrangements)
    int *d_pipe_sync = nullptr;    // pipeline kernel: per launch lane work counter + group completion counters; error flag last
    double h_lane_coef[256][6] = {};  // per channel c1..c4, -a1, -a2 of the normalised cascade (kernel-parameter table of K1a)
    double *d_energy = nullptr;    // [2][energy_cap][ncols][C] raw window energies between the filter warps and the encoder (one set per launch lane)
    int energy_cap = 0;
    int *d_rerun = nullptr;        // [1 + energy_cap] count + utterances the encoder kernel flagged for the exact pass
    int rerun_cap = 0;
    const int16_t *next_pcm16 = nullptr;   // set by the *_i16 entry points for the launch they make (d_pcm == nullptr), then cleared
    unsigned counter_next = 0;
    // launches on different streams share the scratch planes: each launch waits for the previous one's event
    // two scratch slots: launch n uses slot n & 1 and only has to wait for launch n-2, so two launches (on two streams) can be
    // in flight and the drain tail of one overlaps the start of the next; slot -1 = exclusive (waits for / blocks both)
    cudaEvent_t ev_slot[2] = {nullptr, nullptr};
    cudaStream_t slot_stream[2] = {nullptr, nullptr};
    int slot_valid[2] = {0, 0};
    unsigned slot_next = 0;
    // mel
    float *d_mel_w = nullptr;      // packed non-zero mel weights
    int32_t *d_mel_lo = nullptr;   // [C] first non-zero bin
    int32_t *d_mel_n = nullptr;    // [C] number of non-zero bins
    int32_t *d_mel_off = nullptr;  // [C] offset into d_mel_w
    double *d_window = nullptr;    // [n_fft] periodic hann
    double2 *d_twiddle = nullptr;  // [n_fft/4]  exp(-2 pi i q / (n_fft/2))
    double2 *d_twiddle2 = nullptr; // [n_fft/2 + 1]  exp(-2 pi i k / n_fft)
    float *d_mel_scratch = nullptr; // per-CTA [ncols][C] mel power / dB plane
    double mel_tw1[62] = {0};      // host copy of the twiddles of FFT stages 1-5 (kernel-parameter constants of the warp-per-frame kernel)
    int grid_warp = 0;             // grid of mel_finish_kernel<0> (warp-per-frame arrangement; 0: packed weights too long, block kernel only)
    uint32_t *d_mel_sched = nullptr; // [mel_sched_len][32][2] projection schedule of the warp-per-frame kernel
    int mel_sched_len = 0;
    int finish_threads = 256;      // block size of mel_finish_kernel<0>: the channels rounded up to whole warps
    float *d_mel_power = nullptr;  // [mel_power_cap][ncols][C] mel power between the two kernels of the warp-per-frame arrangement
    int mel_power_cap = 0;
};

struct lsm_reservoir {
    lsm_reservoir_params p;
    int n_pad = 0;                 // row pitch of the dense weight plane (multiple of 32)
    int32_t *d_wt = nullptr;       // dense [N][n_pad]: row = PRESYNAPTIC j, column = postsynaptic i
    double *d_wt64 = nullptr;      // strict reservoirs (lsm_reservoir_create_f64): the same plane in fp64, summed in ascending presynaptic order
    int w64 = 0;
    int32_t *d_in_rowptr = nullptr, *d_in_col = nullptr;
    double *d_in_val = nullptr;
    double *d_leak = nullptr;
    int32_t *d_out_slot = nullptr; // [N] position in the output list or -1
    int32_t *d_in_row = nullptr;   // [N] single input row, -1 none, -2 several
    int max_in_per_neuron = 0;
    int lean = 0;                  // uniform leak and gain, one input row per driven neuron, relabelled layout (reservoir_core.cuh)
    int32_t *d_ext_id = nullptr;   // [n_pad] lean layout: external index of internal neuron slot (>= N: padding)
    int zero_row = 0;              // index of the all-zero weight row
    double *gather_out[8] = {};    // fused all-gather destinations (lsm_reservoir_set_gather), n_gather = 0: off
    long long gather_row0 = 0;
    int n_gather = 0;
    int skip_dead_time = 0;        // theta > 0 and 0 <= leak <= 1 everywhere: silent stretches of an utterance are exact no-ops
    double c_off = 0.0, c_on = 0.0;
    int hi_magic = 0;
    double leak0 = 0.0, gain0 = 0.0;
    // dense (tensor-core) arm, reservoir_dense.cu: host tables in internal neuron order, digit-plane count, workspace
    int mode = 0;                  // lsm_reservoir_set_mode: 0 event-driven, 1 dense
    int32_t *h_dense_in_row = nullptr, *h_dense_out_int = nullptr;   // [n_pad] input row or -1 / [n_out] internal index of output o
    double *h_dense_gain = nullptr, *h_dense_leak = nullptr;         // [n_pad]
    int dense_planes = 0, dense_top_signed = 0;
    struct lsm_dense_ws *dense = nullptr;
};

#define LSM_FAIL(ctx, code, ...)                                  \
    do {                                                          \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);    \
        return (code);                                            \
    } while (0)

#define LSM_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            LSM_FAIL(ctx, LSM_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,          \
                     cudaGetErrorString(e_));                                                 \
    } while (0)

// grow-only device / pinned staging buffers owned by the ctx
int lsm_stage_device(lsm_ctx *ctx, int slot, size_t bytes, void **out);
int lsm_stage_pinned(lsm_ctx *ctx, int slot, size_t bytes, void **out);

// stage launchers (no host sync)
int lsm_launch_gammatone(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes,
                         double *d_spec_norm, cudaStream_t st);
int lsm_gammatone_grid(lsm_ctx *ctx, const lsm_frontend_params *p, int *grid);
struct GtArgs;
void lsm_gammatone_fill_args(const lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes, double *d_spec_norm, GtArgs *out);
bool lsm_lanes_eligible(const lsm_frontend *fe, const void *d_pcm);
int lsm_frontend_ensure_energy(lsm_ctx *ctx, lsm_frontend *fe, int B);
int lsm_frontend_ensure_rerun(lsm_ctx *ctx, lsm_frontend *fe, int B);
bool lsm_pipeline_lanes_eligible(const lsm_frontend *fe, const lsm_reservoir *res, const void *d_pcm, bool i16);
int lsm_launch_pipeline_lanes(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B,
                              uint8_t *d_spikes_or_null, uint32_t feature_mask, int nan_to_num, double *d_features,
                              cudaStream_t st, int lane, long long row0);
int lsm_pipeline_lanes_device(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B, uint8_t *d_spikes,
                              uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st);
int lsm_launch_audit(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, double *d_out, cudaStream_t st);
extern "C" int lsm_gammatone_error_bound(const double *h_table, int32_t channels, int32_t n_samples, double *h_kappa);
int lsm_frontend_order_before(lsm_ctx *ctx, lsm_frontend *fe, cudaStream_t st, int slot = -1);   // call before a launch that uses fe's scratch
int lsm_frontend_order_after(lsm_ctx *ctx, lsm_frontend *fe, cudaStream_t st, int slot = -1);    // ... and right after it
int lsm_frontend_wait_idle(lsm_ctx *ctx, lsm_frontend *fe);                                      // host waits for every launch of fe
int lsm_fused_npt(const lsm_frontend *fe, const lsm_reservoir *res);
int lsm_launch_fused(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B,
                     uint8_t *d_spikes_or_null, uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st,
                     long long row0);
int lsm_launch_fused_args(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, GtArgs &a, cudaStream_t st, bool launch, int *wave,
                          int max_grid, int forced_slot);
int lsm_fused_wave(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res);
bool lsm_ws_eligible(const lsm_frontend *fe, const lsm_reservoir *res);
int lsm_launch_fused_ws(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, GtArgs &a, cudaStream_t st, int *wave, bool launch);
struct ResArgs;
void lsm_reservoir_fill_args(const lsm_reservoir *res, const uint8_t *d_spikes, int B, uint32_t feature_mask,
                             int nan_to_num, double *d_features, uint8_t *d_raster, ResArgs *out);
int lsm_launch_mel(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes,
                   double *d_spec_norm, cudaStream_t st);
int lsm_launch_reservoir(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int B,
                         uint32_t feature_mask, int nan_to_num, double *d_features, uint8_t *d_raster,
                         cudaStream_t st, int *d_diag = nullptr, long long row0 = 0);
bool lsm_mel_fused_ok(const lsm_frontend *fe, const lsm_reservoir *res);
int lsm_launch_mel_fused(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B, uint8_t *d_spikes_or_null,
                         uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st, long long row0);
void lsm_reservoir_geometry(int N, int *npt, int *threads, int *n_pad);
// dense arm (reservoir_dense.cu)
const char *lsm_dense_unsupported(const lsm_reservoir *res);
void lsm_dense_ws_free(struct lsm_dense_ws *w);
int lsm_launch_reservoir_dense(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int B, uint32_t feature_mask,
                               int nan_to_num, double *d_features, uint8_t *d_raster, cudaStream_t st);
int lsm_reservoir_dense_probe_launch(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_s, int B, int32_t *d_acc, cudaStream_t st);
int lsm_mel_create(lsm_ctx *ctx, lsm_frontend *fe, const float *h_basis);
void lsm_mel_destroy(lsm_frontend *fe);
int lsm_mel_set_tables(lsm_ctx *ctx, lsm_frontend *fe, const double *h_win, const double *h_tw, const double *h_tw2);

// ---------------------------------------------------------------------------------------------
// fp64 helpers: every value-producing operation is an explicit round-to-nearest intrinsic, so
// nvcc can neither contract a*b+c into an FMA nor re-associate.  The CPU oracle performs the
// same IEEE-754 operations in the same order (oracle/lsm_oracle.c, -ffp-contract=off).
__device__ __forceinline__ double mul64(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add64(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub64(double a, double b) { return __dsub_rn(a, b); }

// Deterministic log10 for finite x > 0: fdlibm e_log.c kernel + FreeBSD e_log10.c recombination,
// operation for operation the same as oracle_log10() in oracle/lsm_oracle.c.
__device__ __forceinline__ double lsm_log10(double x)
{
    const double LG1 = 6.666666666666735130e-01, LG2 = 3.999999999940941908e-01,
                 LG3 = 2.857142874366239149e-01, LG4 = 2.222219843214978396e-01,
                 LG5 = 1.818357216161805012e-01, LG6 = 1.531383769920937332e-01,
                 LG7 = 1.479819860511658591e-01;
    const double IVLN10HI = 4.34294481878168880939e-01, IVLN10LO = 2.50829467116452752298e-11,
                 LOG10_2HI = 3.01029995663611771306e-01, LOG10_2LO = 3.69423907715893078616e-13;
    int hx = __double2hiint(x);
    int lx = __double2loint(x);
    int k = 0;
    if (hx < 0x00100000) {
        x = mul64(x, 18014398509481984.0);
        hx = __double2hiint(x);
        lx = __double2loint(x);
        k -= 54;
    }
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int i = (hx + 0x95f64) & 0x100000;
    x = __hiloint2double(hx | (i ^ 0x3ff00000), lx);
    k += (i >> 20);
    double f = sub64(x, 1.0);
    double hfsq = mul64(mul64(0.5, f), f);
    double s = div_rn_inline(f, add64(2.0, f));                 // = __ddiv_rn without its slow-path branch (sqrt_rn.cuh)
    double z = mul64(s, s);
    double w = mul64(z, z);
    double t1 = mul64(w, add64(LG2, mul64(w, add64(LG4, mul64(w, LG6)))));
    double t2 = mul64(z, add64(LG1, mul64(w, add64(LG3, mul64(w, add64(LG5, mul64(w, LG7)))))));
    double R = add64(t2, t1);
    double hi = sub64(f, hfsq);
    hi = __hiloint2double(__double2hiint(hi), 0);
    double lo = add64(sub64(sub64(f, hi), hfsq), mul64(s, add64(hfsq, R)));
    double val_hi = mul64(hi, IVLN10HI);
    double dk = (double)k;
    double y2 = mul64(dk, LOG10_2HI);
    double val_lo = add64(add64(mul64(dk, LOG10_2LO), mul64(add64(lo, hi), IVLN10LO)), mul64(lo, IVLN10HI));
    double ww = add64(y2, val_hi);
    val_lo = add64(val_lo, add64(sub64(y2, ww), val_hi));
    val_hi = ww;
    return add64(val_lo, val_hi);
}

__device__ __forceinline__ double warp_max_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
