// liblsmb200.so — C ABI (include/lsm_b200.h): context, handles, host-buffer entry points and the
// chunked audio -> features pipeline.  No CPU fallback anywhere: every entry point needs a ctx,
// and a ctx needs a CUDA device.
#include <math.h>
#include <stdlib.h>

#include <complex>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "lsm_common.cuh"

// ------------------------------------------------------------------------------------ context
extern "C" int lsm_ctx_create(lsm_ctx **out, int device_ordinal)
{
    if (!out) return LSM_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_ordinal < 0 || device_ordinal >= n)
        return LSM_ERR_CUDA;
    lsm_ctx *ctx = new (std::nothrow) lsm_ctx();
    if (!ctx) return LSM_ERR_NOMEM;
    ctx->device = device_ordinal;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) { delete ctx; return LSM_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) { delete ctx; return LSM_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        // built for sm_100a only; fail loudly rather than at the first launch
        delete ctx;
        return LSM_ERR_UNSUPPORTED;
    }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LSM_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    for (int i = 0; i < 2; ++i) cudaStreamCreateWithFlags(&ctx->copy_stream[i], cudaStreamNonBlocking);
    for (int i = 0; i < 12; ++i) cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming);
    for (int i = 0; i < 4; ++i) { cudaEventCreateWithFlags(&ctx->ev_stage_full[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_stage_free[i], cudaEventDisableTiming); }
    *out = ctx;
    return LSM_OK;
}

extern "C" void lsm_ctx_destroy(lsm_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 16; ++i) if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
    for (int i = 0; i < 4; ++i) if (ctx->h_pin[i]) cudaFreeHost(ctx->h_pin[i]);
    for (int i = 0; i < 12; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 4; ++i) { if (ctx->ev_stage_full[i]) cudaEventDestroy(ctx->ev_stage_full[i]); if (ctx->ev_stage_free[i]) cudaEventDestroy(ctx->ev_stage_free[i]); }
    for (int i = 0; i < 2; ++i) if (ctx->copy_stream[i]) cudaStreamDestroy(ctx->copy_stream[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    lsm_copy_pool_free(ctx->copy_pool);
    delete ctx;
}

extern "C" const char *lsm_last_error(const lsm_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }

extern "C" int lsm_set_stream(lsm_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return LSM_ERR_INVALID;
    ctx->stream = (cudaStream_t)cuda_stream;
    return LSM_OK;
}

extern "C" int lsm_reset_stream(lsm_ctx *ctx)
{
    if (!ctx) return LSM_ERR_INVALID;
    ctx->stream = ctx->own_stream;
    return LSM_OK;
}

extern "C" int lsm_sync(lsm_ctx *ctx)
{
    if (!ctx) return LSM_ERR_INVALID;
    LSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSM_OK;
}

extern "C" int lsm_ctx_set_host_feed(lsm_ctx *ctx, int32_t mode)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (mode != LSM_FEED_ZERO_COPY && mode != LSM_FEED_COPY_ENGINE) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_ctx_set_host_feed: mode must be 0 or 1");
    ctx->host_feed = mode;
    return LSM_OK;
}

extern "C" int64_t lsm_launch_count(const lsm_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int lsm_sm_count(const lsm_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

int lsm_stage_device(lsm_ctx *ctx, int slot, size_t bytes, void **out)
{
    if (ctx->d_stage_bytes[slot] < bytes) {
        // growing: asynchronous work (the launch lanes) may still be using the old buffer
        if (ctx->d_stage[slot]) { LSM_CUDA(ctx, cudaDeviceSynchronize()); LSM_CUDA(ctx, cudaFree(ctx->d_stage[slot])); ctx->d_stage[slot] = nullptr; ctx->d_stage_bytes[slot] = 0; }
        if (cudaMalloc(&ctx->d_stage[slot], bytes) != cudaSuccess) LSM_FAIL(ctx, LSM_ERR_NOMEM, "cudaMalloc(%zu) failed", bytes);
        ctx->d_stage_bytes[slot] = bytes;
    }
    *out = ctx->d_stage[slot];
    return LSM_OK;
}

int lsm_stage_pinned(lsm_ctx *ctx, int slot, size_t bytes, void **out)
{
    if (ctx->h_pin_bytes[slot] < bytes) {
        if (ctx->h_pin[slot]) { LSM_CUDA(ctx, cudaFreeHost(ctx->h_pin[slot])); ctx->h_pin[slot] = nullptr; ctx->h_pin_bytes[slot] = 0; }
        if (cudaMallocHost(&ctx->h_pin[slot], bytes) != cudaSuccess) LSM_FAIL(ctx, LSM_ERR_NOMEM, "cudaMallocHost(%zu) failed", bytes);
        ctx->h_pin_bytes[slot] = bytes;
    }
    *out = ctx->h_pin[slot];
    return LSM_OK;
}

template <typename T>
static int upload(lsm_ctx *ctx, T **dst, const T *src, size_t n)
{
    *dst = nullptr;
    if (n == 0) n = 1;  // keep pointers non-null
    LSM_CUDA(ctx, cudaMalloc((void **)dst, n * sizeof(T)));
    if (src) LSM_CUDA(ctx, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return LSM_OK;
}

// Pageable host buffers for one kernel stage: pieces of the batch flow through a ring of four pinned slots (filled and emptied by
// the ctx's host copy threads) and double-buffered device staging on the two launch lanes; defined with the copy pool below.
typedef std::function<int(const void *d_in, void *d_out, int n, cudaStream_t st, long long off)> lsm_piece_fn;
static int host_ring_staged(lsm_ctx *ctx, int B, int piece, size_t in_per, size_t out_per, const void *h_in, void *h_out,
                            int lanes, const lsm_piece_fn &launch);
static bool is_pageable(const void *p);

// ------------------------------------------------------------------------------------ stage 1
extern "C" int lsm_frontend_create(lsm_ctx *ctx, const lsm_frontend_params *p, const void *h_table,
                                   const int32_t *h_zoom_i0, const double *h_zoom_f, lsm_frontend **out)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!p || !h_table || !out) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_frontend_create: null argument");
    *out = nullptr;
    if (p->channels <= 0 || p->channels > 1024) LSM_FAIL(ctx, LSM_ERR_INVALID, "channels %d outside 1..1024", p->channels);
    if (p->n_thresholds <= 0 || p->n_thresholds > 8) LSM_FAIL(ctx, LSM_ERR_INVALID, "n_thresholds %d outside 1..8", p->n_thresholds);
    if (p->redundancy <= 0 || p->n_bins <= 1 || p->n_samples <= 0) LSM_FAIL(ctx, LSM_ERR_INVALID, "bad redundancy/n_bins/n_samples");
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    lsm_frontend *fe = new (std::nothrow) lsm_frontend();
    if (!fe) LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
    fe->p = *p;
    int rc = LSM_OK;
    if (p->kind == LSM_FILTERBANK_GAMMATONE) {
        // the kernel keeps exactly three windows open: 2*hop < nwin <= 3*hop (reference: 400 / 160)
        if (!(p->hop > 0 && 2 * p->hop < p->nwin && p->nwin <= 3 * p->hop && p->nwin <= p->n_samples)) {
            delete fe;
            LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "gammatone window/hop %d/%d: need 2*hop < nwin <= 3*hop", p->nwin, p->hop);
        }
        if (p->channels > 256) { delete fe; LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "gammatone channels %d > 256", p->channels); }
        const double *t = (const double *)h_table;
        for (int c = 0; c < p->channels; ++c)
            if (t[10 * c + 5] != 0.0) { delete fe; LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "gammatone design with A2 != 0"); }
        fe->ncols = 1 + (p->n_samples - p->nwin) / p->hop;
        for (int c = 0; c < p->channels; ++c) {
            // same expressions as gt_filter_fast evaluates on the device
            const double *r = t + 10 * c;
            for (int k = 0; k < 4; ++k) fe->h_lane_coef[c][k] = r[1 + k] / r[0];
            fe->h_lane_coef[c][4] = -(r[7] / r[6]);
            fe->h_lane_coef[c][5] = -(r[8] / r[6]);
        }
        rc = upload(ctx, &fe->d_coefs, t, (size_t)p->channels * 10);
        // worst-case distance between the speculative and the reference-order arrangement of this design (error_bound.cu)
        if (rc == LSM_OK && lsm_gammatone_error_bound(t, p->channels, p->n_samples, fe->h_kappa) != LSM_OK) {
            rc = LSM_ERR_INVALID;
            snprintf(ctx->err, sizeof(ctx->err), "gammatone design: the error bound of the speculative filter is not finite");
        }
        if (rc == LSM_OK) rc = upload(ctx, &fe->d_kappa, fe->h_kappa, (size_t)p->channels);
        if (rc == LSM_OK) rc = lsm_gammatone_grid(ctx, p, &fe->grid);
        if (rc == LSM_OK) rc = upload<double>(ctx, &fe->d_scratch, nullptr, 2 * (size_t)fe->grid * fe->ncols * p->channels);   // two slots
        if (rc == LSM_OK) rc = upload<int>(ctx, &fe->d_counters, nullptr, 128 + 64 * 256);
        if (rc == LSM_OK && cudaMemset(fe->d_counters, 0, (128 + 64 * 256) * sizeof(int)) != cudaSuccess) rc = LSM_ERR_CUDA;
        fe->mode = getenv("LSM_EXACT_FILTER") ? LSM_FILTER_EXACT : LSM_FILTER_SPECULATIVE;
    } else if (p->kind == LSM_FILTERBANK_MEL) {
        if (p->n_fft <= 0 || (p->n_fft & (p->n_fft - 1)) || p->mel_hop <= 0) { delete fe; LSM_FAIL(ctx, LSM_ERR_INVALID, "mel n_fft must be a power of two"); }
        fe->ncols = 1 + p->n_samples / p->mel_hop;
        rc = lsm_mel_create(ctx, fe, (const float *)h_table);
    } else {
        delete fe;
        LSM_FAIL(ctx, LSM_ERR_INVALID, "unknown filterbank kind %d", p->kind);
    }
    if (rc == LSM_OK && fe->ncols != p->n_bins) {
        if (!h_zoom_i0 || !h_zoom_f) { rc = LSM_ERR_INVALID; snprintf(ctx->err, sizeof(ctx->err), "zoom table required: %d columns -> %d bins", fe->ncols, p->n_bins); }
        else {
            for (int j = 0; j < p->n_bins && rc == LSM_OK; ++j)
                if (h_zoom_i0[j] < 0 || h_zoom_i0[j] >= fe->ncols) { rc = LSM_ERR_INVALID; snprintf(ctx->err, sizeof(ctx->err), "zoom index out of range"); }
            if (rc == LSM_OK) rc = upload(ctx, &fe->d_zoom_i0, h_zoom_i0, (size_t)p->n_bins);
            if (rc == LSM_OK) rc = upload(ctx, &fe->d_zoom_f, h_zoom_f, (size_t)p->n_bins);
        }
    }
    if (rc != LSM_OK) { lsm_frontend_destroy(fe); return rc; }
    *out = fe;
    return LSM_OK;
}

extern "C" void lsm_frontend_destroy(lsm_frontend *fe)
{
    if (!fe) return;
    for (int k = 0; k < 2; ++k) {
        if (fe->slot_valid[k]) cudaEventSynchronize(fe->ev_slot[k]);
        if (fe->ev_slot[k]) cudaEventDestroy(fe->ev_slot[k]);
    }
    cudaFree(fe->d_energy); cudaFree(fe->d_rerun); cudaFree(fe->d_kappa); cudaFree(fe->d_xmax); cudaFree(fe->d_pipe_sync);
    cudaFree(fe->d_coefs); cudaFree(fe->d_zoom_i0); cudaFree(fe->d_zoom_f); cudaFree(fe->d_scratch); cudaFree(fe->d_counters);
    lsm_mel_destroy(fe);
    delete fe;
}

extern "C" int lsm_frontend_set_mode(lsm_ctx *ctx, lsm_frontend *fe, int mode, double delta_db)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || (mode != LSM_FILTER_EXACT && mode != LSM_FILTER_SPECULATIVE) || !(delta_db >= 0.0) || !(delta_db < 1e300))
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_frontend_set_mode: mode must be 0 or 1 and 0 <= delta_db < 1e300");
    fe->mode = mode;
    fe->spec_delta = delta_db;
    return LSM_OK;
}

extern "C" int lsm_frontend_set_bound_scale(lsm_ctx *ctx, lsm_frontend *fe, double scale)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || !(scale >= 0.0) || !(scale < 1e30)) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_frontend_set_bound_scale: 0 <= scale < 1e30");
    fe->bound_scale = scale;
    return LSM_OK;
}

extern "C" int lsm_frontend_audit(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int32_t B, double *d_out)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || B < 0 || (B > 0 && (!d_pcm || !d_out))) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_frontend_audit: bad argument");
    if (fe->p.kind != LSM_FILTERBANK_GAMMATONE) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "lsm_frontend_audit: gammatone front ends only");
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    return lsm_launch_audit(ctx, fe, d_pcm, B, d_out, ctx->stream);
}

extern "C" int lsm_frontend_reruns(lsm_ctx *ctx, lsm_frontend *fe, int64_t *h_out, int reset)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || !h_out) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_frontend_reruns: null argument");
    *h_out = 0;
    if (fe->p.kind != LSM_FILTERBANK_GAMMATONE) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = lsm_frontend_wait_idle(ctx, fe);
    if (rc != LSM_OK) return rc;
    int v = 0;
    LSM_CUDA(ctx, cudaMemcpy(&v, fe->d_counters + 64, sizeof(int), cudaMemcpyDeviceToHost));
    if (reset) LSM_CUDA(ctx, cudaMemset(fe->d_counters + 64, 0, sizeof(int)));
    *h_out = v;
    return LSM_OK;
}

extern "C" int lsm_frontend_mel_tables(lsm_ctx *ctx, lsm_frontend *fe, const double *h_window, const double *h_tw,
                                       const double *h_tw2)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || !h_window || !h_tw || !h_tw2) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_frontend_mel_tables: null argument");
    if (fe->p.kind != LSM_FILTERBANK_MEL) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_frontend_mel_tables: not a mel front end");
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    return lsm_mel_set_tables(ctx, fe, h_window, h_tw, h_tw2);
}

// The per-CTA scratch planes (and the work-counter ring) of a front end are shared by all its launches.  Launches on one
// stream are ordered by the stream; when the stream changes (torch's stream for the device-pointer calls, the ctx's
// own stream for the *_host calls) the new launch first waits for the previous one.
int lsm_frontend_order_before(lsm_ctx *ctx, lsm_frontend *fe, cudaStream_t st, int slot)
{
    for (int k = 0; k < 2; ++k)
        if ((slot < 0 || slot == k) && fe->slot_valid[k] && fe->slot_stream[k] != st)
            LSM_CUDA(ctx, cudaStreamWaitEvent(st, fe->ev_slot[k], 0));
    return LSM_OK;
}

int lsm_frontend_order_after(lsm_ctx *ctx, lsm_frontend *fe, cudaStream_t st, int slot)
{
    for (int k = 0; k < 2; ++k) {
        if (slot >= 0 && slot != k) continue;
        if (!fe->ev_slot[k]) LSM_CUDA(ctx, cudaEventCreateWithFlags(&fe->ev_slot[k], cudaEventDisableTiming));
        LSM_CUDA(ctx, cudaEventRecord(fe->ev_slot[k], st));
        fe->slot_stream[k] = st;
        fe->slot_valid[k] = 1;
    }
    return LSM_OK;
}

int lsm_frontend_wait_idle(lsm_ctx *ctx, lsm_frontend *fe)
{
    for (int k = 0; k < 2; ++k)
        if (fe->slot_valid[k]) LSM_CUDA(ctx, cudaEventSynchronize(fe->ev_slot[k]));
    return LSM_OK;
}

static int frontend_launch(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes,
                           double *d_spec, cudaStream_t st)
{
    if (fe->p.kind == LSM_FILTERBANK_GAMMATONE) return lsm_launch_gammatone(ctx, fe, d_pcm, B, d_spikes, d_spec, st);
    return lsm_launch_mel(ctx, fe, d_pcm, B, d_spikes, d_spec, st);
}

extern "C" int lsm_frontend_encode(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int32_t B,
                                   uint8_t *d_spikes, double *d_spec_norm_or_null)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || B < 0 || (B > 0 && (!d_pcm || !d_spikes))) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_frontend_encode: bad argument");
    if (B == 0) return LSM_OK;   // empty batch: nothing to do (create_dataset.py:164-166 prints and returns)
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    return frontend_launch(ctx, fe, d_pcm, B, d_spikes, d_spec_norm_or_null, ctx->stream);
}

extern "C" int lsm_frontend_encode_host(lsm_ctx *ctx, lsm_frontend *fe, const float *h_pcm, int32_t B,
                                        uint8_t *h_spikes)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || B < 0 || (B > 0 && (!h_pcm || !h_spikes))) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_frontend_encode_host: bad argument");
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t pcm_bytes = (size_t)B * fe->p.n_samples * sizeof(float);
    const size_t spk_bytes = (size_t)B * fe->p.channels * fe->p.redundancy * fe->p.n_bins * fe->p.n_thresholds;
    void *d_pcm, *d_spk;
    int rc;
    if (B >= 512 && is_pageable(h_pcm) && is_pageable(h_spikes) && !getenv("LSM_NO_PAGEABLE_RING"))
        return host_ring_staged(ctx, B, 768, pcm_bytes / B, spk_bytes / B, h_pcm, h_spikes, 2,
                                [&](const void *d_in, void *d_out, int n, cudaStream_t st, long long) {
                                    return frontend_launch(ctx, fe, (const float *)d_in, n, (uint8_t *)d_out, nullptr, st);
                                });
    if ((rc = lsm_stage_device(ctx, 0, pcm_bytes, &d_pcm)) != LSM_OK) return rc;
    if ((rc = lsm_stage_device(ctx, 1, spk_bytes, &d_spk)) != LSM_OK) return rc;
    LSM_CUDA(ctx, cudaMemcpyAsync(d_pcm, h_pcm, pcm_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = frontend_launch(ctx, fe, (const float *)d_pcm, B, (uint8_t *)d_spk, nullptr, ctx->stream)) != LSM_OK) return rc;
    LSM_CUDA(ctx, cudaMemcpyAsync(h_spikes, d_spk, spk_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    LSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSM_OK;
}

// ------------------------------------------------------------------------------------ host design helpers
// Slaney's ERBSpace / MakeERBFilters as published in gammatone==1.0.3 (call site create_dataset.py:51-58); the same
// formulas as lsm_speech_classifier_b200/filterbank.py, for C callers.  Transcendentals come from the host libm, so
// the last bit may differ from numpy's: pass the same table to every implementation you want to agree bit for bit.
extern "C" int lsm_gammatone_design(double fs, int32_t channels, double f_min, double *h_out)
{
    if (!h_out || channels <= 0 || fs <= 0 || f_min <= 0 || f_min >= fs / 2) return LSM_ERR_INVALID;
    const double ear_q = 9.26449, min_bw = 24.7, pi = 3.14159265358979323846;
    const double c = ear_q * min_bw, hi = fs / 2, T = 1 / fs;
    const double rt_pos = sqrt(3 + pow(2.0, 1.5)), rt_neg = sqrt(3 - pow(2.0, 1.5));
    for (int i = 1; i <= channels; ++i) {
        const double cf = -c + exp(((double)i / channels) * (-log(hi + c) + log(f_min + c))) * (hi + c);
        const double erb = cf / ear_q + min_bw;
        const double B = 1.019 * 2 * pi * erb;
        const double arg = 2 * cf * pi * T;
        const std::complex<double> vec = std::exp(std::complex<double>(0, 2 * arg));
        const double B1 = -2 * cos(arg) / exp(B * T), B2 = exp(-2 * B * T);
        const double common = -T * exp(-(B * T));
        const double k[4] = {cos(arg) + rt_pos * sin(arg), cos(arg) - rt_pos * sin(arg),
                             cos(arg) + rt_neg * sin(arg), cos(arg) - rt_neg * sin(arg)};
        const std::complex<double> g = std::exp(std::complex<double>(-B * T, arg));
        const std::complex<double> q = T * exp(B * T) / (-1 / exp(B * T) + 1.0 + vec * (1 - exp(B * T)));
        const double gain = std::abs((vec - g * k[0]) * (vec - g * k[1]) * (vec - g * k[2]) * (vec - g * k[3]) * q * q * q * q);
        double *row = h_out + 10 * (size_t)(channels - i);      // flipud: row 0 = lowest centre frequency
        row[0] = T; row[1] = common * k[0]; row[2] = common * k[1]; row[3] = common * k[2]; row[4] = common * k[3];
        row[5] = 0.0; row[6] = 1.0; row[7] = B1; row[8] = B2; row[9] = gain;
    }
    return LSM_OK;
}

extern "C" int lsm_zoom_table(int32_t n_in, int32_t n_out, int32_t *h_i0, double *h_f)
{
    if (!h_i0 || !h_f || n_in < 2 || n_out < 2) return LSM_ERR_INVALID;
    const double zz = (double)(n_in - 1) / (double)(n_out - 1);
    for (int j = 0; j < n_out; ++j) {
        const double cc = (double)j * zz, fl = floor(cc);
        h_i0[j] = (int32_t)fl;
        h_f[j] = cc - fl;
    }
    return LSM_OK;
}

// ------------------------------------------------------------------------------------ encoder alone
namespace {
struct EncArgs { double thr[8], lower[8]; float thr32[8], lower32[8]; };

template <typename T>
__global__ void hysteresis_kernel(const T *__restrict__ spec, int rows, int n_bins, int K, int R, EncArgs e,
                                  uint8_t *__restrict__ spikes)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;   // (utterance, channel)
    if (row >= rows) return;
    const T *v = spec + (size_t)row * n_bins;
    uint8_t *out0 = spikes + (size_t)row * R * n_bins * K;
    unsigned on = 0;
    for (int j = 0; j < n_bins; ++j) {
        const T x = v[j];
        for (int k = 0; k < K; ++k) {
            const bool is_on = (on >> k) & 1u;
            bool rise, fall;
            if (sizeof(T) == 4) { rise = (float)x > e.thr32[k]; fall = (float)x < e.lower32[k]; }
            else { rise = (double)x > e.thr[k]; fall = (double)x < e.lower[k]; }
            if (!is_on && rise) on |= 1u << k;
            else if (is_on && fall) on &= ~(1u << k);
        }
        for (int r = 0; r < R; ++r)
            for (int k = 0; k < K; ++k) out0[(size_t)r * n_bins * K + (size_t)j * K + k] = (on >> k) & 1u;
    }
}
}  // namespace

extern "C" int lsm_hysteresis_encode(lsm_ctx *ctx, const void *d_spec, int32_t is_f32, int32_t B, int32_t C, int32_t n_bins,
                                     const double *h_thr_desc, const double *h_lower, int32_t K, int32_t redundancy,
                                     uint8_t *d_spikes)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (B < 0 || C <= 0 || n_bins <= 0 || K <= 0 || K > 8 || redundancy <= 0 || !h_thr_desc || !h_lower ||
        (B > 0 && (!d_spec || !d_spikes)))
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_hysteresis_encode: bad argument");
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    EncArgs e = {};
    for (int k = 0; k < K; ++k) {
        e.thr[k] = h_thr_desc[k]; e.lower[k] = h_lower[k];
        e.thr32[k] = (float)h_thr_desc[k]; e.lower32[k] = (float)h_lower[k];
    }
    const int rows = B * C;
    const int threads = 128, blocks = (rows + threads - 1) / threads;
    if (is_f32) hysteresis_kernel<float><<<blocks, threads, 0, ctx->stream>>>((const float *)d_spec, rows, n_bins, K, redundancy, e, d_spikes);
    else hysteresis_kernel<double><<<blocks, threads, 0, ctx->stream>>>((const double *)d_spec, rows, n_bins, K, redundancy, e, d_spikes);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return LSM_OK;
}

// ------------------------------------------------------------------------------------ stage 2+3
// h_w_q: integer weights (multiples of 2^-w_shift), or h_w_val: fp64 weights of a strict reservoir - exactly one of the two
static int reservoir_create_impl(lsm_ctx *ctx, const lsm_reservoir_params *p,
                                 const int32_t *h_w_rowptr, const int32_t *h_w_col, const int32_t *h_w_q, const double *h_w_val,
                                 const int32_t *h_in_rowptr, const int32_t *h_in_col, const double *h_in_val,
                                 const double *h_leak, const int32_t *h_out_idx, lsm_reservoir **out)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!p || !h_w_rowptr || !h_in_rowptr || !h_leak || !h_out_idx || !out)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_create: null argument");
    *out = nullptr;
    const int N = p->num_neurons;
    if (N <= 0 || p->num_inputs <= 0 || p->num_steps <= 0 || p->n_out < 0 || p->n_out > N || p->refractory < 0 ||
        p->w_shift < 0 || p->w_shift > 52)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_create: bad dimensions");
    const int64_t nnz = h_w_rowptr[N];
    const bool strict = h_w_val != nullptr;
    if (nnz > 0 && (!h_w_col || (!h_w_q && !h_w_val))) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_create: null weights");
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    lsm_reservoir *res = new (std::nothrow) lsm_reservoir();
    if (!res) LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
    res->p = *p;
    int npt, threads;
    lsm_reservoir_geometry(N, &npt, &threads, &res->n_pad);
    if (threads > 1024) { delete res; LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "num_neurons %d > 16384 not supported by the event-driven kernel", N); }
    const int n_pad = res->n_pad;
    const int64_t nin = h_in_rowptr[N];
    for (int64_t q = 0; q < nin; ++q)
        if (h_in_col[q] < 0 || h_in_col[q] >= p->num_inputs) { delete res; LSM_FAIL(ctx, LSM_ERR_INVALID, "input row out of range"); }
    std::vector<int32_t> in_row(n_pad, -1);
    for (int i = 0; i < N; ++i) {
        const int d = h_in_rowptr[i + 1] - h_in_rowptr[i];
        if (d > res->max_in_per_neuron) res->max_in_per_neuron = d;
        if (d == 1) in_row[i] = h_in_col[h_in_rowptr[i]];
        else if (d > 1) in_row[i] = -2;
    }
    // lean kernel variant (reservoir_core.cuh): uniform leak, uniform input gain (and no -0.0), every driven neuron has exactly
    // one input row and no row drives two neurons, the rows fit the "internal neuron 8r" slots, theta > 0, refractory period
    // fits a 4-bit counter, and the input gain is a multiple of the weight quantum (so the folded constant is exact)
    res->lean = res->max_in_per_neuron <= 1 && !getenv("LSM_NO_LEAN") && !strict;     // strict reservoirs keep the caller's neuron order
    res->w64 = strict ? 1 : 0;
    res->leak0 = h_leak[0];
    res->gain0 = nin > 0 ? h_in_val[0] + 0.0 : 0.0;
    for (int i = 1; i < N && res->lean; ++i)
        if (memcmp(&h_leak[i], &h_leak[0], sizeof(double)) != 0) res->lean = 0;
    for (int64_t q = 0; q < nin && res->lean; ++q)
        if (memcmp(&h_in_val[q], &h_in_val[0], sizeof(double)) != 0) res->lean = 0;
    if (res->lean) {
        std::vector<char> seen(p->num_inputs, 0);
        for (int i = 0; i < N && res->lean; ++i)
            if (in_row[i] >= 0) { if (seen[in_row[i]]) res->lean = 0; seen[in_row[i]] = 1; }
        // the lean kernel treats internal slot 8r as driven by row r: every row must drive a neuron
        for (int r = 0; r < p->num_inputs && res->lean; ++r)
            if (!seen[r]) res->lean = 0;
        const double g = ldexp(res->gain0, p->w_shift);
        if (p->num_inputs > n_pad / 8 || !(p->theta > 0.0) || p->refractory > 15 || g != floor(g) ||
            fabs(res->gain0) > ldexp(1.0, 31 - p->w_shift))
            res->lean = 0;
    }
    // neuron relabelling: perm[external] = internal.  Lean: the neuron driven by input row r becomes internal 8r, the others
    // fill the remaining slots in ascending order; otherwise the identity.
    std::vector<int32_t> perm(N), ext_id(n_pad, N);
    if (res->lean) {
        std::vector<char> taken(n_pad, 0);
        for (int i = 0; i < N; ++i)
            if (in_row[i] >= 0) { perm[i] = 8 * in_row[i]; taken[perm[i]] = 1; }
        int next = 0;
        for (int i = 0; i < N; ++i)
            if (in_row[i] < 0) {
                while (taken[next]) ++next;
                perm[i] = next; taken[next] = 1;
            }
    } else {
        for (int i = 0; i < N; ++i) perm[i] = i;
    }
    for (int i = 0; i < N; ++i) ext_id[perm[i]] = i;
    res->zero_row = res->lean ? n_pad : N;
    res->skip_dead_time = (p->theta > 0.0 && !getenv("LSM_NO_DEAD_TIME_SKIP")) ? 1 : 0;
    for (int i = 0; i < N; ++i)
        if (!(h_leak[i] >= 0.0 && h_leak[i] <= 1.0)) res->skip_dead_time = 0;
    res->hi_magic = (1075 - p->w_shift) << 20;
    res->c_off = ldexp(1.0, 52 - p->w_shift) + ldexp(1.0, 31 - p->w_shift);
    res->c_on = res->c_off - res->gain0;
    // dense presynaptic-major plane: wt[j][i] = weight of j -> i (internal labels); one extra all-zero row pads spike lists
    std::vector<int32_t> wt(strict ? 1 : (size_t)(res->zero_row + 1) * n_pad, 0);
    std::vector<double> wt64(strict ? (size_t)N * n_pad : 0, 0.0);
    std::vector<int64_t> rowabs(N, 0);
    for (int i = 0; i < N; ++i) {
        int prev = -1;
        for (int64_t q = h_w_rowptr[i]; q < h_w_rowptr[i + 1]; ++q) {
            const int j = h_w_col[q];
            if (j < 0 || j >= N) { delete res; LSM_FAIL(ctx, LSM_ERR_INVALID, "presynaptic index out of range"); }
            if (strict) {
                // one edge per (post, pre) pair, ascending inside a row: the order of the sum is part of the definition
                if (j <= prev) { delete res; LSM_FAIL(ctx, LSM_ERR_INVALID, "row %d: presynaptic indices must be strictly ascending", i); }
                prev = j;
                wt64[(size_t)j * n_pad + i] = h_w_val[q];
                continue;
            }
            wt[(size_t)perm[j] * n_pad + perm[i]] += h_w_q[q];
            rowabs[i] += h_w_q[q] < 0 ? -(int64_t)h_w_q[q] : (int64_t)h_w_q[q];
        }
        if (rowabs[i] >= (1ll << 31)) { delete res; LSM_FAIL(ctx, LSM_ERR_INVALID, "row %d: sum of |weights| overflows the exact int32 accumulator", i); }
    }
    std::vector<int32_t> out_slot(n_pad, -1);
    for (int o = 0; o < p->n_out; ++o) {
        if (h_out_idx[o] < 0 || h_out_idx[o] >= N) { delete res; LSM_FAIL(ctx, LSM_ERR_INVALID, "output index out of range"); }
        out_slot[perm[h_out_idx[o]]] = o;
    }
    // dense arm (reservoir_dense.cu): per-neuron tables in internal order and the number of 8-bit digit planes the weights need
    res->h_dense_in_row = (int32_t *)malloc(sizeof(int32_t) * n_pad);
    res->h_dense_gain = (double *)malloc(sizeof(double) * n_pad);
    res->h_dense_leak = (double *)malloc(sizeof(double) * n_pad);
    res->h_dense_out_int = (int32_t *)malloc(sizeof(int32_t) * (p->n_out > 0 ? p->n_out : 1));
    if (!res->h_dense_in_row || !res->h_dense_gain || !res->h_dense_leak || !res->h_dense_out_int) {
        lsm_reservoir_destroy(res);
        LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
    }
    for (int i = 0; i < n_pad; ++i) { res->h_dense_in_row[i] = -1; res->h_dense_gain[i] = 0.0; res->h_dense_leak[i] = 0.0; }
    for (int i = 0; i < N; ++i) {
        res->h_dense_in_row[perm[i]] = in_row[i];
        res->h_dense_gain[perm[i]] = in_row[i] >= 0 ? h_in_val[h_in_rowptr[i]] : 0.0;
        res->h_dense_leak[perm[i]] = h_leak[i];
    }
    for (int o = 0; o < p->n_out; ++o) res->h_dense_out_int[o] = perm[h_out_idx[o]];
    {
        int32_t lo = 0, hi = 0;
        for (size_t q = 0; q < wt.size(); ++q) { if (wt[q] < lo) lo = wt[q]; if (wt[q] > hi) hi = wt[q]; }
        if (lo < 0 || hi >= (1 << 24)) { res->dense_planes = 4; res->dense_top_signed = 1; }
        else res->dense_planes = hi >= (1 << 16) ? 3 : (hi >= (1 << 8) ? 2 : 1);
    }
    if (strict) res->dense_planes = 0;                    // the dense arm contracts integer digit planes
    int rc = upload(ctx, &res->d_wt, wt.data(), wt.size());
    if (rc == LSM_OK && strict) rc = upload(ctx, &res->d_wt64, wt64.data(), wt64.size());
    if (rc == LSM_OK) rc = upload(ctx, &res->d_in_rowptr, h_in_rowptr, (size_t)N + 1);
    if (rc == LSM_OK) rc = upload(ctx, &res->d_in_col, h_in_col, (size_t)nin);
    if (rc == LSM_OK) rc = upload(ctx, &res->d_in_val, h_in_val, (size_t)nin);
    if (rc == LSM_OK) rc = upload(ctx, &res->d_leak, h_leak, (size_t)N);
    if (rc == LSM_OK) rc = upload(ctx, &res->d_out_slot, out_slot.data(), out_slot.size());
    if (rc == LSM_OK) rc = upload(ctx, &res->d_in_row, in_row.data(), in_row.size());
    if (rc == LSM_OK) rc = upload(ctx, &res->d_ext_id, ext_id.data(), ext_id.size());
    if (rc != LSM_OK) { lsm_reservoir_destroy(res); return rc; }
    *out = res;
    return LSM_OK;
}

extern "C" int lsm_reservoir_create(lsm_ctx *ctx, const lsm_reservoir_params *p,
                                    const int32_t *h_w_rowptr, const int32_t *h_w_col, const int32_t *h_w_q,
                                    const int32_t *h_in_rowptr, const int32_t *h_in_col, const double *h_in_val,
                                    const double *h_leak, const int32_t *h_out_idx, lsm_reservoir **out)
{
    if (ctx && h_w_rowptr && p && p->num_neurons > 0 && h_w_rowptr[p->num_neurons] > 0 && !h_w_q)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_create: null weights");
    return reservoir_create_impl(ctx, p, h_w_rowptr, h_w_col, h_w_q, nullptr, h_in_rowptr, h_in_col, h_in_val, h_leak, h_out_idx, out);
}

extern "C" int lsm_reservoir_create_f64(lsm_ctx *ctx, const lsm_reservoir_params *p,
                                        const int32_t *h_w_rowptr, const int32_t *h_w_col, const double *h_w_val,
                                        const int32_t *h_in_rowptr, const int32_t *h_in_col, const double *h_in_val,
                                        const double *h_leak, const int32_t *h_out_idx, lsm_reservoir **out)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!h_w_val) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_create_f64: null weights");
    return reservoir_create_impl(ctx, p, h_w_rowptr, h_w_col, nullptr, h_w_val, h_in_rowptr, h_in_col, h_in_val, h_leak, h_out_idx, out);
}

extern "C" void lsm_reservoir_destroy(lsm_reservoir *res)
{
    if (!res) return;
    cudaFree(res->d_wt); cudaFree(res->d_wt64); cudaFree(res->d_in_rowptr); cudaFree(res->d_in_col); cudaFree(res->d_in_val);
    cudaFree(res->d_leak); cudaFree(res->d_out_slot); cudaFree(res->d_in_row); cudaFree(res->d_ext_id);
    free(res->h_dense_in_row); free(res->h_dense_gain); free(res->h_dense_leak); free(res->h_dense_out_int);
    lsm_dense_ws_free(res->dense);
    delete res;
}

extern "C" int lsm_reservoir_set_gather(lsm_ctx *ctx, lsm_reservoir *res, double *const *d_gather, int32_t n, int64_t row0)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!res || n < 0 || n > 8 || (n > 0 && !d_gather) || row0 < 0) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_set_gather: bad argument (at most 8 destinations)");
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int k = 0; k < n; ++k) {
        if (!d_gather[k]) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_set_gather: null destination %d", k);
        // a destination in another GPU's memory (IPC-mapped): kernels of this device may only store there once peer access is on
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, d_gather[k]) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
            cudaGetLastError();
            LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_set_gather: destination %d is not device memory", k);
        }
        if (at.device != ctx->device) {
            int can = 0;
            LSM_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, at.device));
            if (!can) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "lsm_reservoir_set_gather: device %d cannot access device %d", ctx->device, at.device);
        }
    }
    if (n > 0) {
        // peer access from this device to every other visible device (idempotent; IPC-mapped destinations report their owner)
        int ndev = 0;
        LSM_CUDA(ctx, cudaGetDeviceCount(&ndev));
        for (int d = 0; d < ndev; ++d) {
            if (d == ctx->device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ctx->device, d) != cudaSuccess || !can) { cudaGetLastError(); continue; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(d, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); LSM_FAIL(ctx, LSM_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) -> %s", d, cudaGetErrorString(e)); }
            cudaGetLastError();
        }
    }
    res->n_gather = n;
    res->gather_row0 = row0;
    for (int k = 0; k < 8; ++k) res->gather_out[k] = k < n ? d_gather[k] : nullptr;
    return LSM_OK;
}

// ------------------------------------------------------------------------------------ peer buffers (fused all-gather)
// Gather matrices that other ranks' kernels store into: plain cudaMalloc memory exported / imported through CUDA IPC.
extern "C" int lsm_peer_buffer_create(lsm_ctx *ctx, int64_t bytes, void **d_ptr, void *h_handle64)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!d_ptr || !h_handle64 || bytes <= 0) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_peer_buffer_create: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    *d_ptr = nullptr;
    if (cudaMalloc(d_ptr, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); LSM_FAIL(ctx, LSM_ERR_NOMEM, "cudaMalloc(%lld) for a peer buffer failed", (long long)bytes); }
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) {
        cudaFree(*d_ptr); *d_ptr = nullptr; cudaGetLastError();
        LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    }
    memcpy(h_handle64, &h, 64);
    return LSM_OK;
}

extern "C" int lsm_peer_buffer_open(lsm_ctx *ctx, const void *h_handle64, void **d_ptr)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!d_ptr || !h_handle64) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_peer_buffer_open: bad argument");
    // the mapping must belong to the context this ctx's kernels run in: open it with the ctx's device current
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, 64);
    *d_ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e)); }
    return LSM_OK;
}

extern "C" int lsm_peer_buffer_close(lsm_ctx *ctx, void *d_ptr)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!d_ptr) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    LSM_CUDA(ctx, cudaDeviceSynchronize());
    LSM_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return LSM_OK;
}

extern "C" int lsm_peer_buffer_destroy(lsm_ctx *ctx, void *d_ptr)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!d_ptr) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    LSM_CUDA(ctx, cudaDeviceSynchronize());
    LSM_CUDA(ctx, cudaFree(d_ptr));
    return LSM_OK;
}

// stage 2+3 by the arm lsm_reservoir_set_mode selected (event-driven by default)
static int launch_reservoir_by_mode(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int B, uint32_t feature_mask,
                                    int nan_to_num, double *d_features, uint8_t *d_raster, cudaStream_t st)
{
    if (res->mode == LSM_RESERVOIR_DENSE)
        return lsm_launch_reservoir_dense(ctx, res, d_spikes, B, feature_mask, nan_to_num, d_features, d_raster, st);
    return lsm_launch_reservoir(ctx, res, d_spikes, B, feature_mask, nan_to_num, d_features, d_raster, st);
}

extern "C" int lsm_reservoir_run(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int32_t B,
                                 uint32_t feature_mask, int32_t nan_to_num, double *d_features,
                                 uint8_t *d_raster_or_null)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!res || B < 0 || (B > 0 && !d_spikes)) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_run: bad argument");
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_reservoir_by_mode(ctx, res, d_spikes, B, feature_mask, nan_to_num, d_features, d_raster_or_null, ctx->stream);
}

extern "C" int lsm_reservoir_set_mode(lsm_ctx *ctx, lsm_reservoir *res, int32_t mode)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!res || (mode != LSM_RESERVOIR_EVENT && mode != LSM_RESERVOIR_DENSE)) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_set_mode: bad argument");
    if (mode == LSM_RESERVOIR_DENSE)
        if (const char *why = lsm_dense_unsupported(res)) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "dense reservoir arm: %s", why);
    res->mode = mode;
    return LSM_OK;
}

extern "C" int lsm_reservoir_dense_probe(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_s, int32_t B, int32_t *d_acc)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!res || B < 0 || (B > 0 && (!d_s || !d_acc))) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_dense_probe: bad argument");
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    return lsm_reservoir_dense_probe_launch(ctx, res, d_s, B, d_acc, ctx->stream);
}

extern "C" int lsm_reservoir_diagnostics(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int32_t B, int32_t *d_diag)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!res || B < 0 || (B > 0 && (!d_spikes || !d_diag))) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_diagnostics: bad argument");
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    return lsm_launch_reservoir(ctx, res, d_spikes, B, 0u, 0, nullptr, nullptr, ctx->stream, d_diag);
}

extern "C" int lsm_reservoir_run_host(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *h_spikes, int32_t B,
                                      uint32_t feature_mask, int32_t nan_to_num, double *h_features,
                                      uint8_t *h_raster_or_null)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!res || B < 0 || (B > 0 && !h_spikes)) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_run_host: bad argument");
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    const lsm_reservoir_params &p = res->p;
    const int nkeys = __builtin_popcount(feature_mask & 0xFFu);
    const size_t spk_bytes = (size_t)B * p.num_inputs * p.num_steps;
    const size_t feat_bytes = (size_t)B * nkeys * p.n_out * sizeof(double);
    const size_t ras_bytes = (size_t)B * p.num_steps * p.num_neurons;
    void *d_spk = nullptr, *d_feat = nullptr, *d_ras = nullptr;
    int rc;
    if (B >= 512 && h_features && !h_raster_or_null && res->mode == LSM_RESERVOIR_EVENT && is_pageable(h_spikes) && is_pageable(h_features) &&
        !getenv("LSM_NO_PAGEABLE_RING")) {
        // reservoirs whose statistics live in a global slab (one slab per ctx) keep to one lane
        const int lanes = p.num_neurons > 8192 ? 1 : 2;
        return host_ring_staged(ctx, B, 1024, spk_bytes / B, feat_bytes / B, h_spikes, h_features, lanes,
                                [&](const void *d_in, void *d_out, int n, cudaStream_t st, long long off) {
                                    return lsm_launch_reservoir(ctx, res, (const uint8_t *)d_in, n, feature_mask, nan_to_num, (double *)d_out,
                                                                nullptr, st, nullptr, off);
                                });
    }
    if ((rc = lsm_stage_device(ctx, 1, spk_bytes, &d_spk)) != LSM_OK) return rc;
    if (h_features && (rc = lsm_stage_device(ctx, 2, feat_bytes, &d_feat)) != LSM_OK) return rc;
    if (h_raster_or_null && (rc = lsm_stage_device(ctx, 3, ras_bytes, &d_ras)) != LSM_OK) return rc;
    LSM_CUDA(ctx, cudaMemcpyAsync(d_spk, h_spikes, spk_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_reservoir_by_mode(ctx, res, (const uint8_t *)d_spk, B, feature_mask, nan_to_num,
                                       (double *)d_feat, (uint8_t *)d_ras, ctx->stream)) != LSM_OK) return rc;
    if (h_features) LSM_CUDA(ctx, cudaMemcpyAsync(h_features, d_feat, feat_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_raster_or_null) LSM_CUDA(ctx, cudaMemcpyAsync(h_raster_or_null, d_ras, ras_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    LSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSM_OK;
}

// Can the device address this buffer directly?  Pinned/registered host memory (UVA alias) or device/managed memory.
static bool device_visible(const void *h, void **d)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return false; }
    if ((at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) && at.devicePointer) {
        *d = at.devicePointer;
        return true;
    }
    return false;
}


// ------------------------------------------------------------------------------------ host copy workers
// Pageable caller buffers (plain numpy arrays: what create_dataset / extract_all_features style callers hand over) reach the
// zero-copy kernel through a ring of pinned staging buffers; the copies into and out of the ring are split over a few host
// threads (one thread moves ~10 GB/s, the kernel consumes 25 GB/s of PCM), created on first use and owned by the ctx.
struct lsm_copy_pool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    char *dst = nullptr;
    const char *src = nullptr;
    size_t bytes = 0;
    unsigned generation = 0;
    int pending = 0;
    bool stop = false;

    explicit lsm_copy_pool(int n)
    {
        for (int i = 0; i < n; ++i) threads.emplace_back([this, i, n] { run(i, n); });
    }
    ~lsm_copy_pool()
    {
        { std::lock_guard<std::mutex> l(m); stop = true; }
        cv_work.notify_all();
        for (auto &t : threads) t.join();
    }
    void run(int i, int n)
    {
        unsigned seen = 0;
        for (;;) {
            char *d; const char *s_; size_t b;
            {
                std::unique_lock<std::mutex> l(m);
                cv_work.wait(l, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation; d = dst; s_ = src; b = bytes;
            }
            // 4 KiB-aligned slices
            const size_t per = ((b / n) + 4095) & ~(size_t)4095;
            const size_t lo = (size_t)i * per < b ? (size_t)i * per : b, hi = lo + per < b ? lo + per : b;
            if (hi > lo) memcpy(d + lo, s_ + lo, hi - lo);
            {
                std::lock_guard<std::mutex> l(m);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // blocking: returns when all slices have been copied
    void copy(void *d, const void *s_, size_t b)
    {
        if (b < (1u << 20) || threads.empty()) { memcpy(d, s_, b); return; }
        std::unique_lock<std::mutex> l(m);
        dst = (char *)d; src = (const char *)s_; bytes = b; pending = (int)threads.size(); ++generation;
        cv_work.notify_all();
        cv_done.wait(l, [&] { return pending == 0; });
    }
};

void lsm_copy_pool_free(lsm_copy_pool *p) { delete p; }

static lsm_copy_pool *copy_pool(lsm_ctx *ctx)
{
    if (!ctx->copy_pool) {
        int n = (int)std::thread::hardware_concurrency() / 2;
        if (const char *e = getenv("LSM_COPY_THREADS")) n = atoi(e);
        n = n < 1 ? 1 : (n > 8 ? 8 : n);
        ctx->copy_pool = new (std::nothrow) lsm_copy_pool(n);
    }
    return ctx->copy_pool;
}

// ------------------------------------------------------------------------------------ whole path
static cudaStream_t lane_stream(lsm_ctx *ctx, int lane) { return lane == 0 ? ctx->own_stream : ctx->copy_stream[0]; }

static bool is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

static int host_ring_staged(lsm_ctx *ctx, int B, int piece, size_t in_per, size_t out_per, const void *h_in, void *h_out,
                            int lanes, const lsm_piece_fn &launch)
{
    lsm_copy_pool *pool = copy_pool(ctx);
    if (!pool) LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
    if (piece > B) piece = B;
    const size_t in_bytes = ((size_t)piece * in_per + 255) & ~(size_t)255, out_bytes = ((size_t)piece * out_per + 255) & ~(size_t)255;
    const size_t slot_bytes = in_bytes + out_bytes;
    void *ringp, *devp;
    int rc;
    if ((rc = lsm_stage_pinned(ctx, 1, 4 * slot_bytes, &ringp)) != LSM_OK) return rc;
    if ((rc = lsm_stage_device(ctx, 0, 2 * slot_bytes, &devp)) != LSM_OK) return rc;
    const int n_pieces = (B + piece - 1) / piece;
    auto pin_in = [&](int k) { return (char *)ringp + (size_t)(k & 3) * slot_bytes; };
    auto pin_out = [&](int k) { return pin_in(k) + in_bytes; };
    auto dev_in = [&](int k) { return (char *)devp + (size_t)(k % lanes) * slot_bytes; };
    auto dev_out = [&](int k) { return dev_in(k) + in_bytes; };
    auto count = [&](int k) { return k + 1 < n_pieces ? piece : B - k * piece; };
    for (int k = 0; k < n_pieces + 2; ++k) {
        if (k >= 2) {
            const int j = k - 2;
            LSM_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage_free[j & 3]));
            pool->copy((char *)h_out + (size_t)j * piece * out_per, pin_out(j), (size_t)count(j) * out_per);
        }
        if (k < n_pieces) {
            const int n = count(k);
            pool->copy(pin_in(k), (const char *)h_in + (size_t)k * piece * in_per, (size_t)n * in_per);
            cudaStream_t ls = lane_stream(ctx, k % lanes);
            LSM_CUDA(ctx, cudaMemcpyAsync(dev_in(k), pin_in(k), (size_t)n * in_per, cudaMemcpyHostToDevice, ls));
            if ((rc = launch(dev_in(k), dev_out(k), n, ls, (long long)k * piece)) != LSM_OK) return rc;
            LSM_CUDA(ctx, cudaMemcpyAsync(pin_out(k), dev_out(k), (size_t)n * out_per, cudaMemcpyDeviceToHost, ls));
            LSM_CUDA(ctx, cudaEventRecord(ctx->ev_stage_free[k & 3], ls));
        }
    }
    return LSM_OK;
}


// The warp-specialised kernel's units give up (and say so) if a group never completes; surfaced at the synchronous calls.
static int check_pipe_error(lsm_ctx *ctx, lsm_frontend *fe)
{
    if (!fe->d_pipe_sync) return LSM_OK;
    int flag = 0;
    LSM_CUDA(ctx, cudaMemcpy(&flag, fe->d_pipe_sync + 2 * ((size_t)fe->energy_cap / 32 + 8), sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) LSM_FAIL(ctx, LSM_ERR_CUDA, "pipeline kernel: an encoder unit timed out waiting for its filter units");
    return LSM_OK;
}

// Utterances per launch of the warp-specialised kernel: half the batch (two launches share the SMs), in whole 32-utterance
// groups, at most 8192 (the energy planes between the two roles take 100 KB per utterance).
static int lanes_piece(int B)
{
    if (getenv("LSM_PIPE_ONE_PIECE")) return B > 8192 ? 8192 : B;       // experiment: no split across the two lanes
    int n = ((B / 2 + 31) / 32) * 32;
    if (B < 64) n = B;
    return n > 8192 ? 8192 : n;
}

// The warp-specialised kernel on device-resident PCM (float32, or PCM16 when d_pcm is null and fe->next_pcm16 is set), ordered
// after / before the work on `st`.  A launch has one CTA per SM and two launches share an SM, so the batch goes out in pieces
// that alternate between the two launch lanes: the fill and drain phases of one overlap the steady state of the other.
int lsm_pipeline_lanes_device(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B, uint8_t *d_spikes,
                              uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st)
{
    const int L = fe->p.n_samples;
    const size_t spk_per = (size_t)res->p.num_inputs * res->p.num_steps;
    const size_t feat_per = (size_t)__builtin_popcount(feature_mask & 0xFFu) * res->p.n_out;
    const int16_t *pcm16 = d_pcm ? nullptr : fe->next_pcm16;
    cudaStream_t l0 = lane_stream(ctx, 0), l1 = lane_stream(ctx, 1);
    const bool on_lane = st == l0 || st == l1;
    // variant 2 (energy-unit filter role, dynamic unit queue): whole calls on the caller's stream, alternating the two resource
    // lanes from call to call, so that callers with two streams have two launches in flight
    const bool whole_calls = !on_lane && getenv("LSM_PIPELINE") && atoi(getenv("LSM_PIPELINE")) == 2 && !pcm16;
    const int piece = (on_lane || whole_calls) ? (B > 8192 ? 8192 : B) : lanes_piece(B);
    int rc = LSM_OK, used = 0;
    if (!on_lane && !whole_calls && B > piece) LSM_CUDA(ctx, cudaEventRecord(ctx->ev[8], st));
    for (int off = 0, k = 0; off < B && rc == LSM_OK; off += piece, ++k) {
        const int n = B - off < piece ? B - off : piece;
        int lane = k & 1;
        cudaStream_t ls = lane == 0 ? l0 : l1;
        if (on_lane) { lane = st == l1 ? 1 : 0; ls = st; }
        else if (whole_calls) { lane = (int)(fe->slot_next++ & 1u); ls = st; }
        else if (B <= piece) { lane = 0; ls = st; }          // a single small launch stays on the caller's stream
        else if (k < 2) LSM_CUDA(ctx, cudaStreamWaitEvent(ls, ctx->ev[8], 0));
        fe->next_pcm16 = pcm16 ? pcm16 + (size_t)off * L : nullptr;
        rc = lsm_launch_pipeline_lanes(ctx, fe, res, d_pcm ? d_pcm + (size_t)off * L : nullptr, n,
                                       d_spikes ? d_spikes + (size_t)off * spk_per : nullptr, feature_mask, nan_to_num,
                                       d_features + (size_t)off * feat_per, ls, lane, off);
        if (ls != st) used |= 1 << lane;
    }
    fe->next_pcm16 = pcm16;
    if (rc != LSM_OK) return rc;
    for (int lane = 0; lane < 2; ++lane)
        if (used & (1 << lane)) {
            LSM_CUDA(ctx, cudaEventRecord(ctx->ev[9 + lane], lane == 0 ? l0 : l1));
            LSM_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev[9 + lane], 0));
        }
    return LSM_OK;
}

// Host buffers through the warp-specialised kernel: the batch goes out in pieces that alternate between the two launch lanes,
// each piece copied to the device by the copy engine (the kernel reads every sample sixteen times, from L2), filtered, and
// its feature rows written straight to the host buffer when the device can address it (pinned), else copied back.
// only_lane < 0: both lanes, then wait; 0 / 1: that lane only, asynchronous.
static int pipeline_lanes_host(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const void *h_pcm, bool i16, int B,
                               uint32_t feature_mask, int nan_to_num, double *h_features, uint8_t *h_spikes_or_null, int only_lane,
                               bool warp_specialised = true)
{
    const int L = fe->p.n_samples;
    const size_t esz = i16 ? sizeof(int16_t) : sizeof(float);
    const size_t spk_per = (size_t)res->p.num_inputs * res->p.num_steps;
    const size_t feat_per = (size_t)__builtin_popcount(feature_mask & 0xFFu) * res->p.n_out;
    void *dv_feat = nullptr;
    const bool feat_direct = device_visible(h_features, &dv_feat);
    const int piece = only_lane >= 0 ? (B > 8192 ? 8192 : B) : lanes_piece(B);
    int rc;
    for (int off = 0, k = 0; off < B; off += piece, ++k) {
        const int n = B - off < piece ? B - off : piece;
        const int lane = only_lane >= 0 ? only_lane : (k & 1);
        cudaStream_t ls = lane_stream(ctx, lane);
        void *d_in = nullptr, *d_feat = nullptr, *d_spk = nullptr;
        // PCM staging: a ring of four device buffers filled by the copy engine on its own stream, so that the copy of piece
        // k+1 (and k+2) runs while the kernels of piece k are busy whatever lane they are on; events hand a buffer from the
        // copy to its kernel and back
        const int slot = (int)(ctx->stage_next++ & 3u);
        {
            // all four buffers at once, so that growing one never frees memory another piece is still using
            const size_t need = (size_t)piece * L * esz;
            if (ctx->d_stage_bytes[8] < need)
                for (int q = 0; q < 4; ++q) { void *tmp; if ((rc = lsm_stage_device(ctx, 8 + q, need, &tmp)) != LSM_OK) return rc; }
            d_in = ctx->d_stage[8 + slot];
        }
        if (!feat_direct && (rc = lsm_stage_device(ctx, 12 + lane, (size_t)piece * feat_per * sizeof(double), &d_feat)) != LSM_OK) return rc;
        if (h_spikes_or_null && (rc = lsm_stage_device(ctx, 14 + lane, (size_t)piece * spk_per, &d_spk)) != LSM_OK) return rc;
        cudaStream_t cs = ctx->copy_stream[1];
        if (ctx->stage_busy[slot]) LSM_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_stage_free[slot], 0));      // its last kernel has read it
        LSM_CUDA(ctx, cudaMemcpyAsync(d_in, (const char *)h_pcm + (size_t)off * L * esz, (size_t)n * L * esz, cudaMemcpyDefault, cs));
        LSM_CUDA(ctx, cudaEventRecord(ctx->ev_stage_full[slot], cs));
        LSM_CUDA(ctx, cudaStreamWaitEvent(ls, ctx->ev_stage_full[slot], 0));
        double *out = feat_direct ? (double *)dv_feat + (size_t)off * feat_per : (double *)d_feat;
        fe->next_pcm16 = i16 ? (const int16_t *)d_in : nullptr;
        if (warp_specialised)
            rc = lsm_launch_pipeline_lanes(ctx, fe, res, i16 ? nullptr : (const float *)d_in, n, (uint8_t *)d_spk, feature_mask, nan_to_num,
                                           out, ls, lane, off);
        else       // the lane = channel fused kernel on the staged copy
            rc = lsm_launch_fused(ctx, fe, res, i16 ? nullptr : (const float *)d_in, n, (uint8_t *)d_spk, feature_mask, nan_to_num, out, ls, off);
        fe->next_pcm16 = nullptr;
        if (rc != LSM_OK) return rc;
        LSM_CUDA(ctx, cudaEventRecord(ctx->ev_stage_free[slot], ls));
        ctx->stage_busy[slot] = 1;
        if (!feat_direct)
            LSM_CUDA(ctx, cudaMemcpyAsync(h_features + (size_t)off * feat_per, d_feat, (size_t)n * feat_per * sizeof(double), cudaMemcpyDeviceToHost, ls));
        if (h_spikes_or_null)
            LSM_CUDA(ctx, cudaMemcpyAsync(h_spikes_or_null + (size_t)off * spk_per, d_spk, (size_t)n * spk_per, cudaMemcpyDeviceToHost, ls));
    }
    if (only_lane < 0) {
        LSM_CUDA(ctx, cudaStreamSynchronize(lane_stream(ctx, 0)));
        LSM_CUDA(ctx, cudaStreamSynchronize(lane_stream(ctx, 1)));
        return check_pipe_error(ctx, fe);
    }
    return LSM_OK;
}

extern "C" int lsm_pipeline_run(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm,
                                int32_t B, uint32_t feature_mask, int32_t nan_to_num, uint8_t *d_spikes,
                                double *d_features)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || !res || B < 0 || (B > 0 && (!d_pcm || !d_features)))
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_pipeline_run: bad argument");
    if (fe->p.channels * fe->p.redundancy != res->p.num_inputs || fe->p.n_bins * fe->p.n_thresholds != res->p.num_steps)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "front end emits %dx%d spike trains, reservoir expects %dx%d",
                 fe->p.channels * fe->p.redundancy, fe->p.n_bins * fe->p.n_thresholds, res->p.num_inputs, res->p.num_steps);
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    // default shape: the warp-specialised kernel, two half batches on the ctx's two launch lanes (they share the SMs)
    if (lsm_pipeline_lanes_eligible(fe, res, d_pcm, false))
        return lsm_pipeline_lanes_device(ctx, fe, res, d_pcm, B, d_spikes, feature_mask, nan_to_num, d_features, ctx->stream);
    // one fused kernel when the pair allows it (spikes handed over in shared memory; d_spikes optional)
    if (lsm_fused_npt(fe, res))
        return lsm_launch_fused(ctx, fe, res, d_pcm, B, d_spikes, feature_mask, nan_to_num, d_features, ctx->stream, 0);
    if (lsm_mel_fused_ok(fe, res))
        return lsm_launch_mel_fused(ctx, fe, res, d_pcm, B, d_spikes, feature_mask, nan_to_num, d_features, ctx->stream, 0);
    if (!d_spikes) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_pipeline_run: this configuration runs as two kernels and needs a d_spikes buffer");
    int rc = frontend_launch(ctx, fe, d_pcm, B, d_spikes, nullptr, ctx->stream);
    if (rc != LSM_OK) return rc;
    return lsm_launch_reservoir(ctx, res, d_spikes, B, feature_mask, nan_to_num, d_features, nullptr, ctx->stream);
}

extern "C" int lsm_pipeline_is_fused(const lsm_frontend *fe, const lsm_reservoir *res)
{
    return (fe && res && (lsm_fused_npt(fe, res) || lsm_mel_fused_ok(fe, res))) ? 1 : 0;
}

extern "C" int lsm_pipeline_run_host(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *h_pcm,
                                     int32_t B, uint32_t feature_mask, int32_t nan_to_num, double *h_features,
                                     uint8_t *h_spikes_or_null)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || !res || B < 0 || (B > 0 && (!h_pcm || !h_features)))
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_pipeline_run_host: bad argument");
    if (fe->p.channels * fe->p.redundancy != res->p.num_inputs || fe->p.n_bins * fe->p.n_thresholds != res->p.num_steps)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "front end emits %dx%d spike trains, reservoir expects %dx%d",
                 fe->p.channels * fe->p.redundancy, fe->p.n_bins * fe->p.n_thresholds, res->p.num_inputs, res->p.num_steps);
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int L = fe->p.n_samples;
    const size_t spk_per = (size_t)res->p.num_inputs * res->p.num_steps;
    const int nkeys = __builtin_popcount(feature_mask & 0xFFu);
    const size_t feat_per = (size_t)nkeys * res->p.n_out;
    // Default shape: the warp-specialised kernel, fed by the copy engine, two halves on the two launch lanes
    if (lsm_pipeline_lanes_eligible(fe, res, (const void *)256, false))
        return pipeline_lanes_host(ctx, fe, res, h_pcm, false, B, feature_mask, nan_to_num, h_features, h_spikes_or_null, -1);
    // Zero-copy path: with pinned host buffers the one fused persistent kernel reads each PCM sample exactly once
    // straight over PCIe as it consumes it (~15 GB/s at full speed) and writes the feature rows straight back, so
    // there is no staging copy, no chunking and no per-chunk drain tail.  Pageable buffers take the staged path below.
    {
        void *dv_pcm = nullptr, *dv_feat = nullptr;
        if (lsm_fused_npt(fe, res) && !getenv("LSM_NO_ZEROCOPY") && device_visible(h_pcm, &dv_pcm) &&
            device_visible(h_features, &dv_feat)) {
            void *d_spk = nullptr;
            int rc0;
            if (h_spikes_or_null && (rc0 = lsm_stage_device(ctx, 1, (size_t)B * spk_per, &d_spk)) != LSM_OK) return rc0;
            const int wave = lsm_fused_wave(ctx, fe, res);
            if (!h_spikes_or_null && wave > 0 && B >= 2 * wave && !getenv("LSM_NO_SPLIT")) {
                // two launches on the two lanes (two scratch slots): the second half fills the drain tail of the first
                const int n0 = B / 2;
                if ((rc0 = lsm_launch_fused(ctx, fe, res, (const float *)dv_pcm, n0, nullptr, feature_mask, nan_to_num,
                                            (double *)dv_feat, ctx->own_stream, 0)) != LSM_OK) return rc0;
                if ((rc0 = lsm_launch_fused(ctx, fe, res, (const float *)dv_pcm + (size_t)n0 * L, B - n0, nullptr, feature_mask, nan_to_num,
                                            (double *)dv_feat + (size_t)n0 * feat_per, ctx->copy_stream[0], n0)) != LSM_OK) return rc0;
                LSM_CUDA(ctx, cudaStreamSynchronize(ctx->own_stream));
                LSM_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream[0]));
                return LSM_OK;
            }
            if ((rc0 = lsm_launch_fused(ctx, fe, res, (const float *)dv_pcm, B, (uint8_t *)d_spk, feature_mask, nan_to_num,
                                        (double *)dv_feat, ctx->stream, 0)) != LSM_OK) return rc0;
            if (h_spikes_or_null)
                LSM_CUDA(ctx, cudaMemcpyAsync(h_spikes_or_null, d_spk, (size_t)B * spk_per, cudaMemcpyDeviceToHost, ctx->stream));
            LSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            return LSM_OK;
        }
    }
    // Pageable buffers, fused pair: a ring of four pinned pieces.  Host threads copy piece k into the ring while the kernels of
    // pieces k-1 and k-2 run on the two launch lanes reading their pieces over PCIe (zero-copy) and writing their feature rows
    // into the pinned half of the same slot; finished rows are copied out to the caller's array two pieces later.
    if (lsm_fused_npt(fe, res) && !h_spikes_or_null && !getenv("LSM_NO_PAGEABLE_RING")) {
        lsm_copy_pool *pool = copy_pool(ctx);
        if (!pool) LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
        int piece = lsm_fused_wave(ctx, fe, res);
        if (piece <= 0) piece = 512;
        while (piece > 1024) piece /= 2;                         // ~600 utterances: 38 MB of PCM per slot
        if (piece > B) piece = B;
        const size_t in_bytes = (size_t)piece * L * sizeof(float), out_bytes = (size_t)piece * feat_per * sizeof(double);
        const size_t slot_bytes = ((in_bytes + 255) & ~(size_t)255) + out_bytes;
        void *ringp;
        int rc;
        if ((rc = lsm_stage_pinned(ctx, 0, 4 * slot_bytes, &ringp)) != LSM_OK) return rc;
        const int n_pieces = (B + piece - 1) / piece;
        auto slot_in = [&](int k) { return (char *)ringp + (size_t)(k & 3) * slot_bytes; };
        auto slot_out = [&](int k) { return slot_in(k) + ((in_bytes + 255) & ~(size_t)255); };
        auto count = [&](int k) { return k + 1 < n_pieces ? piece : B - k * piece; };
        for (int k = 0; k < n_pieces + 2; ++k) {
            if (k >= 2) {
                // piece k-2 is complete: its rows leave the ring (and its slot is free for piece k+2)
                const int j = k - 2;
                LSM_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage_free[j & 3]));
                pool->copy(h_features + (size_t)j * piece * feat_per, slot_out(j), (size_t)count(j) * feat_per * sizeof(double));
            }
            if (k < n_pieces) {
                const int n = count(k);
                pool->copy(slot_in(k), h_pcm + (size_t)k * piece * L, (size_t)n * L * sizeof(float));
                void *dv_in = nullptr, *dv_out = nullptr;
                if (!device_visible(slot_in(k), &dv_in) || !device_visible(slot_out(k), &dv_out))
                    LSM_FAIL(ctx, LSM_ERR_CUDA, "pinned staging ring is not device-visible");
                cudaStream_t ls = lane_stream(ctx, k & 1);
                if ((rc = lsm_launch_fused(ctx, fe, res, (const float *)dv_in, n, nullptr, feature_mask, nan_to_num, (double *)dv_out, ls,
                                           (long long)k * piece)) != LSM_OK) return rc;
                LSM_CUDA(ctx, cudaEventRecord(ctx->ev_stage_free[k & 3], ls));
            }
        }
        return LSM_OK;
    }
    // chunked, three legs on three streams: H2D(c+1) | kernels(c) | D2H(c-1).  A chunk is one full wave of the
    // persistent front-end grid (every CTA gets exactly one utterance, so a chunk has no drain tail); a short
    // remainder is folded into the last chunk.  The first chunk's H2D and the last chunk's D2H are the only
    // copies that are not hidden.
    const bool fused = lsm_fused_npt(fe, res) != 0;
    const bool mel_fused = !fused && lsm_mel_fused_ok(fe, res);
    // pageable arrays, fused mel pair: the same pinned ring + host copy threads as the stage calls (cudaMemcpyAsync from pageable
    // memory is a staged copy inside the driver at 6-8 GB/s, far below what the kernels take)
    if (mel_fused && !h_spikes_or_null && B >= 512 && is_pageable(h_pcm) && is_pageable(h_features) && !getenv("LSM_NO_PAGEABLE_RING"))
        return host_ring_staged(ctx, B, 768, (size_t)L * sizeof(float), feat_per * sizeof(double), h_pcm, h_features, 2,
                                [&](const void *d_in, void *d_out, int n, cudaStream_t st, long long off) {
                                    return lsm_launch_mel_fused(ctx, fe, res, (const float *)d_in, n, nullptr, feature_mask, nan_to_num,
                                                                (double *)d_out, st, off);
                                });
    int wave = fe->grid > 0 ? fe->grid : 1024;
    if (fused) { const int w = lsm_fused_wave(ctx, fe, res); if (w > 0) wave = w; }
    int n_chunks = B / wave;
    if (n_chunks < 1) n_chunks = 1;
    const int chunk = wave < B ? wave : B;                       // all chunks but the last
    const int last = B - (n_chunks - 1) * chunk;                 // chunk + remainder (< 2*chunk)
    const int cap = last > chunk ? last : chunk;
    void *d_pcm[2], *d_spk[2], *d_feat[2];
    int rc;
    void *base;
    if ((rc = lsm_stage_device(ctx, 0, 2 * (size_t)cap * L * sizeof(float), &base)) != LSM_OK) return rc;
    d_pcm[0] = base; d_pcm[1] = (char *)base + (size_t)cap * L * sizeof(float);
    if ((rc = lsm_stage_device(ctx, 1, 2 * (size_t)cap * spk_per, &base)) != LSM_OK) return rc;
    d_spk[0] = base; d_spk[1] = (char *)base + (size_t)cap * spk_per;
    if ((rc = lsm_stage_device(ctx, 2, 2 * (size_t)cap * feat_per * sizeof(double), &base)) != LSM_OK) return rc;
    d_feat[0] = base; d_feat[1] = (char *)base + (size_t)cap * feat_per * sizeof(double);
    cudaStream_t s_in = ctx->copy_stream[0], s_out = ctx->copy_stream[1], s_k = ctx->stream;
    // ev[0..1]: H2D of buffer b done; ev[2..3]: kernels on buffer b done; ev[4..5]: D2H of buffer b done
    for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1;
        const int n = (c + 1 < n_chunks) ? chunk : last;
        const size_t off = (size_t)c * chunk;
        if (c >= 2) LSM_CUDA(ctx, cudaStreamWaitEvent(s_in, ctx->ev[2 + b], 0));   // kernels of chunk c-2 released d_pcm[b]
        LSM_CUDA(ctx, cudaMemcpyAsync(d_pcm[b], h_pcm + off * L, (size_t)n * L * sizeof(float), cudaMemcpyHostToDevice, s_in));
        LSM_CUDA(ctx, cudaEventRecord(ctx->ev[b], s_in));
        LSM_CUDA(ctx, cudaStreamWaitEvent(s_k, ctx->ev[b], 0));
        if (c >= 2) LSM_CUDA(ctx, cudaStreamWaitEvent(s_k, ctx->ev[4 + b], 0));    // D2H of chunk c-2 released d_spk/d_feat[b]
        if (fused) {
            if ((rc = lsm_launch_fused(ctx, fe, res, (const float *)d_pcm[b], n, h_spikes_or_null ? (uint8_t *)d_spk[b] : nullptr,
                                       feature_mask, nan_to_num, (double *)d_feat[b], s_k, (long long)off)) != LSM_OK) return rc;
        } else if (mel_fused) {
            if ((rc = lsm_launch_mel_fused(ctx, fe, res, (const float *)d_pcm[b], n, h_spikes_or_null ? (uint8_t *)d_spk[b] : nullptr,
                                           feature_mask, nan_to_num, (double *)d_feat[b], s_k, (long long)off)) != LSM_OK) return rc;
        } else {
            if ((rc = frontend_launch(ctx, fe, (const float *)d_pcm[b], n, (uint8_t *)d_spk[b], nullptr, s_k)) != LSM_OK) return rc;
            if ((rc = lsm_launch_reservoir(ctx, res, (const uint8_t *)d_spk[b], n, feature_mask, nan_to_num,
                                           (double *)d_feat[b], nullptr, s_k, nullptr, (long long)off)) != LSM_OK) return rc;
        }
        LSM_CUDA(ctx, cudaEventRecord(ctx->ev[2 + b], s_k));
        LSM_CUDA(ctx, cudaStreamWaitEvent(s_out, ctx->ev[2 + b], 0));
        LSM_CUDA(ctx, cudaMemcpyAsync(h_features + off * feat_per, d_feat[b], (size_t)n * feat_per * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        if (h_spikes_or_null)
            LSM_CUDA(ctx, cudaMemcpyAsync(h_spikes_or_null + off * spk_per, d_spk[b], (size_t)n * spk_per, cudaMemcpyDeviceToHost, s_out));
        LSM_CUDA(ctx, cudaEventRecord(ctx->ev[4 + b], s_out));
    }
    LSM_CUDA(ctx, cudaStreamSynchronize(s_out));
    LSM_CUDA(ctx, cudaStreamSynchronize(s_k));
    LSM_CUDA(ctx, cudaStreamSynchronize(s_in));
    return LSM_OK;
}

// PCM16 input (what a Speech Commands WAV file holds): the int16 -> float32 conversion of the ingest
// (create_dataset.py:22-36, librosa.load) happens in the kernel, exactly; half the bytes cross PCIe.  Fused pairs only.
static int run_i16(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const int16_t *pcm16, int32_t B, uint32_t feature_mask,
                   int32_t nan_to_num, uint8_t *d_spikes, double *d_features, cudaStream_t st, const char *who)
{
    if (fe->p.channels * fe->p.redundancy != res->p.num_inputs || fe->p.n_bins * fe->p.n_thresholds != res->p.num_steps)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "front end emits %dx%d spike trains, reservoir expects %dx%d",
                 fe->p.channels * fe->p.redundancy, fe->p.n_bins * fe->p.n_thresholds, res->p.num_inputs, res->p.num_steps);
    if (!lsm_fused_npt(fe, res)) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "%s: PCM16 input needs a front end / reservoir pair that runs fused", who);
    fe->next_pcm16 = pcm16;
    {
        cudaPointerAttributes at;
        const bool on_device = cudaPointerGetAttributes(&at, pcm16) == cudaSuccess && at.type == cudaMemoryTypeDevice;
        cudaGetLastError();
        if (on_device && lsm_pipeline_lanes_eligible(fe, res, pcm16, true)) {
            const int rc = lsm_pipeline_lanes_device(ctx, fe, res, nullptr, B, d_spikes, feature_mask, nan_to_num, d_features, st);
            fe->next_pcm16 = nullptr;
            return rc;
        }
    }
    const int rc = lsm_launch_fused(ctx, fe, res, nullptr, B, d_spikes, feature_mask, nan_to_num, d_features, st, 0);
    fe->next_pcm16 = nullptr;
    return rc;
}

extern "C" int lsm_pipeline_run_i16(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const int16_t *d_pcm16, int32_t B,
                                    uint32_t feature_mask, int32_t nan_to_num, uint8_t *d_spikes, double *d_features)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || !res || B < 0 || (B > 0 && (!d_pcm16 || !d_features))) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_pipeline_run_i16: bad argument");
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    return run_i16(ctx, fe, res, d_pcm16, B, feature_mask, nan_to_num, d_spikes, d_features, ctx->stream, "lsm_pipeline_run_i16");
}

extern "C" int lsm_pipeline_run_host_async_i16(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const int16_t *h_pcm16,
                                               int32_t B, uint32_t feature_mask, int32_t nan_to_num, double *h_features,
                                               int32_t lane)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || !res || B < 0 || (B > 0 && (!h_pcm16 || !h_features)) || lane < 0 || lane > 1)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_pipeline_run_host_async_i16: bad argument");
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    void *dv_pcm = nullptr, *dv_feat = nullptr;
    if (!device_visible(h_pcm16, &dv_pcm) || !device_visible(h_features, &dv_feat))
        LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "lsm_pipeline_run_host_async_i16 needs pinned (or device) buffers");
    if (fe->p.channels * fe->p.redundancy == res->p.num_inputs && fe->p.n_bins * fe->p.n_thresholds == res->p.num_steps &&
        lsm_pipeline_lanes_eligible(fe, res, (const void *)256, true))
        return pipeline_lanes_host(ctx, fe, res, h_pcm16, true, B, feature_mask, nan_to_num, h_features, nullptr, lane);
    if (ctx->host_feed == LSM_FEED_COPY_ENGINE && lsm_fused_npt(fe, res) &&
        fe->p.channels * fe->p.redundancy == res->p.num_inputs && fe->p.n_bins * fe->p.n_thresholds == res->p.num_steps)
        return pipeline_lanes_host(ctx, fe, res, h_pcm16, true, B, feature_mask, nan_to_num, h_features, nullptr, lane, false);
    return run_i16(ctx, fe, res, (const int16_t *)dv_pcm, B, feature_mask, nan_to_num, nullptr, (double *)dv_feat,
                   lane == 0 ? ctx->own_stream : ctx->copy_stream[0], "lsm_pipeline_run_host_async_i16");
}

// Asynchronous variant for pinned host buffers only (the zero-copy path): enqueue on one of the ctx's two launch lanes and
// return.  Two calls on alternating lanes are in flight together (two scratch slots), so the drain tail of one batch overlaps
// the start of the next.  lsm_sync_all waits for both lanes.
extern "C" int lsm_pipeline_run_host_async(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *h_pcm,
                                           int32_t B, uint32_t feature_mask, int32_t nan_to_num, double *h_features,
                                           int32_t lane)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!fe || !res || B < 0 || (B > 0 && (!h_pcm || !h_features)) || lane < 0 || lane > 1)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_pipeline_run_host_async: bad argument");
    if (fe->p.channels * fe->p.redundancy != res->p.num_inputs || fe->p.n_bins * fe->p.n_thresholds != res->p.num_steps)
        LSM_FAIL(ctx, LSM_ERR_INVALID, "front end emits %dx%d spike trains, reservoir expects %dx%d",
                 fe->p.channels * fe->p.redundancy, fe->p.n_bins * fe->p.n_thresholds, res->p.num_inputs, res->p.num_steps);
    if (B == 0) return LSM_OK;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    void *dv_pcm = nullptr, *dv_feat = nullptr;
    if (!lsm_fused_npt(fe, res) || !device_visible(h_pcm, &dv_pcm) || !device_visible(h_features, &dv_feat))
        LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "lsm_pipeline_run_host_async needs pinned (or device) buffers and a pair that runs fused");
    if (lsm_pipeline_lanes_eligible(fe, res, (const void *)256, false))
        return pipeline_lanes_host(ctx, fe, res, h_pcm, false, B, feature_mask, nan_to_num, h_features, nullptr, lane);
    if (ctx->host_feed == LSM_FEED_COPY_ENGINE)
        return pipeline_lanes_host(ctx, fe, res, h_pcm, false, B, feature_mask, nan_to_num, h_features, nullptr, lane, false);
    return lsm_launch_fused(ctx, fe, res, (const float *)dv_pcm, B, nullptr, feature_mask, nan_to_num, (double *)dv_feat,
                            lane == 0 ? ctx->own_stream : ctx->copy_stream[0], 0);
}

extern "C" void *lsm_lane_stream(lsm_ctx *ctx, int32_t lane)
{
    if (!ctx || lane < 0 || lane > 1) return nullptr;
    return (void *)(lane == 0 ? ctx->own_stream : ctx->copy_stream[0]);
}

extern "C" int lsm_sync_all(lsm_ctx *ctx)
{
    if (!ctx) return LSM_ERR_INVALID;
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    LSM_CUDA(ctx, cudaStreamSynchronize(ctx->own_stream));
    for (int i = 0; i < 2; ++i) LSM_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream[i]));
    if (ctx->stream != ctx->own_stream) LSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LSM_OK;
}

// ------------------------------------------------------------------------------------ density
namespace {
__global__ void density_kernel(const uint8_t *__restrict__ x, long long n, unsigned long long *out)
{
    unsigned long long s = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n16 = n / 16;
    const uint4 *x4 = reinterpret_cast<const uint4 *>(x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = __ldg(x4 + i);
        // bytes are small counts (0/1 in practice): byte-wise sums via SAD against zero
        s += __vsadu4(v.x, 0) + __vsadu4(v.y, 0) + __vsadu4(v.z, 0) + __vsadu4(v.w, 0);
    }
    for (long long i = n16 * 16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}
}  // namespace

extern "C" int lsm_spike_density(lsm_ctx *ctx, const uint8_t *d_spikes, int64_t n_bytes, int64_t *h_out)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!h_out || n_bytes < 0 || (n_bytes > 0 && !d_spikes)) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_spike_density: bad argument");
    h_out[0] = 0; h_out[1] = n_bytes;
    if (n_bytes == 0) return LSM_OK;
    if (((uintptr_t)d_spikes & 15) != 0) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_spike_density: pointer must be 16-byte aligned");
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    void *d_sum;
    int rc;
    if ((rc = lsm_stage_device(ctx, 4, sizeof(unsigned long long), &d_sum)) != LSM_OK) return rc;
    LSM_CUDA(ctx, cudaMemsetAsync(d_sum, 0, sizeof(unsigned long long), ctx->stream));
    density_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d_spikes, (long long)n_bytes, (unsigned long long *)d_sum);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    unsigned long long s = 0;
    LSM_CUDA(ctx, cudaMemcpyAsync(&s, d_sum, sizeof(s), cudaMemcpyDeviceToHost, ctx->stream));
    LSM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h_out[0] = (int64_t)s;
    return LSM_OK;
}

// ------------------------------------------------------------------------------------ fp64 ceiling
namespace {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = __dmul_rn(v[i], b); v[i] = __dadd_rn(v[i], a); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int lsm_fp64_peak_gops(lsm_ctx *ctx, double *h_out)
{
    if (!ctx) return LSM_ERR_INVALID;
    if (!h_out) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_fp64_peak_gops: null argument");
    LSM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int threads = 256, blocks = ctx->sm_count * 4, iters = 20000;   // 8 warps per scheduler
    void *d_out;
    int rc;
    if ((rc = lsm_stage_device(ctx, 5, sizeof(double) * threads * blocks, &d_out)) != LSM_OK) return rc;
    cudaEvent_t e0, e1;
    LSM_CUDA(ctx, cudaEventCreate(&e0));
    LSM_CUDA(ctx, cudaEventCreate(&e1));
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double *)d_out, 200, 1.0, 0.999999);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        LSM_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double *)d_out, iters, 1.0, 0.999999);
        LSM_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        LSM_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        LSM_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double gops = (double)blocks * threads * iters * 16.0 / (ms * 1e6);
        if (gops > best) best = gops;
    }
    ctx->launches += 4;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *h_out = best;
    return LSM_OK;
}
