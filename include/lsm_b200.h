/*
 * lsm_b200.h — C ABI of liblsmb200.so: the B200 (sm_100a) implementation of the
 * audio -> spike train -> liquid-state-machine -> per-neuron feature path of
 * adelitoo/lsm-speech-classifier.
 *
 * The reference has no FFI of its own: it is Python that delegates to three packages.  Each
 * entry point below therefore names the reference CALL SITE (file:line under /root/reference)
 * whose work it replaces.  INTEGRATION.md shows the ctypes stubs a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 (LSM_OK) or a negative lsm_status; lsm_last_error(ctx) has the text
 *   - the caller owns every buffer it passes; the library owns only what lives inside its handles
 *   - "d_" pointers are device memory on the ctx's device, "h_" pointers are host memory
 *   - one lsm_ctx per (process, device); a ctx is not thread-safe; work is enqueued on the ctx's
 *     stream (lsm_set_stream) and is asynchronous unless the name ends in _host or _sync
 *   - there is NO CPU fallback: without a CUDA device lsm_ctx_create fails
 */
#ifndef LSM_B200_H
#define LSM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lsm_ctx lsm_ctx;
typedef struct lsm_frontend lsm_frontend;
typedef struct lsm_reservoir lsm_reservoir;

typedef enum {
    LSM_OK = 0,
    LSM_ERR_INVALID = -1,     /* bad argument / unsupported shape */
    LSM_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed */
    LSM_ERR_NOMEM = -3,
    LSM_ERR_UNSUPPORTED = -4  /* valid request this build cannot serve */
} lsm_status;

enum { LSM_FILTERBANK_GAMMATONE = 0, LSM_FILTERBANK_MEL = 1 };

/* Feature keys, bit k of feature_mask <-> k-th name of FEATURE_SETS['all']
 * (extract_lsm_features.py:20-22).  Output layout is key-major: all output neurons of the
 * lowest selected key, then the next (extract_lsm_features.py:85-87).                      */
enum {
    LSM_F_SPIKE_COUNTS = 1u << 0, LSM_F_SPIKE_VARIANCES = 1u << 1, LSM_F_MEAN_SPIKE_TIMES = 1u << 2,
    LSM_F_FIRST_SPIKE_TIMES = 1u << 3, LSM_F_LAST_SPIKE_TIMES = 1u << 4, LSM_F_MEAN_ISI = 1u << 5,
    LSM_F_ISI_VARIANCES = 1u << 6, LSM_F_BURST_COUNTS = 1u << 7
};

/* ---------------------------------------------------------------- context */
int lsm_ctx_create(lsm_ctx **out, int device_ordinal);
void lsm_ctx_destroy(lsm_ctx *ctx);
const char *lsm_last_error(const lsm_ctx *ctx);
/* Enqueue on an existing cudaStream_t (e.g. torch's current stream; NULL = CUDA's default stream). */
int lsm_set_stream(lsm_ctx *ctx, void *cuda_stream);
/* Go back to the non-blocking stream the ctx created for itself (the initial state). */
int lsm_reset_stream(lsm_ctx *ctx);
int lsm_sync(lsm_ctx *ctx);
/* How the asynchronous host-buffer calls (lsm_pipeline_run_host_async*) bring pinned PCM to the kernel:
 * LSM_FEED_ZERO_COPY (default) - the fused kernel reads the pinned host buffer itself, sample by sample as it filters (every
 *   sample crosses PCIe exactly once, no staging buffer);
 * LSM_FEED_COPY_ENGINE - one cudaMemcpyAsync per call into a device staging buffer on the launch lane, then the kernel on
 *   device memory; the copy of one lane overlaps the kernel of the other.  Better when many GPUs pull from one host
 *   (zero-copy loads of eight GPUs share the host's read bandwidth at small request size).                             */
#define LSM_FEED_ZERO_COPY 0
#define LSM_FEED_COPY_ENGINE 1
int lsm_ctx_set_host_feed(lsm_ctx *ctx, int32_t mode);
/* Kernels launched by this ctx since creation (bench.py's gpu_launches). */
int64_t lsm_launch_count(const lsm_ctx *ctx);
int lsm_sm_count(const lsm_ctx *ctx);

/* ---------------------------------------------------------------- stage 1: PCM -> spike trains
 * Replaces, per utterance, create_dataset.py:148-158:
 *     audio_to_spectrogram (:39-78; gammatone.gtgram.gtgram :51-58 or
 *                           librosa melspectrogram + power_to_db :45-48,
 *                           dB floor :59-60, min-max :62-67, scipy zoom :69-78)
 *     convert_spectrogram_to_spikes_hysteresis (:81-98)
 *     create_pure_redundancy (:101-104)
 */
typedef struct {
    int32_t kind;            /* LSM_FILTERBANK_* (create_dataset.py:43,49) */
    int32_t channels;        /* n_filters */
    int32_t n_samples;       /* 16000 = SAMPLE_RATE*DURATION (create_dataset.py:10-11,28) */
    int32_t nwin, hop;       /* gammatone: gtgram window/hop in samples (400, 160) */
    int32_t n_fft, mel_hop;  /* mel: 2048, 160 (librosa defaults + create_dataset.py:44) */
    int32_t n_bins;          /* TIME_BINS = 100 (create_dataset.py:12) */
    int32_t n_thresholds;    /* <= 8 */
    int32_t redundancy;      /* REDUNDANCY_FACTOR (create_dataset.py:17) */
    double thresholds_desc[8]; /* sorted(thresholds, reverse=True)  (create_dataset.py:87) */
    double lower_bounds[8];    /* threshold - hysteresis_gap in fp64 (create_dataset.py:89) */
} lsm_frontend_params;

/* h_table: gammatone -> double[channels][10] rows {A0,A11,A12,A13,A14,A2,B0,B1,B2,gain}, row 0 = the
 *          lowest centre frequency (gammatone.gtgram.gtgram_xe's flipped make_erb_filters table);
 *          mel -> float[channels][1 + n_fft/2] (librosa.filters.mel).
 * h_zoom_i0/h_zoom_f: int32/double[n_bins], scipy.ndimage.zoom(order=1) source index and fraction
 *          per output bin (ignored when the spectrogram already has n_bins columns).        */
int lsm_frontend_create(lsm_ctx *ctx, const lsm_frontend_params *p, const void *h_table,
                        const int32_t *h_zoom_i0, const double *h_zoom_f, lsm_frontend **out);
void lsm_frontend_destroy(lsm_frontend *fe);
/* Mel only, optional: replace the library's own (host libm) STFT tables with the caller's, so that another
 * implementation filtering with the same tables is reproduced bit for bit.  h_window: double[n_fft] periodic hann;
 * h_tw: double[n_fft/4][2] = exp(-2*pi*i*q/(n_fft/2)); h_tw2: double[n_fft/2+1][2] = exp(-2*pi*i*k/n_fft).   */
int lsm_frontend_mel_tables(lsm_ctx *ctx, lsm_frontend *fe, const double *h_window, const double *h_tw,
                            const double *h_tw2);
/* Gammatone only: how the filter bank is evaluated.  Either way the spike trains that leave the library are those of
 * the reference's operation order (scipy.signal.lfilter x 4, /gain, square, window mean: create_dataset.py:51-58 via
 * gammatone.gtgram), byte for byte.
 *   LSM_FILTER_EXACT        every fp64 operation in the reference's order (35 separate roundings per channel-sample).
 *   LSM_FILTER_SPECULATIVE  (default) a mathematically equivalent arrangement of the same cascade in 13 fused
 *                           multiply-adds.  The library derives, per channel, a worst-case bound on the distance between the
 *                           two arrangements (lsm_gammatone_error_bound) and carries it through dB,
 *                           floor, min-max and zoom: an utterance in which that bound could change the outcome of any
 *                           encoder comparison (threshold, hysteresis bound, silent-clip test) is filtered again in exact
 *                           mode.  delta_db >= 0 widens the test by that many decibels on top of the bound (default 0).
 *                           The optional d_spec_norm dump always uses the exact path.
 * lsm_frontend_set_bound_scale: diagnostic - multiplies the derived bound (1 = the guarantee; 0 switches the bound off, so
 *                           that tests can look at the speculative plane alone).
 * lsm_frontend_reruns: number of utterances that were filtered twice since creation / the last reset (waits for the
 * front end's last launch).                                                                                          */
#define LSM_FILTER_EXACT 0
#define LSM_FILTER_SPECULATIVE 1
int lsm_frontend_set_mode(lsm_ctx *ctx, lsm_frontend *fe, int mode, double delta_db);
int lsm_frontend_set_bound_scale(lsm_ctx *ctx, lsm_frontend *fe, double scale);
int lsm_frontend_reruns(lsm_ctx *ctx, lsm_frontend *fe, int64_t *h_out, int reset);
/* d_pcm: float[B][n_samples].  d_spikes: uint8[B][channels*redundancy][n_bins*n_thresholds] — the
 * X_spikes layout of speech_spike_dataset_pure_redundancy.npz (create_dataset.py:168).
 * d_spec_norm_or_null: optional double[B][channels][n_bins] dump of the normalised, resampled
 * spectrogram (what audio_to_spectrogram returns) for parity tests.                         */
int lsm_frontend_encode(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int32_t B,
                        uint8_t *d_spikes, double *d_spec_norm_or_null);
/* Same through host buffers: H2D, kernel, D2H, synchronous. */
int lsm_frontend_encode_host(lsm_ctx *ctx, lsm_frontend *fe, const float *h_pcm, int32_t B,
                             uint8_t *h_spikes);

/* Diagnostic behind the speculative mode: filters every utterance with BOTH arrangements and reports, per utterance,
 * d_out[8*b + ...]: [0] largest |dB_speculative - dB_exact| over the 98 x channels window cells, [1] the largest ratio of
 * that distance to the cell's derived bound (the guarantee is [1] <= 1), [2] and [3] the same ratio for the plane's maximum
 * and (floored) minimum, [4] and [5] those two bounds in dB, [6] max |sample|, [7] the plane's range in dB.             */
int lsm_frontend_audit(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int32_t B, double *d_out);
/* The bound itself (host code, no device work): h_kappa[c] bounds |window amplitude (speculative) - window amplitude
 * (reference order)| of channel c for an utterance with max |sample| = 1; it scales linearly with the peak level.
 * h_table as for lsm_frontend_create.                                                                               */
int lsm_gammatone_error_bound(const double *h_table, int32_t channels, int32_t n_samples, double *h_kappa);

/* Host-side design helpers for callers without numpy (plain libm; no device work).
 * lsm_gammatone_design: the table gammatone.gtgram.gtgram_xe filters with - make_erb_filters(fs, centre_freqs(fs,
 *   channels, f_min)) flipped so row 0 is the lowest centre frequency; h_out: double[channels][10].
 * lsm_zoom_table: scipy.ndimage.zoom(order=1) source index / fraction for n_in -> n_out columns.           */
int lsm_gammatone_design(double fs, int32_t channels, double f_min, double *h_out);
int lsm_zoom_table(int32_t n_in, int32_t n_out, int32_t *h_i0, double *h_f);

/* The encoder alone, on spectrograms already in device memory: convert_spectrogram_to_spikes_hysteresis
 * (create_dataset.py:81-98) + create_pure_redundancy (:101-104).  d_spec: double[B][C][n_bins] or, when
 * is_f32 != 0, float[B][C][n_bins] compared against float-rounded thresholds as numpy does for a
 * float32 spectrogram.  thr_desc/lower: host arrays of K values (descending thresholds, threshold - gap). */
int lsm_hysteresis_encode(lsm_ctx *ctx, const void *d_spec, int32_t is_f32, int32_t B, int32_t C, int32_t n_bins,
                          const double *h_thr_desc, const double *h_lower, int32_t K, int32_t redundancy,
                          uint8_t *d_spikes);

/* ---------------------------------------------------------------- stage 2+3: spike trains -> features
 * Replaces SNN(simulation_params=...) (extract_lsm_features.py:188) and, per utterance,
 *     lsm.reset(); lsm.set_input_spike_times(sample); lsm.simulate()   (:79-81)
 *     lsm.extract_features_from_spikes() + key selection + nan_to_num (:83-87)
 * The reservoir is generated on the HOST (numpy RandomState, see reservoir.py) and uploaded, so the
 * CPU oracle and the GPU share identical arrays.  Weights are int32 multiples of 2^-w_shift.  */
typedef struct {
    int32_t num_neurons;     /* N */
    int32_t num_inputs;      /* rows of one sample = channels*redundancy */
    int32_t num_steps;       /* T = columns of one sample (400) */
    int32_t refractory;      /* refractory_period */
    int32_t w_shift;         /* 24 */
    int32_t n_out;           /* num_output_neurons */
    double theta;            /* membrane_threshold */
} lsm_reservoir_params;

int lsm_reservoir_create(lsm_ctx *ctx, const lsm_reservoir_params *p,
                         const int32_t *h_w_rowptr, const int32_t *h_w_col, const int32_t *h_w_q, /* CSR, row = postsynaptic */
                         const int32_t *h_in_rowptr, const int32_t *h_in_col, const double *h_in_val, /* CSR, row = neuron */
                         const double *h_leak,            /* double[N] */
                         const int32_t *h_out_idx,        /* int32[n_out], ascending */
                         lsm_reservoir **out);
/* Strict reservoir: the weights as fp64 numbers (snnpy-style normal draws) instead of int32 multiples of
 * 2^-w_shift, and every neuron's recurrent current formed as SURVEY.md 8c S6 words it: the fp64 sum of the weights of its spiking
 * presynaptic neurons, added one by one in ascending presynaptic index (h_w_col strictly ascending inside a row; w_shift is
 * ignored).  Same calls afterwards; such a reservoir runs on the event-driven arm as its own kernel (no fusion with a front end,
 * no dense arm), 2.8 x the time of a quantised one (profiles/r2_config4.md).  The quantised form (lsm_reservoir_create) stays the default of the
 * Python layer because its sums are exact in any order; this form exists so that "computed in its fp64 accumulation order"
 * (BASELINE.json north star) can be had literally.                                                                          */
int lsm_reservoir_create_f64(lsm_ctx *ctx, const lsm_reservoir_params *p,
                             const int32_t *h_w_rowptr, const int32_t *h_w_col, const double *h_w_val,
                             const int32_t *h_in_rowptr, const int32_t *h_in_col, const double *h_in_val,
                             const double *h_leak, const int32_t *h_out_idx, lsm_reservoir **out);
void lsm_reservoir_destroy(lsm_reservoir *res);
/* d_spikes: uint8[B][num_inputs][num_steps].  d_features: double[B][popcount(mask)*n_out], RAW
 * (un-standardised).  nan_to_num != 0 applies extract_lsm_features.py:85's np.nan_to_num on device.
 * d_raster_or_null: optional uint8[B][num_steps][N] = lsm.spike_matrix (Time x Neurons,
 * extract_lsm_features.py:113-116) for parity tests and diagnostics.                        */
int lsm_reservoir_run(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int32_t B,
                      uint32_t feature_mask, int32_t nan_to_num, double *d_features,
                      uint8_t *d_raster_or_null);
int lsm_reservoir_run_host(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *h_spikes, int32_t B,
                           uint32_t feature_mask, int32_t nan_to_num, double *h_features,
                           uint8_t *h_raster_or_null);
/* How lsm_reservoir_run / lsm_reservoir_run_host form the recurrent current of lsm.simulate() (extract_lsm_features.py:81;
 * BASELINE.json north star, kernel 2: "event-driven gather ... or ... a tensor-core tile, choosing whichever ncu shows wins").
 * Both arms produce the same rasters and features bit for bit (integer weights: the sum is exact in any order).
 *   LSM_RESERVOIR_EVENT (default)  one CTA per utterance, all steps in one launch, state in registers / shared memory; work per
 *                                  step = (neurons that fired at t-1) x N integer adds.
 *   LSM_RESERVOIR_DENSE            one launch per time step for the whole batch: spikes[B,N] . W[N,N] on the integer tensor cores
 *                                  (tcgen05.mma kind::i8 over 8-bit digit planes of the weights, int32 accumulators in TMEM),
 *                                  membrane update as the epilogue; work per step = N x N regardless of activity.  Needs at most
 *                                  one input row per neuron; the fused all-gather is not available (LSM_ERR_UNSUPPORTED).
 * profiles/r2_config4.md holds the measured comparison; the event-driven arm wins at every operating point of the reference. */
#define LSM_RESERVOIR_EVENT 0
#define LSM_RESERVOIR_DENSE 1
int lsm_reservoir_set_mode(lsm_ctx *ctx, lsm_reservoir *res, int32_t mode);
/* Diagnostic of the dense arm: one contraction.  d_s: uint8[B][N] spike bytes (any non-zero = fired) in the caller's neuron
 * order, B <= 8192; d_acc: int32[B][N], d_acc[b][i] = sum_j Wq[i][j] * s[b][j] exactly as the tensor cores and the digit-plane
 * recombination form it (tests compare it with the integer matrix product).                                              */
int lsm_reservoir_dense_probe(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_s, int32_t B, int32_t *d_acc);

/* Network diagnostics of run_network_diagnostics (extract_lsm_features.py:92-152) reduced on the device instead of
 * shipping the raster: for each utterance, d_diag[2*b] = neurons (of all N) that fired at least once,
 * d_diag[2*b+1] = total spikes.  Participation % = 100*d_diag[2b]/N, average spikes per neuron = d_diag[2b+1]/N. */
/* Fused feature all-gather (the one collective of the path): from now on every launch that computes features
 * with this reservoir also stores the row of utterance u at row (row0 + u) of each of the n <= 8 matrices d_gather[k]
 * (double[total_rows][n_keys*n_out], device pointers - typically every rank's gather matrix, mapped through CUDA IPC, the
 * caller's own included) directly from the readout epilogue, over NVLink: no collective kernel, nothing to schedule beside the
 * persistent kernels.  The rows are complete when the launch has completed; ranks synchronise (a barrier) before reading.
 * n = 0 switches it off.  The setting is read when a launch is enqueued.                                                    */
int lsm_reservoir_set_gather(lsm_ctx *ctx, lsm_reservoir *res, double *const *d_gather, int32_t n, int64_t row0);
/* Gather matrices for the above across processes: device memory of this ctx's device exported as a 64-byte CUDA IPC handle
 * (create), mapped into another rank's process with that rank's device current (open; the pointer then goes into that
 * rank's d_gather list), unmapped (close) and freed by its owner (destroy).  Peer access over NVLink is enabled on open. */
int lsm_peer_buffer_create(lsm_ctx *ctx, int64_t bytes, void **d_ptr, void *h_handle64);
int lsm_peer_buffer_open(lsm_ctx *ctx, const void *h_handle64, void **d_ptr);
int lsm_peer_buffer_close(lsm_ctx *ctx, void *d_ptr);
int lsm_peer_buffer_destroy(lsm_ctx *ctx, void *d_ptr);
int lsm_reservoir_diagnostics(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int32_t B, int32_t *d_diag);

/* ---------------------------------------------------------------- the whole path, host buffers
 * audio -> features in one call: the loop bodies of create_dataset.py:143-162 and
 * extract_lsm_features.py:78-87 for B utterances.  Copies are chunked and overlapped with the
 * kernels on internal streams.  h_spikes_or_null: optionally also return the spike trains (the
 * stage-1 file content).  Synchronous.  When h_pcm and h_features are pinned host memory (or device memory) the
 * fused kernel addresses them directly: PCM is read across PCIe as it is consumed and no staging copy is made.                                                       */
int lsm_pipeline_run_host(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *h_pcm,
                          int32_t B, uint32_t feature_mask, int32_t nan_to_num, double *h_features,
                          uint8_t *h_spikes_or_null);
/* Device-resident variant: d_pcm -> d_features on the ctx's stream.  Runs as ONE fused kernel when the pair
 * allows it (gammatone: redundancy 1, reservoir width = channels x {4,8,16}; mel: redundancy 1, channels a multiple of
 * 32, reservoir of at most 1024 neurons); d_spikes_or_null then only receives a copy of the spike trains if non-NULL.
 * Otherwise two kernels, and d_spikes is required.                                                         */
int lsm_pipeline_run(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm,
                     int32_t B, uint32_t feature_mask, int32_t nan_to_num, uint8_t *d_spikes_or_null,
                     double *d_features);

/* Asynchronous form of lsm_pipeline_run_host for PINNED host buffers (the zero-copy path; LSM_ERR_UNSUPPORTED otherwise):
 * the call enqueues the fused kernel on launch lane 0 or 1 of the ctx and returns; the kernel reads the PCM and writes the
 * feature rows across PCIe.  Calls on alternating lanes overlap (a front end has two scratch slots), which hides the drain
 * tail of one batch under the start of the next.  lsm_sync_all waits for everything the ctx has enqueued.                 */
int lsm_pipeline_run_host_async(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *h_pcm, int32_t B,
                                uint32_t feature_mask, int32_t nan_to_num, double *h_features, int32_t lane);
int lsm_sync_all(lsm_ctx *ctx);
/* PCM16 input - the samples as a 16 kHz mono WAV file holds them (the ingest step
 * create_dataset.py:22-36 performs with librosa.load).  The kernel converts int16 -> sample / 32768 exactly as the float32
 * path would have received it, so spike trains and features are the same bytes; half as many bytes cross PCIe.
 * lsm_pipeline_run_i16: device (or pinned host) pointers on the ctx stream, like lsm_pipeline_run; d_spikes optional.
 * lsm_pipeline_run_host_async_i16: like lsm_pipeline_run_host_async.  Both need a pair that runs fused.                  */
int lsm_pipeline_run_i16(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const int16_t *d_pcm16, int32_t B,
                         uint32_t feature_mask, int32_t nan_to_num, uint8_t *d_spikes, double *d_features);
int lsm_pipeline_run_host_async_i16(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const int16_t *h_pcm16, int32_t B,
                                    uint32_t feature_mask, int32_t nan_to_num, double *h_features, int32_t lane);
/* The cudaStream_t of launch lane 0 / 1, so that a caller can order its own work (an NCCL all-gather of the feature rows, a copy)
 * after an asynchronous call on that lane.  NULL for a bad argument.                                                       */
void *lsm_lane_stream(lsm_ctx *ctx, int32_t lane);
/* 1 if lsm_pipeline_run / lsm_pipeline_run_host execute this pair fused - the spike train is handed to the reservoir inside the
 * kernel and never written to device memory unless asked for (gammatone: one persistent kernel; mel: the power kernel followed by
 * one epilogue + reservoir kernel) - else 0 (front-end kernel, spike trains in device memory, reservoir kernel).               */
int lsm_pipeline_is_fused(const lsm_frontend *fe, const lsm_reservoir *res);

/* Sum of all spike bytes and their count: the two integers calculate_theoretical_w_critico
 * reduces over X_train[:500] (extract_lsm_features.py:40-44).  h_out: int64[2].            */
int lsm_spike_density(lsm_ctx *ctx, const uint8_t *d_spikes, int64_t n_bytes, int64_t *h_out);

/* ---------------------------------------------------------------- upstream of the path
 * Sample-rate conversion of the ingest step: the resampling inside librosa.load(filepath, sr=16000) at create_dataset.py:26, as
 * librosa's res_type="polyphase" = scipy.signal.resample_poly(y, up, down) (default Kaiser window), bit for bit.
 * d_in: float[B][n_in]; up / down: the reduced rate ratio; h_taps: host float[up][taps_per_phase], the low-pass designed as scipy
 * does (firwin(2 * 10 * max(up, down) + 1, 1 / max(up, down), window=("kaiser", 5.0)) in float32, times up, zero-padded in front by
 * n_pre_pad), transposed and flipped per phase (scipy.signal._upfirdn._pad_h); n_pre_remove: leading outputs to drop
 * ((half_len + n_pre_pad) / down); d_out: float[B][n_out], n_out = ceil(n_in * up / down).  ingest.py builds these arguments.   */
int lsm_resample_poly(lsm_ctx *ctx, const float *d_in, int32_t B, int32_t n_in, int32_t up, int32_t down, const float *h_taps,
                      int32_t taps_per_phase, int32_t n_pre_remove, int32_t n_out, float *d_out);

/* ---------------------------------------------------------------- downstream of the path
 * sklearn.preprocessing.StandardScaler on the device, bit-exact with scikit-learn's dense float64 path
 * (extract_lsm_features.py:199-201).  d_X: double[n][F] row-major; d_mean/d_var/d_scale: double[F].
 * transform may run in place (d_out == d_X).                                                                  */
int lsm_standardize_fit(lsm_ctx *ctx, const double *d_X, int32_t n, int32_t F, double *d_mean, double *d_var, double *d_scale);
int lsm_standardize_transform(lsm_ctx *ctx, const double *d_X, int32_t n, int32_t F, const double *d_mean,
                              const double *d_scale, double *d_out);

/* Multinomial logistic regression readout on the device (train_classifier.py:36-47: LogisticRegression(max_iter=1000).fit /
 * .predict).  scikit-learn's lbfgs objective: mean multinomial log-loss + ||W||^2 / (2 C n), intercept unpenalised; L-BFGS until
 * max|grad| <= tol.  d_X: double[n][F] (standardised features), d_y: int32[n] class indices 0..n_classes-1 (3..64 classes; beyond 16 the
 * passes run per block of 16 classes).
 * h_coef: double[n_classes][F], h_intercept: double[n_classes] (host, scikit-learn's coef_ / intercept_ layout).
 * The unique optimum is reached to solver tolerance, not scikit-learn's iterates bit for bit: the parity bar is the reference's
 * own metric, test accuracy (tests/test_gpu_readout.py: within 0.5 points and >= 99 % identical predictions).                    */
int lsm_logreg_fit(lsm_ctx *ctx, const double *d_X, const int32_t *d_y, int32_t n, int32_t F, int32_t n_classes, double C_reg,
                   int32_t max_iter, double tol, double *h_coef, double *h_intercept, int32_t *h_n_iter);
int lsm_logreg_predict(lsm_ctx *ctx, const double *d_X, int32_t n, int32_t F, int32_t n_classes, const double *h_coef,
                       const double *h_intercept, int32_t *d_pred);

/* Diagnostic: measured ceiling of the fp64 pipe on this device, in 1e9 DADD/DMUL lane-operations per
 * second (independent register chains, 8 warps per scheduler).  K1's roofline denominator in bench.py. */
int lsm_fp64_peak_gops(lsm_ctx *ctx, double *h_out);

#ifdef __cplusplus
}
#endif
#endif /* LSM_B200_H */
