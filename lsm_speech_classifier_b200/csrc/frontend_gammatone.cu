// K1 — gammatone filterbank + dB + min-max + zoom + 4-threshold hysteresis encoder (device code: gammatone_core.cuh).
//
// Kernels in this file:
//   gammatone_encode_kernel<MAXT, MINB, FNPT, LEAN>  persistent, one CTA per utterance in flight, one thread per channel.  FNPT = 0:
//       front end only; FNPT > 0: fused audio -> features (the CTA then simulates the utterance's reservoir, reservoir_core.cuh).
//       Filter modes: exact (gt_filter_exact: every fp64 operation an explicit __d*_rn intrinsic in the oracle's order -
//       oracle/lsm_oracle.c gammatone_energy / db_normalise_zoom / hysteresis_encode_f64 - 35 operations per channel-sample) and,
//       by default, speculative (gt_filter_fast: the same cascade in 13 FMAs; spec_epilogue flags the utterances in which the
//       derived error bound could change a comparison, and those are filtered again exactly inside the kernel, so the spike
//       trains are the exact path's either way).  This kernel reads every PCM sample once, which makes it the zero-copy
//       kernel for pinned host buffers, and it serves every shape; the default shape runs pipeline_lanes.cu instead.
//   gammatone_energy_kernel + gammatone_encode_kernel in mode 2: the stand-alone front end (lane = utterance filter with
//       coefficients in uniform registers, then the encoder).
//   audit_kernel: both filters on the same utterance, their distance against the derived bound (lsm_frontend_audit).
// The filter is bound by the fp64 pipe, not by HBM: per utterance it reads 64 000 B of PCM and writes 51 200 B of spikes.
#include <stdlib.h>

#include <memory>
#include <new>

#include "gammatone_core.cuh"

namespace {

// Units are dispatched in blockIdx order: first the big ones (kLanesJ channels: one PCM stream and one conversion feed
// 52 DFMAs), then single-channel units.  A unit is one warp's serial work (0.8 ms for four channels even on an idle SM),
// so a grid of big units alone ends in a long ragged tail; the small units keep every SM's pipes full to the end.
// The unit index is blockIdx.x: provably CTA-uniform for the compiler, which is what puts the coefficients in uniform registers.
__global__ void __launch_bounds__(32, 16) gammatone_energy_kernel(const __grid_constant__ EnergyArgs a)
{
    const int n_big = a.big_groups * (a.C / kLanesJ);
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        if (u < n_big) {
            const int units = a.C / kLanesJ;              // channel blocks per group (fastest index: PCM locality in L2)
            const int g = u / units;
            energy_unit<kLanesJ>(a, g, (u - g * units) * kLanesJ);
        } else {
            const int s = u - n_big;
            const int g = s / a.C;
            energy_unit<1>(a, a.big_groups + g, s - g * a.C);
        }
    }
}

// FNPT = 0: front end only (spike trains to global memory).  FNPT > 0: fused audio -> features: after
// encoding an utterance the same CTA simulates its reservoir (FNPT neurons per thread, blockDim == C) with
// the spikes handed over as bits in shared memory; the reservoir phase is latency/issue bound and uses
// almost no fp64, so it hides under the filter phases of the other CTAs resident on the SM.
template <int MAXT, int MINB, int FNPT, bool LEAN>
__global__ void __launch_bounds__(MAXT, MINB) gammatone_encode_kernel(const GtArgs a, int *next_utt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_x = reinterpret_cast<double *>(smem_raw);            // [2][kChunkBlocks*hop] PCM as fp64, shifted by kSkew
    __shared__ double s_red[6 * 8];
    __shared__ double s_out[6];
    __shared__ float s_xm[8];
    __shared__ int s_utt;
    __shared__ int s_cnt3[5];

    for (;;) {
        // dynamic work distribution: utterances are handed out one at a time, so every SM stays busy to the end
        if (threadIdx.x == 0) {
            const int i = atomicAdd(next_utt, 1);
            const int n = a.utt_count ? min(*a.utt_count, a.B) : a.B;
            s_utt = i < n ? (a.utt_list ? a.utt_list[i] : i) : -1;
        }
        __syncthreads();
        const int utt = s_utt;
        if (utt < 0) break;
        const PcmRow pcm = {a.pcm ? a.pcm + (size_t)utt * a.L : nullptr, a.pcm16 ? a.pcm16 + (size_t)utt * a.L : nullptr};
        // the utterance's [ncols][C] plane: energies -> dB -> normalised values, in place.  Mode 2 works directly on the
        // utterance's slice of the energy buffer, the other modes on this CTA's scratch plane (L2-resident).
        double *plane = a.mode == 2 ? const_cast<double *>(a.energy_in) + (size_t)utt * a.ncols * a.C
                                    : a.scratch + (size_t)blockIdx.x * a.ncols * a.C;
        bool settled = false;
        if (a.mode != 0) {
            float xm = 0.0f;
            if (a.mode == 1) {
                gt_filter_fast(a, pcm, s_x, plane, xm);
                // peak level of the utterance: every thread staged a share of the samples
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, o));
                if ((threadIdx.x & 31) == 0) s_xm[threadIdx.x >> 5] = xm;
                __syncthreads();
                xm = s_xm[0];
                for (int w = 1; w < (int)(blockDim.x >> 5); ++w) xm = fmaxf(xm, s_xm[w]);
            } else {
                xm = a.xmax_in[utt];
            }
            const bool near = spec_epilogue<FNPT, 0>(a, utt, plane, threadIdx.x, blockDim.x, xm, s_red, s_out, smem_raw);
            // some comparison of this utterance is too close to call on the speculative plane: filter it again, exactly
            settled = !__syncthreads_or(near ? 1 : 0);
            if (!settled && threadIdx.x == 0) atomicAdd(a.reruns, 1);
        }
        if (!settled) {
            double tmax = -INFINITY, tmin = INFINITY;
            gt_filter_exact(a, pcm, s_x, plane, tmax, tmin);
            exact_epilogue<FNPT>(a, utt, plane, tmax, tmin, s_red, s_out, smem_raw);
        }
        if (FNPT > 0) {
            __syncthreads();   // bits complete and visible
            reservoir_simulate<(FNPT > 0 ? FNPT : 4), LEAN>(a.res, utt, smem_raw, s_cnt3, threadIdx.x, blockDim.x);
        }
        __syncthreads();   // plane, shared memory and s_utt are reused by the next utterance
    }
}


// ---- Warp-specialised form of the fused kernel (default shape: 128 channels, 8 neurons per thread, speculative mode).
//      In gammatone_encode_kernel the same 128 threads take turns between the filter (fp64-pipe bound) and the reservoir
//      (issue / latency bound), so whenever some of an SM's CTAs are in their reservoir phase the fp64 pipe runs short of
//      filter warps.  Here a CTA carries both kinds of work side by side:
//        threads   0..127  FILTER group (named barrier 1): lane = channel; gt_filter_fast and the speculative encoder epilogue
//                          (both fp64 work) on utterance after utterance; the spike train goes, as ballot words, into one of two
//                          bit-plane buffers in shared memory
//        threads 128..255  UNIT group (named barrier 2): reservoir (8 neurons per thread) and readout on the bit plane the
//                          filter group finished last
//      The groups hand bit planes over through named barriers (full[slot]: filter arrives, unit syncs; empty[slot]: the other
//      way round), two deep.  Utterances the derived bound cannot settle go to a work list; the exact pass
//      (gammatone_encode_kernel, mode 0, over that list) follows on the same stream and overwrites their rows.
constexpr int kWsFilter = 128, kWsUnit = 128, kWsThreads = kWsFilter + kWsUnit;
constexpr int kBarFilter = 1, kBarUnit = 2, kBarFull = 3 /* +slot */, kBarEmpty = 5 /* +slot */;

__device__ __forceinline__ void bar_arrive(int id, int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <bool LEAN>
__global__ void __launch_bounds__(kWsThreads, 3) gammatone_ws_kernel(const GtArgs a, int *next_utt, const int bits_off, const int bits_bytes,
                                                                     const int unit_off)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_x = reinterpret_cast<double *>(smem_raw);            // filter group: [2][kChunkBlocks*hop] PCM as fp64
    __shared__ double s_red[6 * 8];
    __shared__ double s_out[6];
    __shared__ float s_xm[4];
    __shared__ int s_utt_f;
    __shared__ int s_slot_utt[2];
    __shared__ int s_cnt3[5];

    if (threadIdx.x < kWsFilter) {
        const int tid = threadIdx.x;
        double *plane = a.scratch + (size_t)blockIdx.x * a.ncols * a.C;
        int i = 0;
        for (;; ++i) {
            const int slot = i & 1;
            if (tid == 0) {
                const int k = atomicAdd(next_utt, 1);
                s_utt_f = k < a.B ? k : -1;
            }
            bar_sync(kBarFilter, kWsFilter);
            const int utt = s_utt_f;
            bar_sync(kBarFilter, kWsFilter);                         // everyone has read s_utt_f before it is written again
            if (utt < 0) {
                if (i >= 2) bar_sync(kBarEmpty + slot, kWsThreads);
                if (tid == 0) s_slot_utt[slot] = -1;
                __threadfence_block();
                bar_arrive(kBarFull + slot, kWsThreads);
                break;
            }
            const PcmRow pcm = {a.pcm ? a.pcm + (size_t)utt * a.L : nullptr, a.pcm16 ? a.pcm16 + (size_t)utt * a.L : nullptr};
            float xm = 0.0f;
            gt_filter_fast<kBarFilter>(a, pcm, s_x, plane, xm, tid, kWsFilter);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, o));
            if ((tid & 31) == 0) s_xm[tid >> 5] = xm;
            bar_sync(kBarFilter, kWsFilter);
            xm = fmaxf(fmaxf(s_xm[0], s_xm[1]), fmaxf(s_xm[2], s_xm[3]));
            if (i >= 2) bar_sync(kBarEmpty + slot, kWsThreads);      // the unit group is done with this slot's bit plane
            const bool near = a.ws_debug == 2 ? false
                : spec_epilogue<8, kBarFilter>(a, utt, plane, tid, kWsFilter, xm, s_red, s_out, smem_raw + bits_off + slot * bits_bytes);
            // too close to call on the speculative plane: finish it like the others, and list it for the exact pass
            unsigned any;
            asm volatile("{ .reg .pred p, q; setp.ne.u32 q, %1, 0; bar.red.or.pred p, %2, %3, q; selp.u32 %0, 1, 0, p; }"
                         : "=r"(any) : "r"((unsigned)near), "r"(kBarFilter), "r"(kWsFilter) : "memory");
            if (tid == 0) {
                if (any) {
                    a.rerun_list[1 + atomicAdd(a.rerun_list, 1)] = utt;
                    atomicAdd(a.reruns, 1);
                }
                s_slot_utt[slot] = utt;
            }
            __threadfence_block();
            bar_arrive(kBarFull + slot, kWsThreads);                 // bit plane and utterance index are the unit group's
        }
        // the unit group's release after the last real item (i - 1) has no taker yet: take it, so no barrier is left half-way
        if (i >= 1) bar_sync(kBarEmpty + ((i - 1) & 1), kWsThreads);
    } else {
        const int tid = threadIdx.x - kWsFilter;
        for (int i = 0;; ++i) {
            const int slot = i & 1;
            bar_sync(kBarFull + slot, kWsThreads);
            const int utt = s_slot_utt[slot];
            if (utt < 0) break;
            if (a.ws_debug == 0)
                reservoir_simulate<8, LEAN, false, kBarUnit>(a.res, utt, smem_raw + unit_off, s_cnt3, tid, kWsUnit, 0,
                                                             reinterpret_cast<unsigned *>(smem_raw + bits_off + slot * bits_bytes));
            bar_sync(kBarUnit, kWsUnit);                             // every unit thread is done with the bit plane and the staging area
            bar_arrive(kBarEmpty + slot, kWsThreads);
        }
    }
}

// ---- audit: both arrangements of the filter on the same utterance (diagnostic; lsm_frontend_audit).  Per utterance:
//      out[0] largest |dB_speculative - dB_exact| over all cells         out[1] largest such distance / its cell's bound
//      out[2] |max_speculative - max_exact| / its bound                  out[3] |min_speculative - min_exact| / its bound
//      out[4] bound on the maximum (dB)    out[5] bound on the minimum (dB)    out[6] max |sample|    out[7] range (dB)
template <int MAXT>
__global__ void __launch_bounds__(MAXT) audit_kernel(const GtArgs a, double *scratch2, double *out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_x = reinterpret_cast<double *>(smem_raw);
    __shared__ double s_red[6 * 8];
    __shared__ double s_out[6];
    __shared__ float s_xm[8];
    const int ch = threadIdx.x;
    const bool live = ch < a.C;
    double *plane_s = a.scratch + (size_t)blockIdx.x * a.ncols * a.C;
    double *plane_e = scratch2 + (size_t)blockIdx.x * a.ncols * a.C;
    for (int utt = blockIdx.x; utt < a.B; utt += gridDim.x) {
        const PcmRow pcm = {a.pcm ? a.pcm + (size_t)utt * a.L : nullptr, a.pcm16 ? a.pcm16 + (size_t)utt * a.L : nullptr};
        float xm = 0.0f;
        gt_filter_fast(a, pcm, s_x, plane_s, xm);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, o));
        if ((threadIdx.x & 31) == 0) s_xm[threadIdx.x >> 5] = xm;
        __syncthreads();
        xm = s_xm[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) xm = fmaxf(xm, s_xm[w]);
        const SpecStats st = spec_db_pass<0>(a, plane_s + ch, ch, live, threadIdx.x, blockDim.x, xm, s_red, s_out);
        double tmax = -INFINITY, tmin = INFINITY;
        gt_filter_exact(a, pcm, s_x, plane_e, tmax, tmin);
        group_maxmin<1, 0>(&tmax, &tmin, threadIdx.x, blockDim.x, s_red, s_out);
        const double mn_e = fmax(tmin, sub64(tmax, 80.0));
        double vmax[2] = {0.0, 0.0}, vmin[2] = {0.0, 0.0};
        if (live) {
            const float kx = (float)(a.kappa[ch] * (double)xm) * a.bound_scale;
            bool bad = false;
            for (int c = 0; c < a.ncols; ++c) {
                const double xs = plane_s[(size_t)c * a.C + ch], xe = plane_e[(size_t)c * a.C + ch];
                const double d = fabs(xs - xe);
                const double er = (double)cell_err(xs, kx, bad);
                vmax[0] = fmax(vmax[0], d);
                vmax[1] = fmax(vmax[1], d / er);
            }
            if (bad) vmax[1] = INFINITY;
        }
        group_maxmin<2, 0>(vmax, vmin, threadIdx.x, blockDim.x, s_red, s_out);
        if (threadIdx.x == 0) {
            double *o = out + (size_t)utt * 8;
            o[0] = vmax[0]; o[1] = vmax[1];
            o[2] = fabs(st.mx - tmax) / (double)st.emx;
            o[3] = fabs(st.mn - mn_e) / (double)st.emn;
            o[4] = (double)st.emx; o[5] = (double)st.emn; o[6] = (double)xm; o[7] = st.mx - st.mn;
        }
        __syncthreads();
    }
}

}  // namespace

template <int MAXT, int MINB, int FNPT, bool LEAN>
static int k1_grid(lsm_ctx *ctx, int threads, size_t smem, int *per_sm)
{
    if (smem > 48 * 1024)
        LSM_CUDA(ctx, cudaFuncSetAttribute(gammatone_encode_kernel<MAXT, MINB, FNPT, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, gammatone_encode_kernel<MAXT, MINB, FNPT, LEAN>, threads, smem));
    return LSM_OK;
}

int lsm_gammatone_grid(lsm_ctx *ctx, const lsm_frontend_params *p, int *grid)
{
    const int threads = ((p->channels + 31) / 32) * 32;
    const size_t smem = sizeof(double) * 2 * kChunkBlocks * p->hop;
    int per_sm = 0, rc;
    if (threads > 128) rc = k1_grid<256, 2, 0, true>(ctx, threads, smem, &per_sm);
    else rc = k1_grid<128, 5, 0, true>(ctx, threads, smem, &per_sm);
    if (rc != LSM_OK) return rc;
    if (per_sm < 1) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "gammatone kernel does not fit on an SM (hop %d)", p->hop);
    // persistent: every CTA resident, utterances handed out dynamically.  The scratch planes are sized by this grid; the fused
    // 128-channel kernel runs 6 CTAs per SM (below), so leave room for that.
    if (threads <= 128 && per_sm < 6) per_sm = 6;
    *grid = per_sm * ctx->sm_count;
    return LSM_OK;
}

void lsm_gammatone_fill_args(const lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes, double *d_spec_norm, GtArgs *out)
{
    const lsm_frontend_params &p = fe->p;
    GtArgs &a = *out;
    a.pcm = d_pcm; a.pcm16 = d_pcm ? nullptr : fe->next_pcm16; a.coefs = fe->d_coefs; a.kappa = fe->d_kappa;
    a.zoom_i0 = fe->d_zoom_i0; a.zoom_f = fe->d_zoom_f;
    a.scratch = fe->d_scratch; a.spikes = d_spikes; a.spec_norm = d_spec_norm;
    a.B = B; a.L = p.n_samples; a.C = p.channels; a.nwin = p.nwin; a.hop = p.hop; a.ncols = fe->ncols;
    a.nbins = p.n_bins; a.K = p.n_thresholds; a.R = p.redundancy;
    for (int k = 0; k < 8; ++k) { a.thr[k] = p.thresholds_desc[k]; a.lower[k] = p.lower_bounds[k]; }
    // the normalised-spectrogram dump is defined as the exact path's: asking for it selects the exact filter
    a.mode = d_spec_norm ? 0 : fe->mode;
    a.energy_in = nullptr; a.xmax_in = nullptr;
    a.utt_list = nullptr; a.utt_count = nullptr; a.rerun_list = nullptr; a.ws_debug = 0;
    a.spec_delta = fe->spec_delta;
    a.bound_scale = (float)fe->bound_scale;
    a.reruns = fe->d_counters + 64;
}

// d_counters layout (ints): [0,64) work counters, [64] re-execution count
static int next_counter(lsm_ctx *ctx, lsm_frontend *fe, cudaStream_t st, int **counter, GtArgs *a, int slot)
{
    int rc = lsm_frontend_order_before(ctx, fe, st, slot);
    if (rc != LSM_OK) return rc;
    a->scratch = fe->d_scratch + (slot > 0 ? (size_t)fe->grid * fe->ncols * fe->p.channels : 0);
    // work counter: one int per launch out of a small ring, so back-to-back launches on different streams do not share it
    const unsigned cslot = fe->counter_next++ % 64;
    *counter = fe->d_counters + cslot;
    LSM_CUDA(ctx, cudaMemsetAsync(*counter, 0, sizeof(int), st));
    return LSM_OK;
}

// ------------------------------------------------------------------------------------ lanes arrangement (stand-alone front end)
constexpr int kEnergyMaxUtt = 8192;     // utterances per pass of the two-kernel speculative path (energy buffer = 100 KB each)

// Can the speculative filter run lane = utterance?  8-sample steps must tile the window phases, rows must be 16-byte aligned.
bool lsm_lanes_eligible(const lsm_frontend *fe, const void *d_pcm)
{
    const lsm_frontend_params &p = fe->p;
    if (getenv("LSM_NO_LANES") || !d_pcm) return false;
    const int r_old = p.nwin - 2 * p.hop;
    return p.kind == LSM_FILTERBANK_GAMMATONE && fe->mode == LSM_FILTER_SPECULATIVE && p.hop % 8 == 0 && r_old % 8 == 0 &&
           p.n_samples % 4 == 0 && p.channels % kLanesJ == 0 && p.channels <= 256 && (((uintptr_t)d_pcm) & 15) == 0;
}

// device work list of the exact pass: [0] = count, then utterance indices; two lists (one per launch lane)
int lsm_frontend_ensure_rerun(lsm_ctx *ctx, lsm_frontend *fe, int B)
{
    if (B <= fe->rerun_cap) return LSM_OK;
    { const int rc = lsm_frontend_wait_idle(ctx, fe); if (rc != LSM_OK) return rc; }
    if (fe->d_rerun) { cudaFree(fe->d_rerun); fe->d_rerun = nullptr; fe->rerun_cap = 0; }
    if (cudaMalloc((void **)&fe->d_rerun, 2 * sizeof(int) * ((size_t)B + 1)) != cudaSuccess) LSM_FAIL(ctx, LSM_ERR_NOMEM, "cudaMalloc for the re-execution list failed");
    fe->rerun_cap = B;
    return LSM_OK;
}

// energy planes [2][B][ncols][C] (one set per launch lane) + peak levels
int lsm_frontend_ensure_energy(lsm_ctx *ctx, lsm_frontend *fe, int B)
{
    if (B <= fe->energy_cap) return LSM_OK;
    { const int rc = lsm_frontend_wait_idle(ctx, fe); if (rc != LSM_OK) return rc; }     // nobody is still reading the old buffer
    if (fe->d_energy) { LSM_CUDA(ctx, cudaFree(fe->d_energy)); fe->d_energy = nullptr; fe->energy_cap = 0; }
    if (fe->d_xmax) { LSM_CUDA(ctx, cudaFree(fe->d_xmax)); fe->d_xmax = nullptr; }
    if (fe->d_pipe_sync) { LSM_CUDA(ctx, cudaFree(fe->d_pipe_sync)); fe->d_pipe_sync = nullptr; }
    const size_t bytes = 2 * sizeof(double) * (size_t)B * fe->ncols * fe->p.channels;
    if (cudaMalloc((void **)&fe->d_energy, bytes) != cudaSuccess) LSM_FAIL(ctx, LSM_ERR_NOMEM, "cudaMalloc(%zu) for the energy planes failed", bytes);
    if (cudaMalloc((void **)&fe->d_xmax, 2 * sizeof(float) * (size_t)B) != cudaSuccess) LSM_FAIL(ctx, LSM_ERR_NOMEM, "cudaMalloc for the peak levels failed");
    // pipeline kernel: per lane [utterance counter, 3 spare, one completion counter per 32-utterance group]; last: the error flag
    const size_t n_sync = 2 * ((size_t)B / 32 + 8) + 8;
    if (cudaMalloc((void **)&fe->d_pipe_sync, sizeof(int) * n_sync) != cudaSuccess) LSM_FAIL(ctx, LSM_ERR_NOMEM, "cudaMalloc for the pipeline counters failed");
    LSM_CUDA(ctx, cudaMemset(fe->d_pipe_sync, 0, sizeof(int) * n_sync));
    fe->energy_cap = B;
    return lsm_frontend_ensure_rerun(ctx, fe, B);
}

// K1a: raw window energies of the speculative cascade for utterances [0, B) into fe->d_energy (set 0)
static int launch_energy(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, cudaStream_t st)
{
    const lsm_frontend_params &p = fe->p;
    int rc;
    if ((rc = lsm_frontend_order_before(ctx, fe, st)) != LSM_OK) return rc;
    if ((rc = lsm_frontend_ensure_energy(ctx, fe, B)) != LSM_OK) return rc;
    std::unique_ptr<EnergyArgs> ea_holder(new (std::nothrow) EnergyArgs);     // 12 KB of kernel parameters: not on the stack
    if (!ea_holder) LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
    EnergyArgs &ea = *ea_holder;
    ea.pcm = d_pcm; ea.energy = fe->d_energy; ea.xmax = fe->d_xmax;
    ea.B = B; ea.L = p.n_samples; ea.C = p.channels; ea.nwin = p.nwin; ea.hop = p.hop; ea.ncols = fe->ncols;
    memcpy(ea.coef, fe->h_lane_coef, sizeof(double) * 6 * p.channels);
    const int groups = (B + 31) / 32;
    // Units share an SM's fp64 pipes evenly, so an SM's finishing time is proportional to the work placed on it: hand every SM
    // the same number of big units (a multiple of the SM count; blocks are dealt round-robin) and cut what is left over into
    // single-channel units, which balance four times finer (all-big grid 4.8 ms, balanced 4.1 ms per 2400 utterances).
    const int upg = p.channels / kLanesJ;                       // big units per group
    const long long big_units = (long long)groups * upg;
    const int big_groups = (int)((big_units / ctx->sm_count) * ctx->sm_count / upg);
    ea.big_groups = big_groups;
    ea.n_units = big_groups * upg + (groups - big_groups) * p.channels;
    gammatone_energy_kernel<<<ea.n_units, 32, 0, st>>>(ea);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return LSM_OK;
}

int lsm_launch_gammatone(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes,
                         double *d_spec_norm, cudaStream_t st)
{
    const lsm_frontend_params &p = fe->p;
    const bool lanes = !d_spec_norm && lsm_lanes_eligible(fe, d_pcm);
    if (lanes && B > kEnergyMaxUtt) {
        const size_t spk_per = (size_t)p.channels * p.redundancy * p.n_bins * p.n_thresholds;
        for (int off = 0; off < B; off += kEnergyMaxUtt) {
            const int n = B - off < kEnergyMaxUtt ? B - off : kEnergyMaxUtt;
            const int rc = lsm_launch_gammatone(ctx, fe, d_pcm + (size_t)off * p.n_samples, n, d_spikes + off * spk_per, nullptr, st);
            if (rc != LSM_OK) return rc;
        }
        return LSM_OK;
    }
    GtArgs a;
    lsm_gammatone_fill_args(fe, d_pcm, B, d_spikes, d_spec_norm, &a);
    memset(&a.res, 0, sizeof(a.res));
    if (lanes && B > 0) {
        const int rc = launch_energy(ctx, fe, d_pcm, B, st);
        if (rc != LSM_OK) return rc;
        a.mode = 2;
        a.energy_in = fe->d_energy;
        a.xmax_in = fe->d_xmax;
    }
    const int threads = ((p.channels + 31) / 32) * 32;
    const size_t smem = sizeof(double) * 2 * kChunkBlocks * p.hop;
    const int grid = B < fe->grid ? B : fe->grid;
    if (grid <= 0) return LSM_OK;
    int *counter, rc;
    const int slot = lanes ? -1 : (int)(fe->slot_next++ & 1u);
    if ((rc = next_counter(ctx, fe, st, &counter, &a, slot)) != LSM_OK) return rc;
    if (threads > 128) gammatone_encode_kernel<256, 2, 0, true><<<grid, threads, smem, st>>>(a, counter);
    else gammatone_encode_kernel<128, 5, 0, true><<<grid, threads, smem, st>>>(a, counter);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st, slot);
}

// Can this (front end, reservoir) pair run as one fused kernel?  Gammatone, no redundancy, one thread
// per channel with whole warps, and a reservoir whose padded width is channels x {4, 8, 16} neurons per thread.
int lsm_fused_npt(const lsm_frontend *fe, const lsm_reservoir *res)
{
    const lsm_frontend_params &p = fe->p;
    if (getenv("LSM_NO_FUSE") || res->w64) return 0;      // strict reservoirs run as their own kernel
    if (p.kind != LSM_FILTERBANK_GAMMATONE || p.redundancy != 1 || (p.channels & 31) || p.channels > 256) return 0;
    if (res->p.num_inputs != p.channels || res->p.num_steps != p.n_bins * p.n_thresholds) return 0;
    if (res->n_pad % p.channels) return 0;
    const int npt = res->n_pad / p.channels;
    if (p.channels == 256) return npt == 4 ? 4 : 0;
    return (npt == 8 || npt == 16) ? npt : 0;
}

// launch == false: only report the resident grid (one wave) of this variant through *wave
template <int MAXT, int MINB, int FNPT>
static int launch_fused_t(lsm_ctx *ctx, lsm_frontend *fe, const lsm_reservoir *res, GtArgs &a, int threads, size_t smem,
                          cudaStream_t st, bool launch, int *wave, int max_grid, int forced_slot)
{
    int per_sm = 0, rc;
    if (res->lean) rc = k1_grid<MAXT, MINB, FNPT, true>(ctx, threads, smem, &per_sm);
    else rc = k1_grid<MAXT, MINB, FNPT, false>(ctx, threads, smem, &per_sm);
    if (rc != LSM_OK) return rc;
    if (per_sm < 1) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "fused kernel does not fit on an SM (%zu B shared memory)", smem);
    int grid = per_sm * ctx->sm_count;
    if (grid > fe->grid) grid = fe->grid;      // the dB scratch plane is sized for fe->grid CTAs
    if (wave) *wave = grid;
    if (!launch) return LSM_OK;
    if (grid > a.B) grid = a.B;
    if (max_grid > 0 && grid > max_grid) grid = max_grid;
    int *counter;
    const int slot = forced_slot != -2 ? forced_slot : (int)(fe->slot_next++ & 1u);
    if ((rc = next_counter(ctx, fe, st, &counter, &a, slot)) != LSM_OK) return rc;
    if (res->lean) gammatone_encode_kernel<MAXT, MINB, FNPT, true><<<grid, threads, smem, st>>>(a, counter);
    else gammatone_encode_kernel<MAXT, MINB, FNPT, false><<<grid, threads, smem, st>>>(a, counter);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st, slot);
}

// The lane = channel fused kernel.  a: filled by lsm_gammatone_fill_args + lsm_reservoir_fill_args (mode / work list set by the
// caller).  max_grid > 0 caps the grid (exact pass over a short work list); forced_slot = -2: alternate the scratch slots.
int lsm_launch_fused_args(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, GtArgs &a, cudaStream_t st, bool launch, int *wave,
                          int max_grid, int forced_slot)
{
    const int npt = lsm_fused_npt(fe, res);
    if (!npt) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "this front end / reservoir pair cannot run fused");
    const int threads = fe->p.channels;
    size_t smem = sizeof(double) * 2 * kChunkBlocks * fe->p.hop;
    const size_t smem_res = lsm_res_smem_bytes(a.res.T, a.res.CW, threads * npt, a.res.N);
    if (smem_res > smem) smem = smem_res;
    if (threads == 256) return launch_fused_t<256, 2, 4>(ctx, fe, res, a, threads, smem, st, launch, wave, max_grid, forced_slot);
    // 6 CTAs per SM (80 registers: only the rare exact re-execution spills) fill the issue slots the reservoir phases
    // leave better than 5 (measured 6.48 vs 6.79 ms per launch)
    if (npt == 8) return launch_fused_t<128, 6, 8>(ctx, fe, res, a, threads, smem, st, launch, wave, max_grid, forced_slot);
    return launch_fused_t<128, 4, 16>(ctx, fe, res, a, threads, smem, st, launch, wave, max_grid, forced_slot);
}


// Can the pair run as the warp-specialised kernel?  The reference's default shape in speculative mode.
bool lsm_ws_eligible(const lsm_frontend *fe, const lsm_reservoir *res)
{
    // Opt-in (LSM_WS=1): measured on B200 at 6.0 ms per 2400-utterance step against 5.9 ms for the phases-in-turn kernel
    // (profiles/r2_ws_exp.md: the filter group alone runs at the three-register DFMA ceiling, 4.6 ms, but the encoder epilogue
    // and the reservoir cost 0.6 + 0.8 ms wherever they run - the SM's issue / operand bandwidth is shared, not idle).
    if (!getenv("LSM_WS")) return false;
    return fe->p.channels == kWsFilter && lsm_fused_npt(fe, res) == 8 && fe->mode == LSM_FILTER_SPECULATIVE;
}

// The warp-specialised kernel on utterances [0, B) + the exact pass over the utterances it could not settle.
template <bool LEAN>
static int launch_ws_t(lsm_ctx *ctx, lsm_frontend *fe, GtArgs &a, cudaStream_t st, int *wave, bool launch, int *slot_out)
{
    int rc;
    const size_t smem_f = (sizeof(double) * 2 * kChunkBlocks * fe->p.hop + 127) & ~(size_t)127;
    const size_t bits = (sizeof(unsigned) * (size_t)a.res.T * a.res.CW + 127) & ~(size_t)127;
    const size_t smem_u = lsm_res_smem_bytes(a.res.T, a.res.CW, kWsUnit * 8, a.res.N);
    const size_t smem = smem_f + 2 * bits + smem_u;
    int per_sm = 0;
    LSM_CUDA(ctx, cudaFuncSetAttribute(gammatone_ws_kernel<LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gammatone_ws_kernel<LEAN>, kWsThreads, smem));
    if (per_sm < 1) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "warp-specialised kernel does not fit on an SM (%zu B shared memory)", smem);
    int grid = per_sm * ctx->sm_count;
    if (grid > fe->grid) grid = fe->grid;                  // one scratch plane per CTA out of a slot sized for fe->grid planes
    if (wave) *wave = 2 * grid;                            // utterances in flight: one per group
    if (!launch) return LSM_OK;
    if (grid > a.B) grid = a.B;
    if ((rc = lsm_frontend_ensure_rerun(ctx, fe, a.B)) != LSM_OK) return rc;
    int *counter;
    const int slot = (int)(fe->slot_next++ & 1u);
    *slot_out = slot;
    if ((rc = next_counter(ctx, fe, st, &counter, &a, slot)) != LSM_OK) return rc;
    a.rerun_list = fe->d_rerun + (size_t)slot * (fe->rerun_cap + 1);
    LSM_CUDA(ctx, cudaMemsetAsync(a.rerun_list, 0, sizeof(int), st));
    gammatone_ws_kernel<LEAN><<<grid, kWsThreads, smem, st>>>(a, counter, (int)smem_f, (int)bits, (int)(smem_f + 2 * bits));
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st, slot);
}

int lsm_launch_fused_ws(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, GtArgs &a, cudaStream_t st, int *wave, bool launch)
{
    int rc, slot = 0;
    { const char *e = getenv("LSM_WS_DEBUG"); a.ws_debug = e ? atoi(e) : 0; }
    rc = res->lean ? launch_ws_t<true>(ctx, fe, a, st, wave, launch, &slot) : launch_ws_t<false>(ctx, fe, a, st, wave, launch, &slot);
    if (rc != LSM_OK || !launch) return rc;
    // exact pass over the device work list (typically empty; a few CTAs suffice), same scratch slot, same stream
    GtArgs x = a;
    x.mode = 0; x.rerun_list = nullptr;
    x.utt_list = a.rerun_list + 1; x.utt_count = a.rerun_list;
    return lsm_launch_fused_args(ctx, fe, res, x, st, true, nullptr, 8, slot);
}

// row0: index of utterance 0 of this launch within the API call (offsets the fused all-gather's destination rows)
int lsm_launch_fused(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B,
                     uint8_t *d_spikes_or_null, uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st,
                     long long row0)
{
    if (B <= 0) return LSM_OK;
    GtArgs a;
    lsm_gammatone_fill_args(fe, d_pcm, B, d_spikes_or_null, nullptr, &a);
    lsm_reservoir_fill_args(res, nullptr, B, feature_mask, nan_to_num, d_features, nullptr, &a.res);
    a.res.gather_row0 += row0;
    if (a.mode == 1 && lsm_ws_eligible(fe, res)) return lsm_launch_fused_ws(ctx, fe, res, a, st, nullptr, true);
    return lsm_launch_fused_args(ctx, fe, res, a, st, true, nullptr, 0, -2);
}

// CTAs resident at once for the fused kernel of this pair = utterances per wave (0 if not fusable)
int lsm_fused_wave(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res)
{
    int wave = 0;
    if (!lsm_fused_npt(fe, res)) return 0;
    GtArgs a;
    lsm_gammatone_fill_args(fe, nullptr, 1, nullptr, nullptr, &a);
    lsm_reservoir_fill_args(res, nullptr, 1, 1u, 0, nullptr, nullptr, &a.res);
    if (lsm_launch_fused_args(ctx, fe, res, a, nullptr, false, &wave, 0, -2) != LSM_OK) return 0;
    return wave;
}

// ------------------------------------------------------------------------------------ audit
int lsm_launch_audit(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, double *d_out, cudaStream_t st)
{
    const lsm_frontend_params &p = fe->p;
    GtArgs a;
    lsm_gammatone_fill_args(fe, d_pcm, B, nullptr, nullptr, &a);
    memset(&a.res, 0, sizeof(a.res));
    const int threads = ((p.channels + 31) / 32) * 32;
    const size_t smem = sizeof(double) * 2 * kChunkBlocks * p.hop;
    int grid = B < fe->grid ? B : fe->grid;
    if (grid <= 0) return LSM_OK;
    int rc;
    // both scratch slots: exclusive use of the front end
    if ((rc = lsm_frontend_order_before(ctx, fe, st, -1)) != LSM_OK) return rc;
    a.scratch = fe->d_scratch;
    double *scratch2 = fe->d_scratch + (size_t)fe->grid * fe->ncols * p.channels;
    if (threads > 128) audit_kernel<256><<<grid, threads, smem, st>>>(a, scratch2, d_out);
    else audit_kernel<128><<<grid, threads, smem, st>>>(a, scratch2, d_out);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st, -1);
}
