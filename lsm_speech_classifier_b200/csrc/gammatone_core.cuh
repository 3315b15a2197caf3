// Device code shared by the gammatone kernels (frontend_gammatone.cu: lane = channel kernels, energy kernel, audit kernel;
// pipeline_lanes.cu: the warp-specialised audio -> features kernel): the two filter arrangements, the two encoder epilogues
// and the derived error bound that ties them together.
//
// Replaces, per utterance, /root/reference/create_dataset.py:148-158 on the gammatone branch:
//   gtgram.gtgram(...)                          :51-58   (gammatone==1.0.3: erb_filterbank = 4 cascaded
//                                                          scipy.signal.lfilter biquads per channel, /gain,
//                                                          square, sqrt(mean) over 400-sample windows, hop 160)
//   20*log10(spec+1e-9), floor at max-80         :59-60
//   per-utterance min-max normalisation          :62-67
//   scipy.ndimage.zoom(order=1) 98 -> 100 bins   :69-78
//   convert_spectrogram_to_spikes_hysteresis     :81-98
//   create_pure_redundancy                       :101-104
#pragma once

#include "reservoir_core.cuh"

constexpr int kChunkBlocks = 8;   // hop-blocks of PCM staged per shared-memory buffer (lane = channel kernels)

// x / g with g a per-thread constant: q0 = RN(x*r), e = x - g*q0 (exact, FMA), q = RN(q0 + e*r) is the
// correctly rounded quotient when r = RN(1/g) (Markstein's theorem) as long as the residual does not
// underflow, i.e. for |x| >= 2^-900 (tests/test_oracle_frontend.py checks it against IEEE division).
// No guard is needed for smaller |x|: the only consumer is the square v*v, and with |x| < 2^-900 and
// |1/g| < 2^60 both the exact quotient and this one are far below 2^-538, so the square is exactly +0
// either way.  |x| >= 2^900 cannot occur for float32 PCM.
__device__ __forceinline__ double div_by_const(double x, double g, double r)
{
    const double q0 = mul64(x, r);
    const double e = __fma_rn(-g, q0, x);
    return __fma_rn(e, r, q0);
}

struct GtArgs {
    const float *pcm;       // [B][L] float32 samples, or null when pcm16 is given
    const int16_t *pcm16;   // [B][L] PCM16 samples (optional alternative input)
    const double *coefs;    // [C][10]
    const double *kappa;    // [C] bound on |amplitude(speculative) - amplitude(exact)| per unit max |sample| (error_bound.cu)
    const int32_t *zoom_i0; // [nbins]
    const double *zoom_f;   // [nbins]
    double *scratch;        // [grid][ncols][C]
    uint8_t *spikes;        // [B][C*R][nbins*K]
    double *spec_norm;      // optional [B][C][nbins]
    int B, L, C, nwin, hop, ncols, nbins, K, R;
    const double *energy_in; // mode 2: [B][ncols][C] raw energy sums from an energy kernel (also this utterance's plane)
    const float *xmax_in;   // mode 2: [B] max |sample| of each utterance
    int mode;               // 0 = exact filter only; 1 = speculative filter (lane = channel) in this kernel, exact re-execution
                            // of near-ties; 2 = as 1, but the speculative energies were computed by an energy kernel
    double spec_delta;      // extra dB margin of the near-tie test (on top of the derived bound)
    float bound_scale;      // multiplier of the derived bound (1; other values are diagnostics)
    int *reruns;            // number of utterances filtered twice (speculative mode)
    const int *utt_list;    // optional indirection: work item i is utterance utt_list[i], i < *utt_count (exact re-execution pass)
    const int *utt_count;
    int ws_debug;           // warp-specialised kernel, timing experiments: 1 = no reservoir, 2 = filter group alone (LSM_WS_DEBUG)
    int *rerun_list;        // kernels without the exact path: utterances whose speculative plane was too close to call; [0] = count
    double thr[8], lower[8];
    ResArgs res;            // fused mode only: the reservoir this utterance's spikes feed
};

// One biquad step (scipy.signal.lfilter direct form II transposed, b2 = 0): y = z0 + b0*x;
// z0' = (z1 + x*b1) - y*a1; z1' = -(y*a2).
#define LSM_BIQUAD(y, x, z0, z1, b1)                           \
    {                                                          \
        y = add64(z0, mul64(b0, x));                           \
        z0 = sub64(add64(z1, mul64(x, b1)), mul64(y, a1));     \
        z1 = mul64(y, na2);                                    \
    }

// The four cascaded stages are software-skewed: in one loop iteration stage k works on sample
// s + (3 - k), so the four recurrences are independent instruction chains (ILP 4) while every
// stage still performs exactly the reference's operations in the reference's order.  The stage-4
// output of iteration s is the cascade output for sample s.
constexpr int kSkew = 3;

// One utterance's samples: float32 (the load_audio_file contract, create_dataset.py:22-36) or the PCM16 a WAV file holds.
// (double)int16 * 2^-15 is exactly the double of float32(int16 / 32768), what librosa/soundfile hand to the reference, so both
// forms give the same bits downstream (the int16 -> float32 step of the ingest, done where the sample is used).
struct PcmRow {
    const float *f;
    const int16_t *h;
    __device__ __forceinline__ double at(int i) const
    {
        return h ? __dmul_rn((double)__ldcs(h + i), 0x1p-15) : (double)__ldcs(f + i);
    }
};

// PCM -> fp64 in shared memory; buffer element i of chunk k holds sample k*chunk + i + kSkew.  xm collects max |sample|
// (float32 holds every sample exactly, so the maximum is exact; NaN samples are skipped here and caught by the bound test).
// Barrier of a thread group: the whole CTA (BAR = 0, __syncthreads) or the nthr threads that own named barrier BAR.
template <int BAR>
__device__ __forceinline__ void group_sync(int nthr)
{
    if (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(BAR), "r"(nthr) : "memory");
}

__device__ __forceinline__ void stage_pcm(double *dst, const PcmRow pcm, int base, int chunk, int L, float &xm, int tid, int nthr)
{
    // streaming loads (evict-first): every PCM sample is read once and must not push the CTAs' scratch planes out of L2
    for (int i = tid; i < chunk; i += nthr) {
        const double v = (base + i < L) ? pcm.at(base + i) : 0.0;
        dst[i] = v;
        xm = fmaxf(xm, fabsf((float)v));
    }
}

// sqrt(mean) -> dB of one finished window (create_dataset.py:59) into the CTA's plane
__device__ __forceinline__ void emit_db(double y2w, double *plane, int col, int C, int ch, double &tmax, double &tmin)
{
    const double db = mul64(20.0, lsm_log10(add64(y2w, 1e-9)));
    plane[(size_t)col * C + ch] = db;
    tmax = fmax(tmax, db);
    tmin = fmin(tmin, db);
}

// ---- EXACT filter: the reference's operations in the reference's order (scipy lfilter x4, /gain, square,
//      left-to-right window sums).  35 fp64 operations per channel-sample, none of them fused.
__device__ __forceinline__ void gt_filter_exact(const GtArgs &a, const PcmRow pcm, double *s_x, double *plane, double &tmax,
                                                double &tmin)
{
    const int ch = threadIdx.x;
    const int C = a.C;
    const bool live = ch < C;
    const int hop = a.hop, nwin = a.nwin, ncols = a.ncols;
    const int chunk = kChunkBlocks * hop;
    const int r_old = nwin - 2 * hop;                 // phases at which window m-2 is still open
    const int n_used = (ncols - 1) * hop + nwin;      // samples the reference ever reads
    const int n_blocks = (n_used + hop - 1) / hop;
    const int n_chunks = (n_blocks + kChunkBlocks - 1) / kChunkBlocks;
    float xm_unused = 0.0f;

    // per-channel constants (scipy.signal.lfilter normalises b and a by a[0] = B0 first)
    double b0 = 0, b1_0 = 0, b1_1 = 0, b1_2 = 0, b1_3 = 0, a1 = 0, na2 = 0, gain = 1.0, rgain = 1.0;
    if (live) {
        const double *c = a.coefs + 10 * ch;
        const double a0 = c[6];
        b0 = __ddiv_rn(c[0], a0);
        b1_0 = __ddiv_rn(c[1], a0); b1_1 = __ddiv_rn(c[2], a0);
        b1_2 = __ddiv_rn(c[3], a0); b1_3 = __ddiv_rn(c[4], a0);
        a1 = __ddiv_rn(c[7], a0);
        na2 = -__ddiv_rn(c[8], a0);
        gain = c[9];
        rgain = __ddiv_rn(1.0, gain);
    }
    double z0_0 = 0, z0_1 = 0, z0_2 = 0, z0_3 = 0, z1_0 = 0, z1_1 = 0, z1_2 = 0, z1_3 = 0;
    double y1 = 0, y2 = 0, y3 = 0, y4 = 0;
    double acc_new = 0, acc_mid = 0, acc_old = 0;

    stage_pcm(s_x, pcm, kSkew, chunk, a.L, xm_unused, threadIdx.x, blockDim.x);
    if (live) {
        // prologue: iterations s = -3, -2, -1 fill the skewed pipeline (outputs belong to no sample)
#pragma unroll
        for (int s = 0; s < kSkew; ++s) {
            const double x = pcm.at(s);
            double t1, t2, t3;
            LSM_BIQUAD(t1, x, z0_0, z1_0, b1_0);
            LSM_BIQUAD(t2, y1, z0_1, z1_1, b1_1);
            LSM_BIQUAD(t3, y2, z0_2, z1_2, b1_2);
            LSM_BIQUAD(y4, y3, z0_3, z1_3, b1_3);
            y1 = t1; y2 = t2; y3 = t3;
        }
    }
    __syncthreads();

    for (int ck = 0; ck < n_chunks; ++ck) {
        const double *xs = s_x + (ck & 1) * chunk;
        // prefetch the next chunk into the other buffer while this one is filtered
        if (ck + 1 < n_chunks) stage_pcm(s_x + ((ck + 1) & 1) * chunk, pcm, (ck + 1) * chunk + kSkew, chunk, a.L, xm_unused, threadIdx.x, blockDim.x);
        if (live) {
            for (int bl = 0; bl < kChunkBlocks; ++bl) {
                const int m = ck * kChunkBlocks + bl;       // hop-block index = index of the window that starts here
                if (m >= n_blocks) break;
                const double *xb = xs + bl * hop;
                const int n_here = min(hop, n_used - m * hop);
                const int n_a = min(n_here, r_old);      // phases where windows m, m-1 and m-2 are all open
                // window m starts here: np.add.reduce begins with the first element, and 0.0 + e == e
                acc_new = 0.0;
#define LSM_SAMPLE(xin)                                                                   \
                {                                                                         \
                    double t1, t2, t3;                                                    \
                    LSM_BIQUAD(t1, xin, z0_0, z1_0, b1_0);  /* stage 1, sample s+3 */     \
                    LSM_BIQUAD(t2, y1, z0_1, z1_1, b1_1);   /* stage 2, sample s+2 */     \
                    LSM_BIQUAD(t3, y2, z0_2, z1_2, b1_2);   /* stage 3, sample s+1 */     \
                    LSM_BIQUAD(y4, y3, z0_3, z1_3, b1_3);   /* stage 4, sample s   */     \
                    y1 = t1; y2 = t2; y3 = t3;                                            \
                }
#pragma unroll 4
                for (int p = 0; p < n_a; ++p) {
                    LSM_SAMPLE(xb[p]);
                    const double v = div_by_const(y4, gain, rgain);
                    const double e = mul64(v, v);
                    acc_new = add64(acc_new, e);
                    acc_mid = add64(acc_mid, e);
                    acc_old = add64(acc_old, e);
                }
                // window m-2 complete: sqrt(mean) -> dB
                if (n_a == r_old && m >= 2) emit_db(__dsqrt_rn(__ddiv_rn(acc_old, (double)nwin)), plane, m - 2, C, ch, tmax, tmin);
#pragma unroll 4
                for (int p = n_a; p < n_here; ++p) {
                    LSM_SAMPLE(xb[p]);
                    const double v = div_by_const(y4, gain, rgain);
                    const double e = mul64(v, v);
                    acc_new = add64(acc_new, e);
                    acc_mid = add64(acc_mid, e);
                }
#undef LSM_SAMPLE
                acc_old = acc_mid;
                acc_mid = acc_new;
            }
        }
        __syncthreads();
    }
}

// ---- SPECULATIVE filter: the same cascade in a cheaper, mathematically equivalent arrangement - 13 fused
//      multiply-adds per channel-sample instead of 35 separate operations:
//        * every numerator is normalised to (1 + c_k z^-1); the common factor A0^4 / gain moves to the window level;
//        * direct form, y[n] = (x[n] + c_k x[n-1] - a2 y[n-2]) - a1 y[n-1]: three FMAs, one on the loop-carried path;
//        * one running energy sum; window m-2 = full(m-2) + full(m-1) + head(m) of hop-block sums.
//      Its dB plane differs from the exact one by rounding noise only, and by no more than the bound derived in
//      error_bound.cu; spec_epilogue flags every utterance in which that bound could change one of the
//      encoder's comparisons, and those utterances are filtered again by gt_filter_exact: the spike trains that leave the
//      kernel are the exact path's, byte for byte.
template <int BAR = 0>
__device__ __forceinline__ void gt_filter_fast(const GtArgs &a, const PcmRow pcm, double *s_x, double *plane, float &xm,
                                               const int tid = threadIdx.x, const int nthr = blockDim.x)
{
    const int ch = tid;
    const int C = a.C;
    const bool live = ch < C;
    const int hop = a.hop, nwin = a.nwin, ncols = a.ncols;
    const int chunk = kChunkBlocks * hop;
    const int r_old = nwin - 2 * hop;
    const int n_used = (ncols - 1) * hop + nwin;
    const int n_blocks = (n_used + hop - 1) / hop;
    const int n_chunks = (n_blocks + kChunkBlocks - 1) / kChunkBlocks;
    const bool pairs = !(hop & 1) && !(reinterpret_cast<uintptr_t>(s_x) & 15);

    double c1 = 0, c2 = 0, c3 = 0, c4 = 0, na1 = 0, na2 = 0;
    if (live) {
        const double *c = a.coefs + 10 * ch;
        const double a0 = c[6], A0 = c[0];
        c1 = c[1] / A0; c2 = c[2] / A0; c3 = c[3] / A0; c4 = c[4] / A0;
        na1 = -(c[7] / a0);
        na2 = -(c[8] / a0);
    }
    // stage k at iteration s works on sample s + 4 - k: p_k = its previous output, q_k = the one before
    double xp = 0, p1 = 0, q1 = 0, p2 = 0, q2 = 0, p3 = 0, q3 = 0, p4 = 0, q4 = 0;
    double acc = 0, full1 = 0, full2 = 0;

#define LSM_FAST_SAMPLE(xin)                                                       \
    {                                                                              \
        const double x_ = (xin);                                                   \
        const double n1 = fma(na1, p1, fma(na2, q1, fma(c1, xp, x_)));             \
        const double n2 = fma(na1, p2, fma(na2, q2, fma(c2, q1, p1)));             \
        const double n3 = fma(na1, p3, fma(na2, q3, fma(c3, q2, p2)));             \
        const double n4 = fma(na1, p4, fma(na2, q4, fma(c4, q3, p3)));             \
        xp = x_;                                                                   \
        q1 = p1; p1 = n1; q2 = p2; p2 = n2; q3 = p3; p3 = n3; q4 = p4; p4 = n4;    \
        acc = fma(n4, n4, acc);                                                    \
    }

    stage_pcm(s_x, pcm, kSkew, chunk, a.L, xm, tid, nthr);
    if (live) {
#pragma unroll
        for (int s = 0; s < kSkew; ++s) {
            const double x0 = pcm.at(s);
            xm = fmaxf(xm, fabsf((float)x0));
            LSM_FAST_SAMPLE(x0);
        }
        acc = 0.0;   // (already zero: stage 4 has seen no sample yet)
    }
    group_sync<BAR>(nthr);

    for (int ck = 0; ck < n_chunks; ++ck) {
        const double *xs = s_x + (ck & 1) * chunk;
        if (ck + 1 < n_chunks) stage_pcm(s_x + ((ck + 1) & 1) * chunk, pcm, (ck + 1) * chunk + kSkew, chunk, a.L, xm, tid, nthr);
        if (live) {
            for (int bl = 0; bl < kChunkBlocks; ++bl) {
                const int m = ck * kChunkBlocks + bl;
                if (m >= n_blocks) break;
                const double *xb = xs + bl * hop;
                const int n_here = min(hop, n_used - m * hop);
                const int n_a = min(n_here, r_old);
                acc = 0.0;
                if (pairs && !(n_a & 1)) {
                    // two samples per 16-byte shared-memory load (the buffers and an even hop keep xb 16-byte aligned)
                    const double2 *x2 = reinterpret_cast<const double2 *>(xb);
#pragma unroll 4
                    for (int p = 0; p < n_a / 2; ++p) { const double2 v = x2[p]; LSM_FAST_SAMPLE(v.x); LSM_FAST_SAMPLE(v.y); }
                    if (n_a == r_old && m >= 2) plane[(size_t)(m - 2) * C + ch] = (full2 + full1) + acc;
#pragma unroll 4
                    for (int p = n_a / 2; p < n_here / 2; ++p) { const double2 v = x2[p]; LSM_FAST_SAMPLE(v.x); LSM_FAST_SAMPLE(v.y); }
                    if (n_here & 1) LSM_FAST_SAMPLE(xb[n_here - 1]);
                } else {
#pragma unroll 8
                    for (int p = 0; p < n_a; ++p) LSM_FAST_SAMPLE(xb[p]);
                    // window m-2 complete: its raw energy sum; dB is taken in the epilogue, where the columns give ILP
                    if (n_a == r_old && m >= 2) plane[(size_t)(m - 2) * C + ch] = (full2 + full1) + acc;
#pragma unroll 8
                    for (int p = n_a; p < n_here; ++p) LSM_FAST_SAMPLE(xb[p]);
                }
                full2 = full1;
                full1 = acc;
            }
        }
        group_sync<BAR>(nthr);
    }
#undef LSM_FAST_SAMPLE
}

// ------------------------------------------------------------------------------------------------------------------
// Block-wide reductions for a group of NTHR threads that owns named barrier BAR (0 = the whole CTA, __syncthreads).
// s_red: double[6][8] scratch, s_out: double[6].  Every thread of the group gets all results.
// vmax[k] -> maximum over the group, vmin[k] -> minimum over the group, k < NV
template <int NV, int BAR>
__device__ __forceinline__ void group_maxmin(double *vmax, double *vmin, int tid, int nthr, double *s_red, double *s_out)
{
#pragma unroll
    for (int k = 0; k < NV; ++k) { vmax[k] = warp_max_f64(vmax[k]); vmin[k] = warp_min_f64(vmin[k]); }
    const int nw = nthr >> 5;
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) { s_red[(2 * k) * 8 + (tid >> 5)] = vmax[k]; s_red[(2 * k + 1) * 8 + (tid >> 5)] = vmin[k]; }
    }
    group_sync<BAR>(nthr);
    if (tid < 2 * NV) {
        const bool is_max = (tid & 1) == 0;
        double r = is_max ? -INFINITY : INFINITY;
        for (int w = 0; w < nw; ++w) r = is_max ? fmax(r, s_red[tid * 8 + w]) : fmin(r, s_red[tid * 8 + w]);
        s_out[tid] = r;
    }
    group_sync<BAR>(nthr);
#pragma unroll
    for (int k = 0; k < NV; ++k) { vmax[k] = s_out[2 * k]; vmin[k] = s_out[2 * k + 1]; }
}

// the four (K) Schmitt triggers of one channel for one time bin (create_dataset.py:90-94), bit k of `on` = trigger k
// KK = the number of triggers when known at compile time (4: the reference's), 0 = a.K
template <int KK = 0>
__device__ __forceinline__ void triggers_step(const GtArgs &a, double v, unsigned &on)
{
#pragma unroll
    for (int k = 0; k < (KK ? KK : 8); ++k) {
        if (KK || k < a.K) {
            const bool is_on = (on >> k) & 1u;
            if (!is_on && v > a.thr[k]) on |= (1u << k);
            else if (is_on && v < a.lower[k]) on &= ~(1u << k);
        }
    }
}

// spikes of time bin j: to the reservoir's bit plane in shared memory (fused kernels) and / or to X_spikes rows.
// ch = this thread's channel (0 .. C-1, a whole number of warps).
template <int FNPT, int KK = 0>
__device__ __forceinline__ void put_spikes(const GtArgs &a, unsigned on, int j, int ch, uint8_t *row0, unsigned char *smem_raw)
{
    const int K = KK ? KK : a.K;
    if (FNPT > 0) {
        // word (t, warp) = ballot over this warp's 32 channels
        unsigned *s_bits = reinterpret_cast<unsigned *>(smem_raw);
        const int CW = a.C >> 5;
#pragma unroll
        for (int k = 0; k < (KK ? KK : 8); ++k) {
            if (KK || k < K) {
                const unsigned word = __ballot_sync(0xffffffffu, (on >> k) & 1u);
                if ((ch & 31) == 0) s_bits[(j * K + k) * CW + (ch >> 5)] = word;
            }
        }
    }
    const int T = a.nbins * K;
    for (int r = 0; row0 && r < a.R; ++r) {
        uint8_t *row = row0 + (size_t)r * T + (size_t)j * K;
        if (K == 4) {
            // bytes k = 0..3 of column block j, little endian
            const unsigned w = (on & 1u) | ((on & 2u) << 7) | ((on & 4u) << 14) | ((on & 8u) << 21);
            __stcs(reinterpret_cast<unsigned *>(row), w);        // streaming store: written once, never read here
        } else {
            for (int k = 0; k < K; ++k) row[k] = (on >> k) & 1u;
        }
    }
}

// ---- EXACT epilogue: floor, min-max, zoom, encoder in the reference's operations (create_dataset.py:60-98)
template <int FNPT>
__device__ __forceinline__ void exact_epilogue(const GtArgs &a, int utt, double *plane, double tmax, double tmin,
                                               double *s_red, double *s_out, unsigned char *smem_raw)
{
    const int ch = threadIdx.x, C = a.C, ncols = a.ncols;
    group_maxmin<1, 0>(&tmax, &tmin, threadIdx.x, blockDim.x, s_red, s_out);
    const double mx = tmax, mn0 = tmin;
    const double floor_db = sub64(mx, 80.0);
    const double mn = fmax(mn0, floor_db);             // min of the clamped plane
    const bool degenerate = sub64(mx, mn) < 1e-8;       // create_dataset.py:64-65 -> all zeros
    const double den = add64(sub64(mx, mn), 1e-8);
    if (ch >= C) return;
    const int T = a.nbins * a.K;
    uint8_t *row0 = a.spikes ? a.spikes + ((size_t)utt * C * a.R + (size_t)ch * a.R) * T : nullptr;
    double *dump = a.spec_norm ? a.spec_norm + ((size_t)utt * C + ch) * a.nbins : nullptr;
    // normalise in place (own column of the plane only)
    for (int c = 0; c < ncols; ++c) {
        const double v = fmax(plane[(size_t)c * C + ch], floor_db);
        plane[(size_t)c * C + ch] = __ddiv_rn(sub64(v, mn), den);
    }
    unsigned on = 0;
    for (int j = 0; j < a.nbins; ++j) {
        double v;
        if (degenerate) v = 0.0;
        else if (ncols == a.nbins) v = plane[(size_t)j * C + ch];
        else {
            const int i0 = a.zoom_i0[j];
            const double f = a.zoom_f[j];
            v = mul64(plane[(size_t)i0 * C + ch], sub64(1.0, f));
            if (i0 + 1 < ncols) v = add64(v, mul64(plane[(size_t)(i0 + 1) * C + ch], f));
        }
        if (dump) dump[j] = v;
        triggers_step(a, v, on);
        put_spikes<FNPT>(a, on, j, ch, row0, smem_raw);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The derived bound (error_bound.cu).  For a cell whose speculative dB value is x_db, the exact path's
// value lies within cell_err(x_db, kx) decibels of it, where kx = kappa[channel] * max|sample| bounds the difference of the
// two paths' window amplitudes (absolute, worst case over all inputs of that peak level) and kMiscRel covers the window
// sums, square roots and scale factors (relative).  amp + 1e-9 = 10^(x_db/20), and |20 log10(a') - 20 log10(a)| <=
// (20 / ln 10) d / (a - d) for |a' - a| <= d < a.  Evaluated in float32 with everything rounded up by a percent; the
// caller flags the utterance (exact re-execution) when the relative error r is not small (or is NaN).
constexpr float kMiscRel = 6.0e-14f;        // > (nwin + 32) * 2^-53 for windows up to 500 samples
__device__ __forceinline__ float cell_err(double x_db, float kx, bool &bad)
{
    const float r = kx * exp2f(-0.16609640f * (float)x_db) + kMiscRel;      // d / a, with a = 10^(x/20)
    bad |= !(r < 0.25f);
    return 8.69f * 1.01f * __fdividef(r, 1.0f - r) + 1e-12f;
}

// 20 log10(sqrt(mean square) + 1e-9) for the speculative plane, without a branch: the library's sqrt and log10 each end their
// basic block with a slow-path test, which kept the seven columns a thread has in flight from overlapping and costs ~110 more
// instructions per value than this.  sqrt_rn_inline is exact for 2^-970 <= m < inf and returns 2 m below that (< 1e-290: nothing
// beside the 1e-9), lsm_log10 is the exact path's own log10 (< 1 ulp; argument >= 1e-9); non-finite energies stay non-finite so
// that the utterance is flagged as before.  Accuracy is inside what kMiscRel and cell_err's 1e-12 already allow.
__device__ __forceinline__ double spec_db(const double m)
{
    const double arg = sqrt_rn_inline(m) + 1e-9;
    const double db = 20.0 * lsm_log10(arg);
    return arg < 1.0e300 ? db : arg;
}

constexpr int kSpecCols = 5;                  // columns a thread has in flight in spec_db_pass (7 spill at 80 registers)

struct SpecStats {
    double mx, mn;          // speculative plane: maximum, minimum after the floor
    double floor_db, rden;  // mx - 80, 1 / (mx - mn + 1e-8)
    float emx, emn;         // bounds on |mx_exact - mx|, |mn_exact - mn|
    bool degenerate, near;
};

// First pass of the speculative epilogue for one thread's channel: window energy sums -> dB in place, and the block-wide
// quantities the normalisation needs together with rigorous intervals for the exact path's maximum and minimum:
//   mx_exact in [max(x - err), max(x + err)],  mn_exact in [max(min(x - err), floor - emx), max(min(x + err), floor + emx)].
template <int BAR>
__device__ __forceinline__ SpecStats spec_db_pass(const GtArgs &a, double *col, int ch, bool live, int tid, int nthr, float xmax,
                                                  double *s_red, double *s_out)
{
    const int C = a.C, ncols = a.ncols;
    double vmax[3] = {-INFINITY, -INFINITY, -INFINITY}, vmin[3] = {INFINITY, INFINITY, INFINITY};
    bool bad = false;
    if (live) {
        const double *c = a.coefs + 10 * ch;
        const double s = c[0] / c[6];
        const double G = (s * s) * (s * s) / c[9];
        const double g2n = G * G / (double)a.nwin;      // (A0^4 / gain)^2 / nwin: energy sum -> mean square of the real output
        const float kx = (float)(a.kappa[ch] * (double)xmax) * a.bound_scale;
        int c0 = 0;
        for (; c0 + kSpecCols <= ncols; c0 += kSpecCols) {
            double e[kSpecCols];
#pragma unroll
            for (int u = 0; u < kSpecCols; ++u) e[u] = __ldcg(col + (size_t)(c0 + u) * C);
#pragma unroll
            for (int u = 0; u < kSpecCols; ++u) {
                e[u] = spec_db(e[u] * g2n);
                const double er = (double)cell_err(e[u], kx, bad);
                vmax[0] = fmax(vmax[0], e[u]); vmax[1] = fmax(vmax[1], e[u] - er); vmax[2] = fmax(vmax[2], e[u] + er);
                vmin[0] = fmin(vmin[0], e[u]); vmin[1] = fmin(vmin[1], e[u] - er); vmin[2] = fmin(vmin[2], e[u] + er);
            }
#pragma unroll
            for (int u = 0; u < kSpecCols; ++u) col[(size_t)(c0 + u) * C] = e[u];
        }
        for (; c0 < ncols; ++c0) {
            const double e = spec_db(__ldcg(col + (size_t)c0 * C) * g2n);
            const double er = (double)cell_err(e, kx, bad);
            vmax[0] = fmax(vmax[0], e); vmax[1] = fmax(vmax[1], e - er); vmax[2] = fmax(vmax[2], e + er);
            vmin[0] = fmin(vmin[0], e); vmin[1] = fmin(vmin[1], e - er); vmin[2] = fmin(vmin[2], e + er);
            col[(size_t)c0 * C] = e;
        }
    }
    // "bad" travels through the reduction as +inf in the interval's upper end
    if (bad) vmax[2] = INFINITY;
    group_maxmin<3, BAR>(vmax, vmin, tid, nthr, s_red, s_out);
    SpecStats st;
    st.mx = vmax[0];
    st.floor_db = st.mx - 80.0;
    st.mn = fmax(vmin[0], st.floor_db);
    const double emx = fmax(st.mx - vmax[1], vmax[2] - st.mx) + 1e-12;       // + the rounding of mx - 80 and of the sums above
    const double lo = fmax(vmin[1], st.floor_db - emx), hi = fmax(vmin[2], st.floor_db + emx);
    const double emn = fmax(st.mn - lo, hi - st.mn) + 1e-12;
    st.emx = (float)emx * 1.0001f;
    st.emn = (float)emn * 1.0001f;
    const double range = st.mx - st.mn;
    st.degenerate = range < 1e-8;
    st.rden = 1.0 / (range + 1e-8);
    // the silent-clip test (create_dataset.py:64-65) is a comparison too; non-finite bounds flag the utterance
    st.near = !(fabs(range - 1e-8) > (double)st.emx + (double)st.emn + a.spec_delta) || !(st.emx < 1e30f);
    return st;
}

// near-tie test, Schmitt triggers and spike words of four consecutive bins (spec_epilogue's inner block)
template <int FNPT, int KK>
__device__ __forceinline__ void spec_bins(const GtArgs &a, const double (&v)[4], const float (&mg)[4], const int j0, const int ch,
                                          uint8_t *row0, unsigned char *smem_raw, unsigned &on, bool &near)
{
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j < a.nbins) {
            const double m = (double)mg[u];
#pragma unroll
            for (int k = 0; k < (KK ? KK : 8); ++k)
                if (KK || k < a.K) near |= !(fabs(v[u] - a.thr[k]) > m) | !(fabs(v[u] - a.lower[k]) > m);
            triggers_step<KK>(a, v[u], on);
            put_spikes<FNPT, KK>(a, on, j, ch, row0, smem_raw);
        }
    }
}

// ---- SPECULATIVE epilogue: the same chain (dB, floor, min-max, zoom, encoder) on the speculative energy plane, arranged
//      for throughput - independent columns in flight, library log10, reciprocal instead of division - plus the near-tie
//      test against the derived bound.  Returns true (per thread) if some comparison of the exact path could come out
//      differently; the utterance is then repeated exactly.  The group is NTHR = C threads, thread <-> channel.
template <int FNPT, int BAR>
__device__ __forceinline__ bool spec_epilogue(const GtArgs &a, int utt, double *plane, int tid, int nthr, float xmax,
                                              double *s_red, double *s_out, unsigned char *smem_raw)
{
    const int ch = tid, C = a.C, ncols = a.ncols;
    const bool live = ch < C;
    const SpecStats st = spec_db_pass<BAR>(a, plane + ch, ch, live, tid, nthr, xmax, s_red, s_out);
    bool near = st.near;
    if (!live) return near;
    const float kx = (float)(a.kappa[ch] * (double)xmax) * a.bound_scale;
    const float rdenf = (float)st.rden * 1.001f, deltaf = (float)a.spec_delta;
    const double floor_db = st.floor_db, mn = st.mn, rden = st.rden;
    const double clamp_sure = floor_db - (double)st.emx;       // a cell whose upper end is below this is at the floor on both paths
    const int T = a.nbins * a.K;
    uint8_t *row0 = a.spikes ? a.spikes + ((size_t)utt * C * a.R + (size_t)ch * a.R) * T : nullptr;
    const double *col = plane + ch;
    unsigned on = 0;
    bool bad = false;
    for (int j0 = 0; j0 < a.nbins; j0 += 4) {
        double v[4];
        float mg[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = min(j0 + u, a.nbins - 1);
            int i0 = j;
            double f = 0.0;
            if (ncols != a.nbins) { i0 = __ldg(a.zoom_i0 + j); f = __ldg(a.zoom_f + j); }
            const int i1 = min(i0 + 1, ncols - 1);
            const double d0 = col[(size_t)i0 * C], d1 = col[(size_t)i1 * C];
            const double x0 = (fmax(d0, floor_db) - mn) * rden;
            const double x1 = (fmax(d1, floor_db) - mn) * rden;
            v[u] = st.degenerate ? 0.0 : fma(x1 - x0, f, x0);
            // bound on |v_exact - v|: the two cells (at the floor for certain: only the floor's own error), the minimum and
            // the maximum enter with weights (1-f, f), (1-v) and v; everything divided by the range
            float e0 = cell_err(d0, kx, bad), e1 = cell_err(d1, kx, bad);
            e0 = (d0 + (double)e0 < clamp_sure) ? st.emx : fmaxf(e0, st.emx);
            e1 = (d1 + (double)e1 < clamp_sure) ? st.emx : fmaxf(e1, st.emx);
            const float ff = (float)f, vf = fminf(fmaxf((float)v[u], 0.0f), 1.0f);
            mg[u] = ((1.0f - ff) * e0 + ff * e1 + (1.0f - vf) * st.emn + vf * st.emx + deltaf) * rdenf + 1e-13f;
        }
        if (a.K == 4) spec_bins<FNPT, 4>(a, v, mg, j0, ch, row0, smem_raw, on, near);      // the reference's four triggers: no per-trigger tests
        else spec_bins<FNPT, 0>(a, v, mg, j0, ch, row0, smem_raw, on, near);
    }
    return near | bad;
}

// ---- SPECULATIVE filter, lane = utterance ("lanes" arrangement).  The lane = channel kernel keeps its per-channel
//      coefficients in registers, and a DFMA with three distinct register operands issues at only ~75 % of the fp64
//      pipe's rate on this GPU (tools/fp64_cascade.cu: 14.4 vs 19.2 T lane-ops/s).  Here a warp filters J channels of
//      32 utterances: the coefficients are the same for every lane, sit in the kernel-parameter constant bank and reach the
//      DFMAs through uniform registers (two register operands each), so the cascade runs at the pipe's full rate.
//      Each lane streams its own utterance with 16-byte loads (a sector is consumed by two of them, L1 keeps the line),
//      no shared memory, no barriers.  Output: raw window energy sums energy[utt][col][ch] (one full 32-byte sector per
//      lane and window for J = 4); the encoder epilogue of the second kernel takes it from there.
constexpr int kLanesJ = 4;

struct EnergyArgs {
    const float *pcm;       // [B][L]
    double *energy;         // [B][ncols][C] raw energy sums of the normalised cascade
    float *xmax;            // [B] max |sample| (the error bound scales with it)
    int B, L, C, nwin, hop, ncols;
    int n_units;            // big units + single-channel units
    int big_groups;         // utterance groups [0, big_groups) are cut into units of kLanesJ channels, the rest into single channels
    double coef[256][6];    // per channel: c1..c4 (numerator zeros / A0), -a1, -a2
};

// one unit of work: J channels starting at ch0 for the 32 utterances of group g
template <int J>
__device__ __forceinline__ void energy_unit(const EnergyArgs &a, const int g, const int ch0)
{
    const int lane = threadIdx.x & 31;
    const int utt = g * 32 + lane;
    const bool valid = utt < a.B;
    const float4 *src = reinterpret_cast<const float4 *>(a.pcm + (size_t)(valid ? utt : a.B - 1) * a.L);
    const int n4_max = a.L / 4 - 1;
    const int hop = a.hop, ncols = a.ncols;
    const int r_old = a.nwin - 2 * hop;
    const int n_used = (ncols - 1) * hop + a.nwin;
    double *dst = a.energy + (size_t)utt * ncols * a.C + ch0;

    // Stage k works on the sample that stage k-1 finished one iteration earlier (software skew, as in the lane = channel
    // kernels): per channel four independent 3-FMA chains instead of one of depth eleven.  Iteration i feeds x[i] to stage 1
    // and completes output i-3; outputs -3..-1 are exact zeros (zero state), so the energy sums need no prologue.
    double xp = 0.0;
    float xm = 0.0f;
    double p1[J], q1[J], p2[J], q2[J], p3[J], q3[J], p4[J], q4[J], acc[J], full1[J], full2[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { p1[j] = q1[j] = p2[j] = q2[j] = p3[j] = q3[j] = p4[j] = q4[j] = acc[j] = full1[j] = full2[j] = 0.0; }

#define LSM_LANE_SAMPLE(xf)                                                                             \
    {                                                                                                   \
        const double x_ = (double)(xf);                                                                 \
        _Pragma("unroll") for (int j = 0; j < J; ++j) {                                                 \
            const double *c = a.coef[ch0 + j];                                                          \
            const double y1 = fma(c[4], p1[j], fma(c[5], q1[j], fma(c[0], xp, x_)));                    \
            const double y2 = fma(c[4], p2[j], fma(c[5], q2[j], fma(c[1], q1[j], p1[j])));              \
            const double y3 = fma(c[4], p3[j], fma(c[5], q3[j], fma(c[2], q2[j], p2[j])));              \
            const double y4 = fma(c[4], p4[j], fma(c[5], q4[j], fma(c[3], q3[j], p3[j])));              \
            q1[j] = p1[j]; p1[j] = y1; q2[j] = p2[j]; p2[j] = y2;                                       \
            q3[j] = p3[j]; p3[j] = y3; q4[j] = p4[j]; p4[j] = y4;                                       \
            acc[j] = fma(y4, y4, acc[j]);                                                               \
        }                                                                                               \
        xp = x_;                                                                                        \
    }

    // Groups of 8 iterations aligned with the 16-byte loads; group gi completes outputs 8gi-3 .. 8gi+4, so a window-phase
    // boundary at sample 8gi (the phases are multiples of 8 samples long) falls after the group's third iteration.
    int next_b = r_old;                                // next boundary: first the head of hop-block 0 ...
    int m = 0;                                         // hop-block the boundary belongs to
    bool head = true;                                  // boundary kind: end of the block's head (window m-2 complete) / end of the block
    const int n_groups = n_used / 8;                   // the last boundary (sample n_used) is handled after the loop's last group
    float4 f0 = __ldg(src + 0), f1 = __ldg(src + 1);   // two loads (8 samples) in flight ahead of the arithmetic
    for (int gi = 0; gi <= n_groups; ++gi) {
        const float4 g0 = __ldg(src + min(2 * gi + 2, n4_max)), g1 = __ldg(src + min(2 * gi + 3, n4_max));
        if (ch0 == 0)      // one unit per group reports the peak level (every sample that can reach a window, and a few more)
            xm = fmaxf(fmaxf(fmaxf(xm, fabsf(f0.x)), fmaxf(fabsf(f0.y), fabsf(f0.z))),
                       fmaxf(fmaxf(fabsf(f0.w), fabsf(f1.x)), fmaxf(fmaxf(fabsf(f1.y), fabsf(f1.z)), fabsf(f1.w))));
        LSM_LANE_SAMPLE(f0.x) LSM_LANE_SAMPLE(f0.y) LSM_LANE_SAMPLE(f0.z)
        if (8 * gi == next_b) {
            if (head) {
                if (m >= 2 && valid) {
                    // window m-2 complete: full(m-2) + full(m-1) + head(m)
                    double *o = dst + (size_t)(m - 2) * a.C;
                    if (J == 4) {
                        *reinterpret_cast<double2 *>(o) = make_double2((full2[0] + full1[0]) + acc[0], (full2[1] + full1[1]) + acc[1]);
                        *reinterpret_cast<double2 *>(o + 2) = make_double2((full2[2] + full1[2]) + acc[2], (full2[3] + full1[3]) + acc[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < J; ++j) o[j] = (full2[j] + full1[j]) + acc[j];
                    }
                }
                next_b += hop - r_old;
                head = false;
            } else {
#pragma unroll
                for (int j = 0; j < J; ++j) { full2[j] = full1[j]; full1[j] = acc[j]; acc[j] = 0.0; }
                next_b += r_old;
                head = true;
                ++m;
            }
        }
        LSM_LANE_SAMPLE(f0.w) LSM_LANE_SAMPLE(f1.x) LSM_LANE_SAMPLE(f1.y) LSM_LANE_SAMPLE(f1.z) LSM_LANE_SAMPLE(f1.w)
        f0 = g0; f1 = g1;
    }
#undef LSM_LANE_SAMPLE
    if (ch0 == 0 && valid) a.xmax[utt] = xm;
}

