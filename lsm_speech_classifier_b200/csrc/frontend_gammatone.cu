// K1 — gammatone filterbank + dB + min-max + zoom + 4-threshold hysteresis encoder, one kernel.
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
//
// Mapping: one CTA per utterance in flight (persistent grid, CTAs stride over the batch), one thread
// per channel.  The IIR recurrences are inherently sequential in the reference's rounding order, so
// parallelism is (utterance x channel): 128 chains per utterance, thousands of utterances.  The
// kernel is bound by the fp64 pipe (~35 DADD/DMUL/DFMA per channel-sample), not by HBM: per
// utterance it reads 64 000 B of PCM and writes 51 200 B of spikes.
//
// Bit-exactness: every fp64 operation is an explicit __d*_rn intrinsic in the oracle's order
// (oracle/lsm_oracle.c gammatone_energy / db_normalise_zoom / hysteresis_encode_f64).
#include <stdlib.h>

#include "reservoir_core.cuh"

namespace {

constexpr int kChunkBlocks = 8;   // hop-blocks of PCM staged per shared-memory buffer

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
    const float *pcm;       // [B][L]
    const double *coefs;    // [C][10]
    const int32_t *zoom_i0; // [nbins]
    const double *zoom_f;   // [nbins]
    double *scratch;        // [grid][ncols][C]
    uint8_t *spikes;        // [B][C*R][nbins*K]
    double *spec_norm;      // optional [B][C][nbins]
    int B, L, C, nwin, hop, ncols, nbins, K, R;
    int mode;               // 0 = exact filter only; 1 = speculative filter, exact re-execution of near-ties
    double spec_delta;      // dB margin of the near-tie test
    int *reruns;            // number of utterances filtered twice (speculative mode)
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

// PCM -> fp64 in shared memory; buffer element i of chunk k holds sample k*chunk + i + kSkew
__device__ __forceinline__ void stage_pcm(double *dst, const float *pcm, int base, int chunk, int L)
{
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) dst[i] = (base + i < L) ? (double)__ldg(pcm + base + i) : 0.0;
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
__device__ __forceinline__ void gt_filter_exact(const GtArgs &a, const float *pcm, double *s_x, double *plane, double &tmax,
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

    stage_pcm(s_x, pcm, kSkew, chunk, a.L);
    if (live) {
        // prologue: iterations s = -3, -2, -1 fill the skewed pipeline (outputs belong to no sample)
#pragma unroll
        for (int s = 0; s < kSkew; ++s) {
            const double x = (double)__ldg(pcm + s);
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
        if (ck + 1 < n_chunks) stage_pcm(s_x + ((ck + 1) & 1) * chunk, pcm, (ck + 1) * chunk + kSkew, chunk, a.L);
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
//      Its dB plane differs from the exact one by rounding noise only (measured <= 6e-11 dB on pathological clips,
//      ~2e-12 dB on speech-like ones).  The encoder epilogue flags every utterance in which some normalised value
//      comes within a.spec_delta dB (default 1e-7) of an encoder threshold, of a hysteresis bound or of the
//      degenerate-clip test, and those utterances are filtered again by gt_filter_exact: the spike trains that leave the
//      kernel are the exact path's, byte for byte, as long as the two planes agree to a third of that margin.
__device__ __forceinline__ void gt_filter_fast(const GtArgs &a, const float *pcm, double *s_x, double *plane, double &tmax,
                                               double &tmin)
{
    const int ch = threadIdx.x;
    const int C = a.C;
    const bool live = ch < C;
    const int hop = a.hop, nwin = a.nwin, ncols = a.ncols;
    const int chunk = kChunkBlocks * hop;
    const int r_old = nwin - 2 * hop;
    const int n_used = (ncols - 1) * hop + nwin;
    const int n_blocks = (n_used + hop - 1) / hop;
    const int n_chunks = (n_blocks + kChunkBlocks - 1) / kChunkBlocks;

    double c1 = 0, c2 = 0, c3 = 0, c4 = 0, na1 = 0, na2 = 0, g2n = 0;
    if (live) {
        const double *c = a.coefs + 10 * ch;
        const double a0 = c[6], A0 = c[0];
        c1 = c[1] / A0; c2 = c[2] / A0; c3 = c[3] / A0; c4 = c[4] / A0;
        na1 = -(c[7] / a0);
        na2 = -(c[8] / a0);
        const double s = A0 / a0;
        const double G = (s * s) * (s * s) / c[9];
        g2n = G * G / (double)nwin;
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

    stage_pcm(s_x, pcm, kSkew, chunk, a.L);
    if (live) {
#pragma unroll
        for (int s = 0; s < kSkew; ++s) LSM_FAST_SAMPLE((double)__ldg(pcm + s));
        acc = 0.0;   // (already zero: stage 4 has seen no sample yet)
    }
    __syncthreads();

    for (int ck = 0; ck < n_chunks; ++ck) {
        const double *xs = s_x + (ck & 1) * chunk;
        if (ck + 1 < n_chunks) stage_pcm(s_x + ((ck + 1) & 1) * chunk, pcm, (ck + 1) * chunk + kSkew, chunk, a.L);
        if (live) {
            for (int bl = 0; bl < kChunkBlocks; ++bl) {
                const int m = ck * kChunkBlocks + bl;
                if (m >= n_blocks) break;
                const double *xb = xs + bl * hop;
                const int n_here = min(hop, n_used - m * hop);
                const int n_a = min(n_here, r_old);
                acc = 0.0;
#pragma unroll 8
                for (int p = 0; p < n_a; ++p) LSM_FAST_SAMPLE(xb[p]);
                if (n_a == r_old && m >= 2) emit_db(sqrt(((full2 + full1) + acc) * g2n), plane, m - 2, C, ch, tmax, tmin);
#pragma unroll 8
                for (int p = n_a; p < n_here; ++p) LSM_FAST_SAMPLE(xb[p]);
                full2 = full1;
                full1 = acc;
            }
        }
        __syncthreads();
    }
#undef LSM_FAST_SAMPLE
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
    __shared__ double s_red[2][8];
    __shared__ double s_mm[2];
    __shared__ int s_utt;
    __shared__ int s_cnt3[3];

    const int ch = threadIdx.x;
    const int C = a.C;
    const bool live = ch < C;
    const int ncols = a.ncols;
    double *plane = a.scratch + (size_t)blockIdx.x * ncols * C;     // this CTA's dB plane [ncols][C]

    for (;;) {
        // dynamic work distribution: utterances are handed out one at a time, so every SM stays busy to the end
        if (threadIdx.x == 0) s_utt = atomicAdd(next_utt, 1);
        __syncthreads();
        const int utt = s_utt;
        if (utt >= a.B) break;
        const float *pcm = a.pcm + (size_t)utt * a.L;
        bool exact = a.mode == 0;
        for (;;) {
            double tmax = -INFINITY, tmin = INFINITY;
            if (exact) gt_filter_exact(a, pcm, s_x, plane, tmax, tmin);
            else gt_filter_fast(a, pcm, s_x, plane, tmax, tmin);

            // ---- per-utterance max / min of the dB plane (create_dataset.py:60,62-63)
            {
                const double wmax = warp_max_f64(tmax), wmin = warp_min_f64(tmin);
                if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = wmax; s_red[1][threadIdx.x >> 5] = wmin; }
                __syncthreads();
                if (threadIdx.x == 0) {
                    double mx = -INFINITY, mn = INFINITY;
                    for (int w = 0; w < (blockDim.x >> 5); ++w) { mx = fmax(mx, s_red[0][w]); mn = fmin(mn, s_red[1][w]); }
                    s_mm[0] = mx; s_mm[1] = mn;
                }
                __syncthreads();
            }
            const double mx = s_mm[0];
            const double floor_db = sub64(mx, 80.0);
            const double mn = fmax(s_mm[1], floor_db);        // min of the clamped plane
            const double range = sub64(mx, mn);
            const bool degenerate = range < 1e-8;              // create_dataset.py:64-65 -> all zeros
            const double den = add64(range, 1e-8);
            // speculative pass: margin (in normalised units) inside which a comparison could come out differently
            const double margin = exact ? 0.0 : a.spec_delta / den;
            bool near = !exact && fabs(range - 1e-8) < a.spec_delta;

            if (live) {
                const int T = a.nbins * a.K;
                uint8_t *row0 = a.spikes ? a.spikes + ((size_t)utt * C * a.R + (size_t)ch * a.R) * T : nullptr;
                double *dump = a.spec_norm ? a.spec_norm + ((size_t)utt * C + ch) * a.nbins : nullptr;
                // normalise in place (own column of the plane only)
                for (int c = 0; c < ncols; ++c) {
                    const double v = fmax(plane[(size_t)c * C + ch], floor_db);
                    plane[(size_t)c * C + ch] = __ddiv_rn(sub64(v, mn), den);
                }
                unsigned on = 0;   // bit k = state of trigger k
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
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (k < a.K) {
                            const bool is_on = (on >> k) & 1u;
                            if (!is_on && v > a.thr[k]) on |= (1u << k);
                            else if (is_on && v < a.lower[k]) on &= ~(1u << k);
                            if (!exact) near |= (fabs(v - a.thr[k]) < margin) | (fabs(v - a.lower[k]) < margin);
                        }
                    }
                    if (FNPT > 0) {
                        // hand the spikes to the reservoir phase: word (t, warp) = ballot over this warp's 32 channels
                        unsigned *s_bits = reinterpret_cast<unsigned *>(smem_raw);
                        const int CW = C >> 5;
                        for (int k = 0; k < a.K; ++k) {
                            const unsigned word = __ballot_sync(0xffffffffu, (on >> k) & 1u);
                            if ((threadIdx.x & 31) == 0) s_bits[(j * a.K + k) * CW + (threadIdx.x >> 5)] = word;
                        }
                    }
                    for (int r = 0; row0 && r < a.R; ++r) {
                        uint8_t *row = row0 + (size_t)r * T + (size_t)j * a.K;
                        if (a.K == 4) {
                            // bytes k = 0..3 of column block j, little endian
                            const unsigned w = (on & 1u) | ((on & 2u) << 7) | ((on & 4u) << 14) | ((on & 8u) << 21);
                            *reinterpret_cast<uint32_t *>(row) = w;
                        } else {
                            for (int k = 0; k < a.K; ++k) row[k] = (on >> k) & 1u;
                        }
                    }
                }
            }
            if (exact) break;
            // some comparison of this utterance is too close to call on the speculative plane: filter it again, exactly
            if (!__syncthreads_or(near ? 1 : 0)) break;
            if (threadIdx.x == 0) atomicAdd(a.reruns, 1);
            exact = true;
        }
        if (FNPT > 0) {
            __syncthreads();   // bits complete and visible
            reservoir_simulate<(FNPT > 0 ? FNPT : 4), LEAN>(a.res, utt, smem_raw, s_cnt3);
        }
        __syncthreads();   // plane, shared memory and s_utt are reused by the next utterance
    }
}

}  // namespace

// occupancy target: CTAs per SM the register allocation is bounded for (4, 5 or 6 at <= 128 channels)
static int k1_minb()
{
    const char *e = getenv("LSM_K1_MINB");
    const int v = e ? atoi(e) : 5;
    return v < 4 ? 4 : (v > 6 ? 6 : v);
}

template <int MAXT, int MINB, int FNPT, bool LEAN>
static int k1_grid(lsm_ctx *ctx, int threads, size_t smem, int *per_sm)
{
    if (smem > 48 * 1024)
        LSM_CUDA(ctx, cudaFuncSetAttribute(gammatone_encode_kernel<MAXT, MINB, FNPT, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, gammatone_encode_kernel<MAXT, MINB, FNPT, LEAN>, threads, smem));
    return LSM_OK;
}

static size_t k1_pad_smem()
{
    const char *e = getenv("LSM_K1_PAD_SMEM");      // experiment knob: extra dynamic smem to force lower occupancy
    return e ? (size_t)atoi(e) : 0;
}

int lsm_gammatone_grid(lsm_ctx *ctx, const lsm_frontend_params *p, int *grid)
{
    const int threads = ((p->channels + 31) / 32) * 32;
    const size_t smem = sizeof(double) * 2 * kChunkBlocks * p->hop + k1_pad_smem();
    int per_sm = 0, rc;
    if (threads > 128) rc = k1_grid<256, 2, 0, true>(ctx, threads, smem, &per_sm);
    else if (k1_minb() == 4) rc = k1_grid<128, 4, 0, true>(ctx, threads, smem, &per_sm);
    else if (k1_minb() == 5) rc = k1_grid<128, 5, 0, true>(ctx, threads, smem, &per_sm);
    else rc = k1_grid<128, 6, 0, true>(ctx, threads, smem, &per_sm);
    if (rc != LSM_OK) return rc;
    if (per_sm < 1) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "gammatone kernel does not fit on an SM (hop %d)", p->hop);
    *grid = per_sm * ctx->sm_count;   // persistent: every CTA resident, utterances handed out dynamically
    return LSM_OK;
}

static void fill_args(const lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes, double *d_spec_norm, GtArgs *out)
{
    const lsm_frontend_params &p = fe->p;
    GtArgs &a = *out;
    a.pcm = d_pcm; a.coefs = fe->d_coefs; a.zoom_i0 = fe->d_zoom_i0; a.zoom_f = fe->d_zoom_f;
    a.scratch = fe->d_scratch; a.spikes = d_spikes; a.spec_norm = d_spec_norm;
    a.B = B; a.L = p.n_samples; a.C = p.channels; a.nwin = p.nwin; a.hop = p.hop; a.ncols = fe->ncols;
    a.nbins = p.n_bins; a.K = p.n_thresholds; a.R = p.redundancy;
    for (int k = 0; k < 8; ++k) { a.thr[k] = p.thresholds_desc[k]; a.lower[k] = p.lower_bounds[k]; }
    // the normalised-spectrogram dump is defined as the exact path's: asking for it selects the exact filter
    a.mode = d_spec_norm ? 0 : fe->mode;
    a.spec_delta = fe->spec_delta;
    a.reruns = fe->d_counters + 64;
}

// The per-CTA dB planes (grid x ~100 KB) are written and re-read by the same CTA for every utterance.  Mark that
// region L2-persisting on the launch stream so it is not evicted to HBM by the PCM / spike streams passing through.
static void pin_scratch_in_l2(lsm_ctx *ctx, lsm_frontend *fe, cudaStream_t st)
{
    if (getenv("LSM_NO_L2_PIN")) return;
    const size_t bytes = sizeof(double) * (size_t)fe->grid * fe->ncols * fe->p.channels;
    if (!fe->l2_window_ready) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
        fe->l2_window_bytes = 0;
        if (max_persist > 0 && max_window > 0) {
            size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
                fe->l2_window_bytes = bytes < (size_t)max_window ? bytes : (size_t)max_window;
            fe->l2_hit_ratio = bytes <= want ? 1.0f : (float)want / (float)bytes;
        }
        cudaGetLastError();
        fe->l2_window_ready = 1;
    }
    if (!fe->l2_window_bytes) return;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = fe->d_scratch;
    v.accessPolicyWindow.num_bytes = fe->l2_window_bytes;
    v.accessPolicyWindow.hitRatio = fe->l2_hit_ratio;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
}

static int next_counter(lsm_ctx *ctx, lsm_frontend *fe, cudaStream_t st, int **counter)
{
    int rc = lsm_frontend_order_before(ctx, fe, st);
    if (rc != LSM_OK) return rc;
    pin_scratch_in_l2(ctx, fe, st);
    // work counter: one int per launch out of a small ring, so back-to-back launches on different streams do not share it
    *counter = fe->d_counters + (fe->counter_next++ % 64);
    LSM_CUDA(ctx, cudaMemsetAsync(*counter, 0, sizeof(int), st));
    return LSM_OK;
}

int lsm_launch_gammatone(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes,
                         double *d_spec_norm, cudaStream_t st)
{
    const lsm_frontend_params &p = fe->p;
    GtArgs a;
    fill_args(fe, d_pcm, B, d_spikes, d_spec_norm, &a);
    memset(&a.res, 0, sizeof(a.res));
    const int threads = ((p.channels + 31) / 32) * 32;
    const size_t smem = sizeof(double) * 2 * kChunkBlocks * p.hop + k1_pad_smem();
    const int grid = B < fe->grid ? B : fe->grid;
    if (grid <= 0) return LSM_OK;
    int *counter, rc;
    if ((rc = next_counter(ctx, fe, st, &counter)) != LSM_OK) return rc;
    if (threads > 128) gammatone_encode_kernel<256, 2, 0, true><<<grid, threads, smem, st>>>(a, counter);
    else if (fe->minb == 4) gammatone_encode_kernel<128, 4, 0, true><<<grid, threads, smem, st>>>(a, counter);
    else if (fe->minb == 5) gammatone_encode_kernel<128, 5, 0, true><<<grid, threads, smem, st>>>(a, counter);
    else gammatone_encode_kernel<128, 6, 0, true><<<grid, threads, smem, st>>>(a, counter);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st);
}

// Can this (front end, reservoir) pair run as one fused kernel?  Gammatone, no redundancy, one thread
// per channel with whole warps, and a reservoir whose padded width is channels x {4, 8, 16} neurons per thread.
int lsm_fused_npt(const lsm_frontend *fe, const lsm_reservoir *res)
{
    const lsm_frontend_params &p = fe->p;
    if (getenv("LSM_NO_FUSE")) return 0;
    if (p.kind != LSM_FILTERBANK_GAMMATONE || p.redundancy != 1 || (p.channels & 31) || p.channels > 256) return 0;
    if (res->p.num_inputs != p.channels || res->p.num_steps != p.n_bins * p.n_thresholds) return 0;
    if (res->n_pad % p.channels) return 0;
    const int npt = res->n_pad / p.channels;
    if (p.channels == 256) return npt == 4 ? 4 : 0;
    return (npt == 8 || npt == 16) ? npt : 0;
}

// launch == false: only report the resident grid (one wave) of this variant through *wave
template <int MAXT, int MINB, int FNPT>
static int launch_fused_t(lsm_ctx *ctx, lsm_frontend *fe, const lsm_reservoir *res, const GtArgs &a, int threads, size_t smem,
                          cudaStream_t st, bool launch = true, int *wave = nullptr)
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
    int *counter;
    if ((rc = next_counter(ctx, fe, st, &counter)) != LSM_OK) return rc;
    if (res->lean) gammatone_encode_kernel<MAXT, MINB, FNPT, true><<<grid, threads, smem, st>>>(a, counter);
    else gammatone_encode_kernel<MAXT, MINB, FNPT, false><<<grid, threads, smem, st>>>(a, counter);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st);
}

static int fused_dispatch(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B,
                          uint8_t *d_spikes_or_null, uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st,
                          bool launch, int *wave);

int lsm_launch_fused(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B,
                     uint8_t *d_spikes_or_null, uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st)
{
    if (B <= 0) return LSM_OK;
    return fused_dispatch(ctx, fe, res, d_pcm, B, d_spikes_or_null, feature_mask, nan_to_num, d_features, st, true, nullptr);
}

// CTAs resident at once for the fused kernel of this pair = utterances per wave (0 if not fusable)
int lsm_fused_wave(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res)
{
    int wave = 0;
    if (!lsm_fused_npt(fe, res)) return 0;
    if (fused_dispatch(ctx, fe, res, nullptr, 1, nullptr, 1u, 0, nullptr, nullptr, false, &wave) != LSM_OK) return 0;
    return wave;
}

static int fused_dispatch(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B,
                          uint8_t *d_spikes_or_null, uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st,
                          bool launch, int *wave)
{
    const int npt = lsm_fused_npt(fe, res);
    if (!npt) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "this front end / reservoir pair cannot run fused");
    const lsm_frontend_params &p = fe->p;
    GtArgs a;
    fill_args(fe, d_pcm, B, d_spikes_or_null, nullptr, &a);
    lsm_reservoir_fill_args(res, nullptr, B, feature_mask, nan_to_num, d_features, nullptr, &a.res);
    const int threads = p.channels;
    size_t smem = sizeof(double) * 2 * kChunkBlocks * p.hop;
    const size_t smem_res = lsm_res_smem_bytes(a.res.T, a.res.CW, threads * npt, a.res.N);
    if (smem_res > smem) smem = smem_res;
    if (threads == 256) return launch_fused_t<256, 2, 4>(ctx, fe, res, a, threads, smem, st, launch, wave);
    if (npt == 8) {
        if (fe->minb >= 5) return launch_fused_t<128, 5, 8>(ctx, fe, res, a, threads, smem, st, launch, wave);
        return launch_fused_t<128, 4, 8>(ctx, fe, res, a, threads, smem, st, launch, wave);
    }
    return launch_fused_t<128, 4, 16>(ctx, fe, res, a, threads, smem, st, launch, wave);
}

int lsm_gammatone_minb(void) { return k1_minb(); }
