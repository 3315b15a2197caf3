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
// Kernels in this file:
//   gammatone_encode_kernel<MAXT, MINB, FNPT, LEAN>  persistent, one CTA per utterance in flight, one thread per channel.  FNPT = 0:
//       front end only; FNPT > 0: fused audio -> features (the CTA then simulates the utterance's reservoir, reservoir_core.cuh).
//       Filter modes: exact (gt_filter_exact: every fp64 operation an explicit __d*_rn intrinsic in the oracle's order -
//       oracle/lsm_oracle.c gammatone_energy / db_normalise_zoom / hysteresis_encode_f64 - 35 operations per channel-sample) and,
//       by default, speculative (gt_filter_fast: the same cascade in 13 FMAs; spec_epilogue flags near-ties, which are filtered
//       again exactly inside the kernel, so the spike trains are the exact path's either way).
//   gammatone_energy_kernel + encode_reservoir_kernel / gammatone_encode_kernel in mode 2: the lanes arrangement (lane =
//       utterance, coefficients in uniform registers); default for the stand-alone front end, opt-in for the whole path.
//   spec_fused_kernel: speculative-only variant with the exact pass as a follow-up launch (opt-in, measured slower).
// The filter is bound by the fp64 pipe, not by HBM: per utterance it reads 64 000 B of PCM and writes 51 200 B of spikes.
#include <stdlib.h>

#include <memory>
#include <new>

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
    const float *pcm;       // [B][L] float32 samples, or null when pcm16 is given
    const int16_t *pcm16;   // [B][L] PCM16 samples (optional alternative input)
    const double *coefs;    // [C][10]
    const int32_t *zoom_i0; // [nbins]
    const double *zoom_f;   // [nbins]
    double *scratch;        // [grid][ncols][C]
    uint8_t *spikes;        // [B][C*R][nbins*K]
    double *spec_norm;      // optional [B][C][nbins]
    int B, L, C, nwin, hop, ncols, nbins, K, R;
    const double *energy_in; // mode 2: [B][ncols][C] raw energy sums from gammatone_energy_kernel (also this utterance's plane)
    int mode;               // 0 = exact filter only; 1 = speculative filter (lane = channel) in this kernel, exact re-execution
                            // of near-ties; 2 = as 1, but the speculative energies were computed by gammatone_energy_kernel
    double spec_delta;      // dB margin of the near-tie test
    int *reruns;            // number of utterances filtered twice (speculative mode)
    const int *utt_list;    // optional indirection: work item i is utterance utt_list[i], i < *utt_count (exact re-execution pass)
    const int *utt_count;
    int *rerun_list;        // encode_reservoir_kernel: utterances whose speculative plane was too close to call; [0] = count
    int *sm_rank;           // [256] per-SM arrival counter of this launch (staggered start)
    int stagger_cycles;     // start delay per CTA rank on its SM, 0 = none
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

// PCM -> fp64 in shared memory; buffer element i of chunk k holds sample k*chunk + i + kSkew
__device__ __forceinline__ void stage_pcm(double *dst, const PcmRow pcm, int base, int chunk, int L)
{
    // streaming loads (evict-first): every PCM sample is read once and must not push the CTAs' scratch planes out of L2
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) dst[i] = (base + i < L) ? pcm.at(base + i) : 0.0;
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
__device__ __forceinline__ void gt_filter_fast(const GtArgs &a, const PcmRow pcm, double *s_x, double *plane)
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

    stage_pcm(s_x, pcm, kSkew, chunk, a.L);
    if (live) {
#pragma unroll
        for (int s = 0; s < kSkew; ++s) LSM_FAST_SAMPLE(pcm.at(s));
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
                // window m-2 complete: its raw energy sum; dB is taken in the epilogue, where the columns give ILP
                if (n_a == r_old && m >= 2) plane[(size_t)(m - 2) * C + ch] = (full2 + full1) + acc;
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

// ---- SPECULATIVE filter, lane = utterance ("lanes" arrangement).  The lane = channel kernel above keeps its per-channel
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
    int B, L, C, nwin, hop, ncols;
    int n_units;            // big units + single-channel units
    int big_groups;         // utterance groups [0, big_groups) are cut into units of kLanesJ channels, the rest into single channels
    double coef[256][6];    // per channel: c1..c4 (numerator zeros / A0), -a1, -a2
};

// one unit of work: J channels starting at ch0 for the 32 utterances of group g
template <int J>
__device__ __forceinline__ void energy_unit(const EnergyArgs &a, const int g, const int ch0)
{
    const int lane = threadIdx.x;
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
}

// Units are dispatched in blockIdx order: first the big ones (kLanesJ channels: one PCM stream and one conversion feed
// 52 DFMAs), then single-channel units.  A unit is one warp's serial work (0.8 ms for four channels even on an idle SM),
// so a grid of big units alone ends in a long ragged tail; the small units keep every SM's pipes full to the end.
// The unit index is blockIdx.x + k * gridDim.x: with a grid as large as the unit count every CTA runs one unit (k = 0), with a
// smaller grid (LSM_K1A_PER_SM resident warps per SM, the co-residency experiment) a CTA walks a static list - either way the
// index is provably CTA-uniform for the compiler, which is what puts the coefficients in uniform registers.
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

// per-utterance max / min over the CTA (create_dataset.py:60,62-63); every thread gets both
__device__ __forceinline__ void block_minmax(double tmax, double tmin, double (*s_red)[8], double *s_mm, double &mx, double &mn)
{
    const double wmax = warp_max_f64(tmax), wmin = warp_min_f64(tmin);
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = wmax; s_red[1][threadIdx.x >> 5] = wmin; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double m1 = -INFINITY, m0 = INFINITY;
        for (int w = 0; w < (blockDim.x >> 5); ++w) { m1 = fmax(m1, s_red[0][w]); m0 = fmin(m0, s_red[1][w]); }
        s_mm[0] = m1; s_mm[1] = m0;
    }
    __syncthreads();
    mx = s_mm[0];
    mn = s_mm[1];
}

// the four (K) Schmitt triggers of one channel for one time bin (create_dataset.py:90-94), bit k of `on` = trigger k
__device__ __forceinline__ void triggers_step(const GtArgs &a, double v, unsigned &on)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < a.K) {
            const bool is_on = (on >> k) & 1u;
            if (!is_on && v > a.thr[k]) on |= (1u << k);
            else if (is_on && v < a.lower[k]) on &= ~(1u << k);
        }
    }
}

// spikes of time bin j: to the reservoir's bit plane in shared memory (fused kernels) and / or to X_spikes rows
template <int FNPT>
__device__ __forceinline__ void put_spikes(const GtArgs &a, unsigned on, int j, uint8_t *row0, unsigned char *smem_raw)
{
    if (FNPT > 0) {
        // word (t, warp) = ballot over this warp's 32 channels
        unsigned *s_bits = reinterpret_cast<unsigned *>(smem_raw);
        const int CW = a.C >> 5;
        for (int k = 0; k < a.K; ++k) {
            const unsigned word = __ballot_sync(0xffffffffu, (on >> k) & 1u);
            if ((threadIdx.x & 31) == 0) s_bits[(j * a.K + k) * CW + (threadIdx.x >> 5)] = word;
        }
    }
    const int T = a.nbins * a.K;
    for (int r = 0; row0 && r < a.R; ++r) {
        uint8_t *row = row0 + (size_t)r * T + (size_t)j * a.K;
        if (a.K == 4) {
            // bytes k = 0..3 of column block j, little endian
            const unsigned w = (on & 1u) | ((on & 2u) << 7) | ((on & 4u) << 14) | ((on & 8u) << 21);
            __stcs(reinterpret_cast<unsigned *>(row), w);        // streaming store: written once, never read here
        } else {
            for (int k = 0; k < a.K; ++k) row[k] = (on >> k) & 1u;
        }
    }
}

// ---- EXACT epilogue: floor, min-max, zoom, encoder in the reference's operations (create_dataset.py:60-98)
template <int FNPT>
__device__ __forceinline__ void exact_epilogue(const GtArgs &a, int utt, double *plane, double tmax, double tmin,
                                               double (*s_red)[8], double *s_mm, unsigned char *smem_raw)
{
    const int ch = threadIdx.x, C = a.C, ncols = a.ncols;
    double mx, mn0;
    block_minmax(tmax, tmin, s_red, s_mm, mx, mn0);
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
        put_spikes<FNPT>(a, on, j, row0, smem_raw);
    }
}

// ---- SPECULATIVE epilogue: the same chain (dB, floor, min-max, zoom, encoder) on the speculative energy plane, arranged
//      for throughput - independent columns in flight, library log10, reciprocal instead of division - plus the near-tie
//      test.  Returns true (per thread) if some comparison was within the margin; the CTA then repeats the utterance exactly.
template <int FNPT>
__device__ __forceinline__ bool spec_epilogue(const GtArgs &a, int utt, double *plane, double (*s_red)[8], double *s_mm,
                                              unsigned char *smem_raw)
{
    const int ch = threadIdx.x, C = a.C, ncols = a.ncols;
    const bool live = ch < C;
    double tmax = -INFINITY, tmin = INFINITY;
    if (live) {
        const double *c = a.coefs + 10 * ch;
        const double s = c[0] / c[6];
        const double G = (s * s) * (s * s) / c[9];
        const double g2n = G * G / (double)a.nwin;      // (A0^4 / gain)^2 / nwin: energy sum -> mean square of the real output
        double *col = plane + ch;
        int c0 = 0;
        for (; c0 + 7 <= ncols; c0 += 7) {
            double e[7];
#pragma unroll
            for (int u = 0; u < 7; ++u) e[u] = col[(size_t)(c0 + u) * C];
#pragma unroll
            for (int u = 0; u < 7; ++u) {
                e[u] = 20.0 * log10(sqrt(e[u] * g2n) + 1e-9);
                tmax = fmax(tmax, e[u]);
                tmin = fmin(tmin, e[u]);
            }
#pragma unroll
            for (int u = 0; u < 7; ++u) col[(size_t)(c0 + u) * C] = e[u];
        }
        for (; c0 < ncols; ++c0) {
            const double e = 20.0 * log10(sqrt(col[(size_t)c0 * C] * g2n) + 1e-9);
            tmax = fmax(tmax, e);
            tmin = fmin(tmin, e);
            col[(size_t)c0 * C] = e;
        }
    }
    double mx, mn0;
    block_minmax(tmax, tmin, s_red, s_mm, mx, mn0);
    const double floor_db = mx - 80.0;
    const double mn = fmax(mn0, floor_db);
    const double range = mx - mn;
    const bool degenerate = range < 1e-8;
    const double den = range + 1e-8;
    const double rden = 1.0 / den;
    // margin in normalised units inside which one of the exact path's comparisons could come out differently
    const double margin = a.spec_delta * rden;
    bool near = fabs(range - 1e-8) < a.spec_delta;
    if (!live) return near;
    const int T = a.nbins * a.K;
    uint8_t *row0 = a.spikes ? a.spikes + ((size_t)utt * C * a.R + (size_t)ch * a.R) * T : nullptr;
    const double *col = plane + ch;
    // diagnostic only (LSM_SPEC_DUMP): the speculative normalised spectrogram, to measure its distance from the exact one
    double *dump = a.spec_norm ? a.spec_norm + ((size_t)utt * C + ch) * a.nbins : nullptr;
    unsigned on = 0;
    for (int j0 = 0; j0 < a.nbins; j0 += 4) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = min(j0 + u, a.nbins - 1);
            int i0 = j;
            double f = 0.0;
            if (ncols != a.nbins) { i0 = __ldg(a.zoom_i0 + j); f = __ldg(a.zoom_f + j); }
            const int i1 = min(i0 + 1, ncols - 1);
            const double x0 = (fmax(col[(size_t)i0 * C], floor_db) - mn) * rden;
            const double x1 = (fmax(col[(size_t)i1 * C], floor_db) - mn) * rden;
            v[u] = degenerate ? 0.0 : fma(x1 - x0, f, x0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            if (j < a.nbins) {
                if (dump) dump[j] = v[u];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k < a.K) near |= (fabs(v[u] - a.thr[k]) < margin) | (fabs(v[u] - a.lower[k]) < margin);
                triggers_step(a, v[u], on);
                put_spikes<FNPT>(a, on, j, row0, smem_raw);
            }
        }
    }
    return near;
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
    __shared__ int s_cnt3[5];

    // Every utterance costs the same, so the CTAs resident on an SM would march in lockstep: all in the filter loop
    // (fp64 pipe saturated, issue slots to spare), then all in the encoder / reservoir phases (fp64 pipe idle).  Start
    // the r-th CTA of each SM r x (utterance period / CTAs per SM) late instead, once; the phases then interleave for the
    // rest of the launch and the non-fp64 work hides under the other CTAs' filter loops.
    if (a.stagger_cycles > 0 && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const int r = atomicAdd(a.sm_rank + (smid & 255u), 1);
        const long long t0 = clock64();
        const long long wait = (long long)r * a.stagger_cycles;
        while (clock64() - t0 < wait) __nanosleep(2000);
    }

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
            if (a.mode == 1) gt_filter_fast(a, pcm, s_x, plane);
            const bool near = spec_epilogue<FNPT>(a, utt, plane, s_red, s_mm, smem_raw);
            // some comparison of this utterance is too close to call on the speculative plane: filter it again, exactly
            settled = !__syncthreads_or(near ? 1 : 0);
            if (!settled && threadIdx.x == 0) atomicAdd(a.reruns, 1);
        }
        if (!settled) {
            double tmax = -INFINITY, tmin = INFINITY;
            gt_filter_exact(a, pcm, s_x, plane, tmax, tmin);
            exact_epilogue<FNPT>(a, utt, plane, tmax, tmin, s_red, s_mm, smem_raw);
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

namespace {
// Speculative-only form of the fused kernel (default shape): fast filter, speculative epilogue, reservoir, readout - and nothing
// of the exact path, which keeps the kernel's code and register footprint down.  Flagged utterances are completed like the
// others and appended to the device work list; the host follows with gammatone_encode_kernel in exact mode on that list.
template <int MINB, int FNPT, bool LEAN>
__global__ void __launch_bounds__(128, MINB) spec_fused_kernel(const GtArgs a, int *next_utt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_x = reinterpret_cast<double *>(smem_raw);
    __shared__ double s_red[2][8];
    __shared__ double s_mm[2];
    __shared__ int s_utt;
    __shared__ int s_cnt3[5];
    double *plane = a.scratch + (size_t)blockIdx.x * a.ncols * a.C;
    for (;;) {
        if (threadIdx.x == 0) s_utt = atomicAdd(next_utt, 1);
        __syncthreads();
        const int utt = s_utt;
        if (utt >= a.B) break;
        const PcmRow pcm = {a.pcm ? a.pcm + (size_t)utt * a.L : nullptr, a.pcm16 ? a.pcm16 + (size_t)utt * a.L : nullptr};
        gt_filter_fast(a, pcm, s_x, plane);
        const bool near = spec_epilogue<FNPT>(a, utt, plane, s_red, s_mm, smem_raw);
        if (__syncthreads_or(near ? 1 : 0) && threadIdx.x == 0) {
            a.rerun_list[1 + atomicAdd(a.rerun_list, 1)] = utt;
            atomicAdd(a.reruns, 1);
        }
        if (FNPT > 0) {
            __syncthreads();
            reservoir_simulate<(FNPT > 0 ? FNPT : 4), LEAN>(a.res, utt, smem_raw, s_cnt3);
        }
        __syncthreads();
    }
}
}  // namespace

namespace {
// Second kernel of the lanes arrangement: encoder epilogue on the energy planes of gammatone_energy_kernel, then the
// utterance's reservoir and feature readout.  No filter code, so it runs at a higher occupancy than the fused kernel above
// (the reservoir phase is latency / issue bound).  Utterances whose speculative plane is too close to call are finished
// like the others and also appended to a list; the host follows up with gammatone_encode_kernel in exact mode on that list
// (typically 0-1 utterances per 2400), which overwrites their spike trains and feature rows.
template <int MINB, int FNPT, bool LEAN>
__global__ void __launch_bounds__(128, MINB) encode_reservoir_kernel(const GtArgs a, int *next_utt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[2][8];
    __shared__ double s_mm[2];
    __shared__ int s_utt;
    __shared__ int s_cnt3[5];
    for (;;) {
        if (threadIdx.x == 0) s_utt = atomicAdd(next_utt, 1);
        __syncthreads();
        const int utt = s_utt;
        if (utt >= a.B) break;
        double *plane = const_cast<double *>(a.energy_in) + (size_t)utt * a.ncols * a.C;
        const bool near = spec_epilogue<FNPT>(a, utt, plane, s_red, s_mm, smem_raw);
        if (__syncthreads_or(near ? 1 : 0) && threadIdx.x == 0) {
            a.rerun_list[1 + atomicAdd(a.rerun_list, 1)] = utt;
            atomicAdd(a.reruns, 1);
        }
        __syncthreads();   // bits complete and visible
        reservoir_simulate<FNPT, LEAN>(a.res, utt, smem_raw, s_cnt3);
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
    // persistent: every CTA resident, utterances handed out dynamically.  The scratch planes are sized by this grid; the fused
    // 128-channel kernel runs 6 CTAs per SM by default (below), so leave room for that.
    if (threads <= 128 && per_sm < 6) per_sm = 6;
    *grid = per_sm * ctx->sm_count;
    return LSM_OK;
}

static void fill_args(const lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes, double *d_spec_norm, GtArgs *out)
{
    const lsm_frontend_params &p = fe->p;
    GtArgs &a = *out;
    a.pcm = d_pcm; a.pcm16 = d_pcm ? nullptr : fe->next_pcm16; a.coefs = fe->d_coefs; a.zoom_i0 = fe->d_zoom_i0; a.zoom_f = fe->d_zoom_f;
    a.scratch = fe->d_scratch; a.spikes = d_spikes; a.spec_norm = d_spec_norm;
    a.B = B; a.L = p.n_samples; a.C = p.channels; a.nwin = p.nwin; a.hop = p.hop; a.ncols = fe->ncols;
    a.nbins = p.n_bins; a.K = p.n_thresholds; a.R = p.redundancy;
    for (int k = 0; k < 8; ++k) { a.thr[k] = p.thresholds_desc[k]; a.lower[k] = p.lower_bounds[k]; }
    // the normalised-spectrogram dump is defined as the exact path's: asking for it selects the exact filter
    a.mode = (d_spec_norm && !getenv("LSM_SPEC_DUMP")) ? 0 : fe->mode;
    a.energy_in = nullptr;
    a.utt_list = nullptr; a.utt_count = nullptr; a.rerun_list = nullptr;
    a.spec_delta = fe->spec_delta;
    a.reruns = fe->d_counters + 64;
}

// The per-CTA dB planes (grid x ~100 KB) are written and re-read by the same CTA for every utterance.  Mark that
// region L2-persisting on the launch stream so it is not evicted to HBM by the PCM / spike streams passing through.
static void pin_scratch_in_l2(lsm_ctx *ctx, lsm_frontend *fe, cudaStream_t st)
{
    // Off by default since the streaming (evict-first) hints on the PCM loads and the spike / feature stores: with two scratch
    // slots the window (118 MB) no longer fits the persisting carve-out and made things worse (ncu, 2400 utterances per launch:
    // 775 MB of DRAM traffic with the window, 472 MB without; algorithmic 315 MB).  LSM_L2_PIN=1 turns it back on.
    if (!getenv("LSM_L2_PIN")) return;
    const size_t bytes = 2 * sizeof(double) * (size_t)fe->grid * fe->ncols * fe->p.channels;   // both slots
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

// d_counters layout (ints): [0,64) work counters, [64] re-execution count, [128 + 256*slot, +256) per-SM arrival counters
static int next_counter(lsm_ctx *ctx, lsm_frontend *fe, cudaStream_t st, int **counter, GtArgs *a, int grid, int slot)
{
    int rc = lsm_frontend_order_before(ctx, fe, st, slot);
    if (rc != LSM_OK) return rc;
    a->scratch = fe->d_scratch + (slot > 0 ? (size_t)fe->grid * fe->ncols * fe->p.channels : 0);
    pin_scratch_in_l2(ctx, fe, st);
    // work counter: one int per launch out of a small ring, so back-to-back launches on different streams do not share it
    const unsigned cslot = fe->counter_next++ % 64;
    *counter = fe->d_counters + cslot;
    LSM_CUDA(ctx, cudaMemsetAsync(*counter, 0, sizeof(int), st));
    // staggered start (see the kernel): only when every CTA gets at least two utterances, so the one-off delay pays
    a->sm_rank = fe->d_counters + 128 + 256 * cslot;
    a->stagger_cycles = 0;
    if (a->B >= 2 * grid) {
        const char *e = getenv("LSM_STAGGER");
        const long long n_used = (long long)(fe->ncols - 1) * fe->p.hop + fe->p.nwin;
        const long long per_utt = (long long)fe->p.channels * n_used * (a->mode ? 13 : 35) / 64;   // cycles of one SM's fp64 pipe
        // measured: no gain on this workload (5.51 vs 5.60 ms with the delay), so off unless LSM_STAGGER asks for it
        a->stagger_cycles = e ? atoi(e) : 0;
        (void)per_utt;
        if (a->stagger_cycles > 0) LSM_CUDA(ctx, cudaMemsetAsync(a->sm_rank, 0, 256 * sizeof(int), st));
    }
    return LSM_OK;
}


// ------------------------------------------------------------------------------------ lanes arrangement (K1a)
constexpr int kEnergyMaxUtt = 8192;     // utterances per pass of the two-kernel speculative path (energy buffer = 100 KB each)

// Can the speculative filter run lane = utterance?  8-sample steps must tile the window phases, rows must be 16-byte aligned.
static bool lanes_eligible(const lsm_frontend *fe, const float *d_pcm, bool standalone = false)
{
    const lsm_frontend_params &p = fe->p;
    // Stand-alone front end (stage 1 only: lsm_frontend_encode): on by default, 4.7 ms vs 5.5 ms per 2400 utterances.
    // Whole path: opt-in (LSM_LANES=1) - as two kernels it only ties with the single fused kernel (6.5 vs 6.5 ms per launch,
    // 5.7 vs 5.7 ms per step with two launches in flight) and gives up the zero-copy host path; see DESIGN.md.
    if (getenv("LSM_NO_LANES") || !d_pcm) return false;
    if (!standalone && !getenv("LSM_LANES")) return false;
    const int r_old = p.nwin - 2 * p.hop;
    return p.kind == LSM_FILTERBANK_GAMMATONE && fe->mode == LSM_FILTER_SPECULATIVE && p.hop % 8 == 0 && r_old % 8 == 0 &&
           p.n_samples % 4 == 0 && p.channels % kLanesJ == 0 && p.channels <= 256 && (((uintptr_t)d_pcm) & 15) == 0;
}

// device work list of the exact pass: [0] = count, then utterance indices
static int ensure_rerun(lsm_ctx *ctx, lsm_frontend *fe, int B)
{
    if (B <= fe->rerun_cap) return LSM_OK;
    { const int rc = lsm_frontend_wait_idle(ctx, fe); if (rc != LSM_OK) return rc; }
    if (fe->d_rerun) { cudaFree(fe->d_rerun); fe->d_rerun = nullptr; fe->rerun_cap = 0; }
    // two lists: one per scratch slot (two launches in flight)
    if (cudaMalloc((void **)&fe->d_rerun, 2 * sizeof(int) * ((size_t)B + 1)) != cudaSuccess) LSM_FAIL(ctx, LSM_ERR_NOMEM, "cudaMalloc for the re-execution list failed");
    fe->rerun_cap = B;
    return LSM_OK;
}

static int ensure_energy(lsm_ctx *ctx, lsm_frontend *fe, int B)
{
    if (B <= fe->energy_cap) return LSM_OK;
    { const int rc = lsm_frontend_wait_idle(ctx, fe); if (rc != LSM_OK) return rc; }     // nobody is still reading the old buffer
    if (fe->d_energy) { LSM_CUDA(ctx, cudaFree(fe->d_energy)); fe->d_energy = nullptr; fe->energy_cap = 0; }
    const size_t bytes = sizeof(double) * (size_t)B * fe->ncols * fe->p.channels;
    if (cudaMalloc((void **)&fe->d_energy, bytes) != cudaSuccess) LSM_FAIL(ctx, LSM_ERR_NOMEM, "cudaMalloc(%zu) for the energy planes failed", bytes);
    fe->energy_cap = B;
    return ensure_rerun(ctx, fe, B);
}

// K1a: raw window energies of the speculative cascade for utterances [0, B) into fe->d_energy
static int launch_energy(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, cudaStream_t st)
{
    const lsm_frontend_params &p = fe->p;
    int rc;
    if ((rc = lsm_frontend_order_before(ctx, fe, st)) != LSM_OK) return rc;
    if ((rc = ensure_energy(ctx, fe, B)) != LSM_OK) return rc;
    std::unique_ptr<EnergyArgs> ea_holder(new (std::nothrow) EnergyArgs);     // 12 KB of kernel parameters: not on the stack
    if (!ea_holder) LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
    EnergyArgs &ea = *ea_holder;
    ea.pcm = d_pcm; ea.energy = fe->d_energy;
    ea.B = B; ea.L = p.n_samples; ea.C = p.channels; ea.nwin = p.nwin; ea.hop = p.hop; ea.ncols = fe->ncols;
    memcpy(ea.coef, fe->h_lane_coef, sizeof(double) * 6 * p.channels);
    const int groups = (B + 31) / 32;
    // Units share an SM's fp64 pipes evenly, so an SM's finishing time is proportional to the work placed on it: hand every SM
    // the same number of big units (a multiple of the SM count; blocks are dealt round-robin) and cut what is left over into
    // single-channel units, which balance four times finer (all-big grid 4.8 ms, balanced 4.1 ms per 2400 utterances).
    const int upg = p.channels / kLanesJ;                       // big units per group
    const long long big_units = (long long)groups * upg;
    int big_groups = (int)((big_units / ctx->sm_count) * ctx->sm_count / upg);
    const char *e = getenv("LSM_LANES_SMALL_PCT");
    if (e) big_groups = groups - (groups * atoi(e) + 99) / 100;
    int small_groups = groups - big_groups;
    ea.big_groups = groups - small_groups;
    ea.n_units = ea.big_groups * (p.channels / kLanesJ) + small_groups * p.channels;
    int grid = ea.n_units;
    if (const char *k = getenv("LSM_K1A_PER_SM")) { const int v = atoi(k) * ctx->sm_count; if (v > 0 && v < grid) grid = v; }
    // co-residency experiment (LSM_K1A_CHAIN): energy kernels of one ctx run one after the other whatever their streams, so that
    // energy kernel i+1 starts together with the encoder/reservoir kernel of batch i instead of beside energy kernel i
    const bool chain = getenv("LSM_K1A_CHAIN") != nullptr;
    if (chain) {
        if (!ctx->ev_chain) LSM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming));
        else LSM_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_chain, 0));
    }
    gammatone_energy_kernel<<<grid, 32, 0, st>>>(ea);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    if (chain) LSM_CUDA(ctx, cudaEventRecord(ctx->ev_chain, st));
    return LSM_OK;
}

int lsm_launch_gammatone(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes,
                         double *d_spec_norm, cudaStream_t st)
{
    const lsm_frontend_params &p = fe->p;
    const bool lanes = !d_spec_norm && lanes_eligible(fe, d_pcm, true);
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
    fill_args(fe, d_pcm, B, d_spikes, d_spec_norm, &a);
    memset(&a.res, 0, sizeof(a.res));
    if (lanes && B > 0) {
        const int rc = launch_energy(ctx, fe, d_pcm, B, st);
        if (rc != LSM_OK) return rc;
        a.mode = 2;
        a.energy_in = fe->d_energy;
    }
    const int threads = ((p.channels + 31) / 32) * 32;
    const size_t smem = sizeof(double) * 2 * kChunkBlocks * p.hop + k1_pad_smem();
    const int grid = B < fe->grid ? B : fe->grid;
    if (grid <= 0) return LSM_OK;
    int *counter, rc;
    const int slot = lanes ? -1 : (int)(fe->slot_next++ & 1u);
    if ((rc = next_counter(ctx, fe, st, &counter, &a, grid, slot)) != LSM_OK) return rc;
    if (threads > 128) gammatone_encode_kernel<256, 2, 0, true><<<grid, threads, smem, st>>>(a, counter);
    else if (fe->minb == 4) gammatone_encode_kernel<128, 4, 0, true><<<grid, threads, smem, st>>>(a, counter);
    else if (fe->minb == 5) gammatone_encode_kernel<128, 5, 0, true><<<grid, threads, smem, st>>>(a, counter);
    else gammatone_encode_kernel<128, 6, 0, true><<<grid, threads, smem, st>>>(a, counter);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st, slot);
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
static int launch_fused_t(lsm_ctx *ctx, lsm_frontend *fe, const lsm_reservoir *res, GtArgs &a, int threads, size_t smem,
                          cudaStream_t st, bool launch = true, int *wave = nullptr, int max_grid = 0, int forced_slot = -2)
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
    // the lanes arrangement shares one energy buffer: exclusive; the single-kernel path alternates the two scratch slots
    const int slot = forced_slot != -2 ? forced_slot : ((a.mode == 2 || a.utt_list) ? -1 : (int)(fe->slot_next++ & 1u));
    if ((rc = next_counter(ctx, fe, st, &counter, &a, grid, slot)) != LSM_OK) return rc;
    if (res->lean) gammatone_encode_kernel<MAXT, MINB, FNPT, true><<<grid, threads, smem, st>>>(a, counter);
    else gammatone_encode_kernel<MAXT, MINB, FNPT, false><<<grid, threads, smem, st>>>(a, counter);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st, slot);
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
    const bool lanes = launch && lanes_eligible(fe, d_pcm);
    if (lanes && B > kEnergyMaxUtt) {
        const size_t spk_per = (size_t)p.channels * p.redundancy * p.n_bins * p.n_thresholds;
        const size_t feat_per = (size_t)__builtin_popcount(feature_mask & 0xFFu) * res->p.n_out;
        for (int off = 0; off < B; off += kEnergyMaxUtt) {
            const int n = B - off < kEnergyMaxUtt ? B - off : kEnergyMaxUtt;
            const int rc = fused_dispatch(ctx, fe, res, d_pcm + (size_t)off * p.n_samples, n,
                                          d_spikes_or_null ? d_spikes_or_null + off * spk_per : nullptr, feature_mask, nan_to_num,
                                          d_features + off * feat_per, st, true, nullptr);
            if (rc != LSM_OK) return rc;
        }
        return LSM_OK;
    }
    GtArgs a;
    fill_args(fe, d_pcm, B, d_spikes_or_null, nullptr, &a);
    lsm_reservoir_fill_args(res, nullptr, B, feature_mask, nan_to_num, d_features, nullptr, &a.res);
    if (lanes) {
        const int rc = launch_energy(ctx, fe, d_pcm, B, st);
        if (rc != LSM_OK) return rc;
        a.mode = 2;
        a.energy_in = fe->d_energy;
    }
    const int threads = p.channels;
    size_t smem = sizeof(double) * 2 * kChunkBlocks * p.hop;
    const size_t smem_res = lsm_res_smem_bytes(a.res.T, a.res.CW, threads * npt, a.res.N);
    if (smem_res > smem) smem = smem_res;
    if (lanes && threads == 128 && npt == 8 && res->lean && !getenv("LSM_LANES_OLD_KERNEL")) {
        // lanes arrangement, default shape: K1a (above) -> encode_reservoir_kernel -> exact pass over the flagged utterances
        int rc, per_sm = 0, *counter;
        if (smem_res > 48 * 1024)
            LSM_CUDA(ctx, cudaFuncSetAttribute(encode_reservoir_kernel<6, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_res));
        LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_reservoir_kernel<6, 8, true>, 128, smem_res));
        if (per_sm < 1) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "encode_reservoir_kernel does not fit on an SM");
        if (const char *e = getenv("LSM_ER_PER_SM")) { const int v = atoi(e); if (v > 0 && v < per_sm) per_sm = v; }   // experiment knob
        int grid = per_sm * ctx->sm_count;
        if (grid > B) grid = B;
        a.rerun_list = fe->d_rerun;          // exclusive launch: list of slot 0
        LSM_CUDA(ctx, cudaMemsetAsync(fe->d_rerun, 0, sizeof(int), st));
        if ((rc = next_counter(ctx, fe, st, &counter, &a, grid, -1)) != LSM_OK) return rc;
        a.stagger_cycles = 0;
        encode_reservoir_kernel<6, 8, true><<<grid, 128, smem_res, st>>>(a, counter);
        ctx->launches += 1;
        LSM_CUDA(ctx, cudaGetLastError());
        // exact pass: same kernel as the single-kernel path, exact mode, reading its work list from the device
        GtArgs x = a;
        x.mode = 0; x.energy_in = nullptr; x.rerun_list = nullptr;
        x.utt_list = fe->d_rerun + 1; x.utt_count = fe->d_rerun;
        return launch_fused_t<128, 5, 8>(ctx, fe, res, x, threads, smem, st, true, nullptr, 8);
    }
    if (launch && !lanes && a.mode == 1 && threads == 128 && npt == 8 && res->lean && getenv("LSM_SPLIT_EXACT")) {
        // opt-in variant: spec_fused_kernel (no exact-path code in the hot kernel), then the exact pass over the utterances it
        // flagged.  Measured slower than re-executing inside the kernel (6.37 vs 6.24 ms per step): the follow-up launch is a
        // one-CTA tail per step, and the smaller kernel (92 vs 96 registers) gains no occupancy.
        int rc, per_sm = 0, *counter;
        if ((rc = ensure_rerun(ctx, fe, B)) != LSM_OK) return rc;
        if (smem > 48 * 1024)
            LSM_CUDA(ctx, cudaFuncSetAttribute(spec_fused_kernel<5, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spec_fused_kernel<5, 8, true>, 128, smem));
        if (per_sm < 1) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "spec_fused_kernel does not fit on an SM");
        int grid = per_sm * ctx->sm_count;
        if (grid > fe->grid) grid = fe->grid;
        if (grid > B) grid = B;
        const int slot = (int)(fe->slot_next++ & 1u);
        int *list = fe->d_rerun + (size_t)slot * (fe->rerun_cap + 1);
        if ((rc = next_counter(ctx, fe, st, &counter, &a, grid, slot)) != LSM_OK) return rc;
        LSM_CUDA(ctx, cudaMemsetAsync(list, 0, sizeof(int), st));
        a.rerun_list = list;
        spec_fused_kernel<5, 8, true><<<grid, 128, smem, st>>>(a, counter);
        ctx->launches += 1;
        LSM_CUDA(ctx, cudaGetLastError());
        GtArgs x = a;
        x.mode = 0; x.rerun_list = nullptr; x.utt_list = list + 1; x.utt_count = list;
        return launch_fused_t<128, 5, 8>(ctx, fe, res, x, threads, smem, st, true, nullptr, 8, slot);
    }
    if (threads == 256) return launch_fused_t<256, 2, 4>(ctx, fe, res, a, threads, smem, st, launch, wave);
    if (npt == 8) {
        // 6 CTAs per SM (80 registers: only the rare exact re-execution spills) fill the issue slots the reservoir phases
        // leave better than 5 (measured 6.48 vs 6.79 ms per launch, 5.72 vs 5.82 ms per step); LSM_FUSED_MINB=5 / 4 select the others
        const char *fm = getenv("LSM_FUSED_MINB");
        const int fused_minb = fm ? atoi(fm) : 6;
        if (fused_minb >= 6) return launch_fused_t<128, 6, 8>(ctx, fe, res, a, threads, smem, st, launch, wave);
        if (fused_minb >= 5) return launch_fused_t<128, 5, 8>(ctx, fe, res, a, threads, smem, st, launch, wave);
        return launch_fused_t<128, 4, 8>(ctx, fe, res, a, threads, smem, st, launch, wave);
    }
    return launch_fused_t<128, 4, 16>(ctx, fe, res, a, threads, smem, st, launch, wave);
}

int lsm_gammatone_minb(void) { return k1_minb(); }

