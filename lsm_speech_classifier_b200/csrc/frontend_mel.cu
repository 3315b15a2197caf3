// K1m — mel front end: STFT power -> mel projection -> power_to_db -> min-max -> zoom -> hysteresis encoder.
//
// Replaces, per utterance, /root/reference/create_dataset.py:148-158 on the mel branch:
//   librosa.feature.melspectrogram(y, sr=16000, n_mels=n, hop_length=160)   :45-47  (librosa==0.11.0: n_fft 2048,
//        periodic hann, centre zero padding, |STFT|^2 in float32 from a complex64 STFT, Slaney mel basis)
//   librosa.power_to_db(spec, ref=np.max)                                     :48     (amin 1e-10, top_db 80)
//   min-max normalisation :62-67, scipy zoom 101 -> 100 :69-78, encoder :81-98, redundancy :101-104
//
// Mapping (default, "warp-per-frame"): two kernels.  mel_power_kernel: frames are independent; one warp per frame, the 1024
// packed complex points of the 2048-sample frame (fp64 window x float32 PCM) in its registers, radix-2 FFT in fp64 with one
// transpose through shared memory, untangle to the 1025 real-input bins, power in float32, mel projection by a lane-balanced
// schedule (ascending bins per band) -> mel power [B][frames][C] in device memory.  mel_finish_kernel: one CTA per utterance
// in flight: dB, floor, normalise, zoom, Schmitt triggers (the float32 twin of K1's epilogue) and, fused, the reservoir.
// LSM_MEL_BLOCK=1 selects round 1's arrangement instead (mel_encode_kernel: one 256-thread CTA per utterance, block-wide
// shared-memory FFT per frame, one band per thread), which also serves filter banks whose schedule does not fit.
//
// Bit-exactness: every operation and its order equals oracle/lsm_oracle.c (fft1024 / frame_power / mel_one);
// explicit *_rn intrinsics, no FMA contraction.  librosa's own FFT (pocketfft) has a different internal order;
// the oracle documents that and is itself compared with scipy.fft.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "reservoir_core.cuh"

namespace {

constexpr int kFft = 2048, kHalf = 1024, kThreads = 256;
constexpr int kFR = 1;       // frames per trip through the FFT's barriers (2 was measured: half the barriers per frame, but 4 instead of 5 CTAs per SM: 4.12 vs 3.83 ms)

struct MelArgs {
    const float *pcm;        // [B][L]
    const double *win;       // [2048]
    const double2 *tw;       // [512]
    const double2 *tw2;      // [1025]
    const float *mel_w;      // packed
    const int32_t *mel_lo, *mel_n, *mel_off;
    const int32_t *zoom_i0;
    const double *zoom_f;
    float *scratch;          // [grid][ncols][C]
    float *power;            // [B][ncols][C] mel power: mel_power_kernel -> mel_finish_kernel (which works in place)
    const uint2 *sched;      // [sched_len][32] projection schedule of the warp-per-frame kernel: (weight bits, bin | (band + 1) << 16)
    int sched_len;
    uint8_t *spikes;
    double *spec_norm;       // optional [B][C][nbins]
    int B, L, C, hop, ncols, nbins, K, R;
    double thr[8], lower[8];
    double2 tw1[31];         // twiddles of stages 1-5, stage s at offset 2^(s-1) - 1 (warp-per-frame kernel: constant-bank operands)
    ResArgs res;             // fused variants: the reservoir + readout that follow the encoder in the same CTA
};

__device__ __forceinline__ float block_reduce_f32(float v, bool want_max, float *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, v, o);
        v = want_max ? fmaxf(v, other) : fminf(v, other);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s_red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = want_max ? fmaxf(r, s_red[w]) : fminf(r, s_red[w]);
    return r;
}

// Shared-memory index swizzle (storage layout only - the butterflies and their order are the oracle's): the low four index
// bits are XORed with bits 2-5 and 6-9.  Found by exhaustive search over add/XOR paddings against the kernel's access patterns
// (bit-reversed store, the five two-stage passes, the untangle reads): 898 wavefront units per frame and array, vs 1474 for the
// previous "one pad double every 16" and 834 for a conflict-free layout; no padding words.
__device__ __forceinline__ int phys(int i) { return i ^ ((i >> 2) & 15) ^ ((i >> 6) & 15); }

// one radix-2 DIT butterfly, the oracle's operation order: t = w*a[q]; a[p] = u + t; a[q] = u - t
__device__ __forceinline__ void bfly(double &pr, double &pi, double &qr, double &qi, const double2 w)
{
    const double tr = __dsub_rn(__dmul_rn(w.x, qr), __dmul_rn(w.y, qi));
    const double ti = __dadd_rn(__dmul_rn(w.x, qi), __dmul_rn(w.y, qr));
    const double ur = pr, ui = pi;
    pr = __dadd_rn(ur, tr); pi = __dadd_rn(ui, ti);
    qr = __dsub_rn(ur, tr); qi = __dsub_rn(ui, ti);
}

// Per-utterance epilogue shared by the mel kernels: power_to_db, floor, min-max, zoom, Schmitt triggers; spikes to global rows
// and / or (FUSED) to the reservoir's bit plane at the start of smem_plan.  Called by all nthr threads of the CTA.
// The passes over the plane (global memory, L2-resident) are latency-bound, so every pass issues its loads in groups:
// the maximum over eight frames at a time, the dB pass over four, and normalisation is folded into the zoom, four output
// bins (eight plane values) at a time.
__device__ __forceinline__ float mel_db(const float p, const float ref_db)
{
    return __fsub_rn(__fmul_rn(10.0f, (float)lsm_log10((double)fmaxf(p, 1e-10f))), ref_db);
}

template <int FUSED>
__device__ __forceinline__ void mel_epilogue(const MelArgs &a, const int utt, float *__restrict__ plane, unsigned char *smem_plan,
                                             float *s_red, const int tid, const int nthr)
{
    const int C = a.C, ncols = a.ncols;
    // ---- power_to_db(ref=np.max, amin=1e-10, top_db=80), create_dataset.py:48
    float tmax = -INFINITY;
    for (int m = tid; m < C; m += nthr) {
        const float *col = plane + m;
        int t = 0;
        for (; t + 8 <= ncols; t += 8) {
            float p[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) p[u] = col[(size_t)(t + u) * C];
#pragma unroll
            for (int u = 0; u < 8; ++u) tmax = fmaxf(tmax, p[u]);
        }
        for (; t < ncols; ++t) tmax = fmaxf(tmax, col[(size_t)t * C]);
    }
    const float ref = block_reduce_f32(tmax, true, s_red);
    const double refd = ((double)ref > 1e-10) ? (double)ref : 1e-10;               // scalar path is float64 (numpy 1.26)
    const float ref_db = (float)__dmul_rn(10.0, lsm_log10(refd));
    float dmax = -INFINITY, dmin = INFINITY;
    for (int m = tid; m < C; m += nthr) {
        float *col = plane + m;
        int t = 0;
        for (; t + 4 <= ncols; t += 4) {
            float p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = col[(size_t)(t + u) * C];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float d = mel_db(p[u], ref_db);
                col[(size_t)(t + u) * C] = d;
                dmax = fmaxf(dmax, d); dmin = fminf(dmin, d);
            }
        }
        for (; t < ncols; ++t) {
            const float d = mel_db(col[(size_t)t * C], ref_db);
            col[(size_t)t * C] = d;
            dmax = fmaxf(dmax, d); dmin = fminf(dmin, d);
        }
    }
    const float mx = block_reduce_f32(dmax, true, s_red);
    const float rawmin = block_reduce_f32(dmin, false, s_red);
    const float floor_db = (float)__dsub_rn((double)mx, 80.0);
    const float mn = fmaxf(rawmin, floor_db);                                       // min of the clamped plane
    const float diff = __fsub_rn(mx, mn);
    const bool degenerate = (double)diff < 1e-8;                                    // create_dataset.py:64-65
    const float den = (float)__dadd_rn((double)diff, 1e-8);

    if (FUSED) __syncthreads();                         // every thread is done with the FFT buffers: the bit plane takes their place
    unsigned *s_bits = reinterpret_cast<unsigned *>(smem_plan);
    const int CW = (C + 31) >> 5;
    const bool same_grid = ncols == a.nbins;
    for (int m = tid; m < C; m += nthr) {             // whole warps (fused pairs have C % 32 == 0)
        const int T = a.nbins * a.K;
        uint8_t *row0 = a.spikes ? a.spikes + ((size_t)utt * C * a.R + (size_t)m * a.R) * T : nullptr;
        double *dump = a.spec_norm ? a.spec_norm + ((size_t)utt * C + m) * a.nbins : nullptr;
        const float *col = plane + m;
        unsigned on = 0;
        for (int j0 = 0; j0 < a.nbins; j0 += 4) {
            // normalised (create_dataset.py:62-67) and zoomed (:69-78) values of four bins; the loads first
            int i0[4];
            double f[4];
            float lo[4], hi[4], v4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = min(j0 + u, a.nbins - 1);
                i0[u] = same_grid ? j : __ldg(a.zoom_i0 + j);
                f[u] = same_grid ? 0.0 : __ldg(a.zoom_f + j);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                lo[u] = col[(size_t)i0[u] * C];
                hi[u] = (!same_grid && i0[u] + 1 < ncols) ? col[(size_t)(i0[u] + 1) * C] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float nlo = __fdiv_rn(__fsub_rn(fmaxf(lo[u], floor_db), mn), den);
                if (degenerate) v4[u] = 0.0f;
                else if (same_grid) v4[u] = nlo;
                else {
                    double vd = __dmul_rn((double)nlo, __dsub_rn(1.0, f[u]));
                    if (i0[u] + 1 < ncols) {
                        const float nhi = __fdiv_rn(__fsub_rn(fmaxf(hi[u], floor_db), mn), den);
                        vd = __dadd_rn(vd, __dmul_rn((double)nhi, f[u]));
                    }
                    v4[u] = (float)vd;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                if (j >= a.nbins) break;
                const float v = v4[u];
                if (dump) dump[j] = (double)v;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k < a.K) {
                        const bool is_on = (on >> k) & 1u;
                        if (!is_on && v > (float)a.thr[k]) on |= (1u << k);
                        else if (is_on && v < (float)a.lower[k]) on &= ~(1u << k);
                    }
                }
                if (FUSED) {
                    for (int k = 0; k < a.K; ++k) {
                        const unsigned word = __ballot_sync(0xffffffffu, (on >> k) & 1u);
                        if ((m & 31) == 0) s_bits[(j * a.K + k) * CW + (m >> 5)] = word;
                    }
                }
                if (row0) {
                    for (int r = 0; r < a.R; ++r) {
                        uint8_t *row = row0 + (size_t)r * T + (size_t)j * a.K;
                        if (a.K == 4) {
                            const unsigned w = (on & 1u) | ((on & 2u) << 7) | ((on & 4u) << 14) | ((on & 8u) << 21);
                            *reinterpret_cast<uint32_t *>(row) = w;
                        } else {
                            for (int k = 0; k < a.K; ++k) row[k] = (on >> k) & 1u;
                        }
                    }
                }
            }
        }
    }
}

// FUSED: 0 = front end only (spike trains to global memory); 1 / 2 = the reservoir (lean / generic layout, 4 neurons per thread:
// reservoirs of up to 1024 neurons) and the readout follow in the same CTA, the spike train handed over as bits in shared
// memory (one ballot word per warp of 32 channels and time step), exactly as the fused gammatone kernel does.  The FFT
// buffers and the reservoir's shared-memory plan share one dynamic allocation; the twiddle tables are reloaded per utterance.
constexpr size_t kMelSmemBytes = sizeof(double) * 2 * kHalf * kFR + sizeof(double2) * kHalf + sizeof(float) * (kHalf + 8) * kFR;

template <int FUSED>
__global__ void __launch_bounds__(kThreads, FUSED ? 3 : 5) mel_encode_kernel(const MelArgs a, int *next_utt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_re = reinterpret_cast<double *>(smem_raw), *s_im = s_re + kFR * kHalf;      // [kFR][kHalf] each
    double2 *s_tw = reinterpret_cast<double2 *>(s_im + kFR * kHalf);   // per-stage twiddle tables, stage s at offset 2^(s-1) - 1: T_s[j] = tw[j * (1024 >> s)], j < 2^(s-1)
                                                                 // (contiguous in j: the strided reads of one shared table were up to 16-way bank conflicted)
    float *s_S = reinterpret_cast<float *>(s_tw + kHalf);
    __shared__ float s_red[kThreads / 32];
    __shared__ int s_utt;
    __shared__ int s_cnt[5];

    const int tid = threadIdx.x;
    const int C = a.C, ncols = a.ncols;
    float *plane = a.scratch + (size_t)blockIdx.x * ncols * C;     // mel power / dB plane [ncols][C]

    for (;;) {
        if (tid == 0) s_utt = atomicAdd(next_utt, 1);
        __syncthreads();
        const int utt = s_utt;
        if (utt >= a.B) break;
        for (int q = tid; q < kHalf - 1; q += kThreads) {
            const int s = 32 - __clz(q + 1);                 // stage whose table holds entry q
            const int j = q + 1 - (1 << (s - 1));
            s_tw[q] = __ldg(a.tw + j * (kHalf >> s));
        }
        const float *pcm = a.pcm + (size_t)utt * a.L;

        // kFR frames per trip through the barriers: the same butterflies, half the barriers per frame and twice the
        // independent work between two of them
        for (int t0 = 0; t0 < ncols; t0 += kFR) {
            // ---- windowed frames, packed z[j] = x[2j] + i x[2j+1], stored bit-reversed for the DIT FFT
#pragma unroll
            for (int f = 0; f < kFR; ++f) {
                const int t = t0 + f;
                if (t >= ncols) break;
                const int start = t * a.hop - kHalf;               // centre padding: n_fft/2 zeros each side
                double *re = s_re + f * kHalf, *im = s_im + f * kHalf;
                for (int j = tid; j < kHalf; j += kThreads) {
                    const int i0 = start + 2 * j, i1 = i0 + 1;
                    const double x0 = (i0 >= 0 && i0 < a.L) ? __dmul_rn(__ldg(a.win + 2 * j), (double)__ldg(pcm + i0)) : 0.0;
                    const double x1 = (i1 >= 0 && i1 < a.L) ? __dmul_rn(__ldg(a.win + 2 * j + 1), (double)__ldg(pcm + i1)) : 0.0;
                    const int r = phys((int)(__brev((unsigned)j) >> 22));
                    re[r] = x0; im[r] = x1;
                }
            }
            __syncthreads();
            // ---- 10 radix-2 stages as 5 passes of two stages each: a thread owns the 4 points p, p+h, p+2h, p+3h
            //      and runs stage s on (p,p+h), (p+2h,p+3h) then stage s+1 on (p,p+2h), (p+h,p+3h) in registers.
            //      Exactly the radix-2 operations of the oracle, with half the barriers and shared-memory traffic.
#pragma unroll
            for (int s = 1; s <= 9; s += 2) {
                const int h = 1 << (s - 1);
                const int j = tid & (h - 1);
                const int p = ((tid >> (s - 1)) << (s + 1)) + j;
                const int i0 = phys(p), i1 = phys(p + h), i2 = phys(p + 2 * h), i3 = phys(p + 3 * h);
                const double2 ws = s_tw[h - 1 + j];
                const double2 wa = s_tw[2 * h - 1 + j], wb = s_tw[2 * h - 1 + j + h];
#pragma unroll
                for (int f = 0; f < kFR; ++f) {
                    if (t0 + f >= ncols) break;
                    double *re = s_re + f * kHalf, *im = s_im + f * kHalf;
                    double r0 = re[i0], m0 = im[i0], r1 = re[i1], m1 = im[i1];
                    double r2 = re[i2], m2 = im[i2], r3 = re[i3], m3 = im[i3];
                    bfly(r0, m0, r1, m1, ws);
                    bfly(r2, m2, r3, m3, ws);
                    bfly(r0, m0, r2, m2, wa);
                    bfly(r1, m1, r3, m3, wb);
                    re[i0] = r0; im[i0] = m0; re[i1] = r1; im[i1] = m1;
                    re[i2] = r2; im[i2] = m2; re[i3] = r3; im[i3] = m3;
                }
                __syncthreads();
            }
            // ---- real-input untangle, complex64 rounding, |.|^2 in float32
            for (int k = tid; k <= kHalf; k += kThreads) {
                const int k1 = k & (kHalf - 1), k2 = (kHalf - k) & (kHalf - 1);
                const double2 w = __ldg(a.tw2 + k);
                const double c_re = w.y, c_im = -w.x;                                   // -i * W
                const int p1 = phys(k1), p2 = phys(k2);
#pragma unroll
                for (int f = 0; f < kFR; ++f) {
                    if (t0 + f >= ncols) break;
                    const double *re = s_re + f * kHalf, *im = s_im + f * kHalf;
                    const double zr = re[p1], zi = im[p1], cr = re[p2], ci = -im[p2];
                    const double ar = __dmul_rn(0.5, __dadd_rn(zr, cr)), ai = __dmul_rn(0.5, __dadd_rn(zi, ci));
                    const double br = __dmul_rn(0.5, __dsub_rn(zr, cr)), bi = __dmul_rn(0.5, __dsub_rn(zi, ci));
                    const double xr = __dadd_rn(ar, __dsub_rn(__dmul_rn(c_re, br), __dmul_rn(c_im, bi)));
                    const double xi = __dadd_rn(ai, __dadd_rn(__dmul_rn(c_re, bi), __dmul_rn(c_im, br)));
                    const float r32 = (float)xr, i32 = (float)xi;
                    const double r64 = (double)r32, i64 = (double)i32;
                    const float mag = (float)__dsqrt_rn(__dadd_rn(__dmul_rn(r64, r64), __dmul_rn(i64, i64)));
                    s_S[f * (kHalf + 8) + k] = __fmul_rn(mag, mag);
                }
            }
            __syncthreads();
            // ---- mel projection: one band per thread and frame, ascending bins, float32 multiply then add
            for (int q = tid; q < C * kFR; q += kThreads) {
                const int f = q / C, m = q - f * C;
                if (t0 + f >= ncols) continue;
                const float *w = a.mel_w + __ldg(a.mel_off + m);
                const int lo = __ldg(a.mel_lo + m), n = __ldg(a.mel_n + m);
                const float *S = s_S + f * (kHalf + 8);
                float acc = 0.0f;
                for (int qq = 0; qq < n; ++qq) acc = __fadd_rn(acc, __fmul_rn(__ldg(w + qq), S[lo + qq]));
                plane[(size_t)(t0 + f) * C + m] = acc;
            }
            // the next trip's loads touch s_re/s_im only, and ten barriers separate this read of s_S from its next write
        }

        mel_epilogue<FUSED>(a, utt, plane, smem_raw, s_red, tid, kThreads);
        if (FUSED) {
            __syncthreads();                                 // the bit plane is complete
            reservoir_simulate<4, FUSED == 1, false, 0>(a.res, utt, smem_raw, s_cnt, tid, kThreads);
        }
        __syncthreads();
    }
}

// ---- warp-per-frame arrangement (default): mel_power_kernel (STFT power -> mel power, frames are independent) followed by
//      mel_finish_kernel (per-utterance epilogue, optionally the reservoir).  The same butterflies, untangle and projection as
//      mel_encode_kernel - every output of a radix-2 butterfly depends only on its two inputs and its twiddle, so the schedule is
//      free - but one warp owns a frame and the 1024 complex points live in its registers, 32 per lane:
//        stages 1-5   lane l holds positions 32 l .. 32 l + 31 (a closed group of the first five stages); the twiddles depend on the
//                     register slot only and come from the kernel-parameter constant bank
//        transpose    through the warp's own 16 KB of shared memory (XOR-swizzled 16-byte slots, conflict-free both ways)
//        stages 6-10  lane l holds positions l, l + 32, ..: twiddle T_s[l + 32 kk], one conflict-free 16-byte load per 16 / 2^(s-6)
//                     butterflies
//        untangle     bins l, l + 32, ..: own value from registers, the mirrored bin from shared memory
//        projection   the power spectrum in shared memory, bands l, l + 32, .. per lane, four bands interleaved
//      No block barrier at all in the power kernel: the block kernel spent 5.7 stall cycles per issue at its six barriers per
//      frame (profiles/r2_summary.md, capture L).  A frame needs ~250 registers per lane, i.e. 8 warps per SM, so everything that
//      is latency-bound (dB conversion, normalisation, Schmitt triggers, the reservoir) runs in the second kernel at full occupancy.
constexpr int kPT = 256, kPW = kPT / 32;
constexpr int kSchedMax = 160;                                               // longest projection schedule held in shared memory (steps)
constexpr int kZeroBin = kHalf + 2;                                          // a spectrum slot that always holds 0 (schedule padding)
constexpr size_t kMelPowerSmemBase = sizeof(double2) * kHalf * (2 + kPW) + sizeof(double2) * (kHalf + 2);

__host__ __device__ constexpr int brev5(int r) { return ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | ((r & 8) >> 2) | ((r & 16) >> 4); }

// The two warps that share an SM sub-partition (warp w and w + 4) meet at a 64-thread named barrier after every phase while both
// have a frame, so that they walk the ~80 KB of straight-line code together and share its instruction fetches (3.16 -> 3.05 ms;
// all eight warps in step: 3.06).
__device__ __forceinline__ void pair_sync(const bool both, const int bar)
{
    if (both) asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void warp_frame(const MelArgs &a, const float *pcm, const int t, const double2 *s_win, const double2 *s_tw,
                                           const double2 *s_tw2, const uint2 *s_sched, double2 *buf, float *plane_row, const int lane,
                                           const bool both, const int bar)
{
    double re[32], im[32];
    const int start = t * a.hop - kHalf;                           // centre padding: n_fft/2 zeros each side
    const int jl = (int)(__brev((unsigned)lane) >> 27);
    // ---- windowed frame, packed z[j] = x[2j] + i x[2j+1]; position 32 l + r of the bit-reversed order holds j = 32 brev5(r) + brev5(l)
    if (start >= 0 && start + kFft <= a.L && ((reinterpret_cast<uintptr_t>(pcm + start) & 7) == 0)) {
        const float2 *x2 = reinterpret_cast<const float2 *>(pcm + start);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int j = 32 * brev5(r) + jl;
            const double2 w = s_win[j];
            const float2 x = __ldg(x2 + j);
            re[r] = __dmul_rn(w.x, (double)x.x);
            im[r] = __dmul_rn(w.y, (double)x.y);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int j = 32 * brev5(r) + jl;
            const double2 w = s_win[j];
            const int i0 = start + 2 * j, i1 = i0 + 1;
            re[r] = (i0 >= 0 && i0 < a.L) ? __dmul_rn(w.x, (double)__ldg(pcm + i0)) : 0.0;
            im[r] = (i1 >= 0 && i1 < a.L) ? __dmul_rn(w.y, (double)__ldg(pcm + i1)) : 0.0;
        }
    }
    pair_sync(both, bar);
    // ---- stages 1-5 in registers
#pragma unroll
    for (int s = 1; s <= 5; ++s) {
        const int h = 1 << (s - 1);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            if (r & h) continue;
            bfly(re[r], im[r], re[r + h], im[r + h], a.tw1[h - 1 + (r & (h - 1))]);
        }
    }
    pair_sync(both, bar);
    // ---- transpose: position P = 32 g + r lives in 16-byte slot 32 g + (r ^ g)
#pragma unroll
    for (int r = 0; r < 32; ++r) buf[32 * lane + (r ^ lane)] = make_double2(re[r], im[r]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const double2 v = buf[32 * k + (lane ^ k)];
        re[k] = v.x; im[k] = v.y;
    }
    // ---- stages 6-10: slot k holds position 32 k + lane; stage s pairs slots k, k + 2^(s-6); twiddle index (32 k + lane) mod 2^(s-1)
#pragma unroll
    for (int s = 6; s <= 10; ++s) {
        const int hk = 1 << (s - 6), h = 32 * hk;
#pragma unroll
        for (int kk = 0; kk < hk; ++kk) {
            const double2 w = s_tw[h - 1 + lane + 32 * kk];
#pragma unroll
            for (int b = 0; b < 16 / hk; ++b) {
                const int k = kk + 2 * hk * b;
                bfly(re[k], im[k], re[k + hk], im[k + hk], w);
            }
        }
    }
    // ---- real-input untangle, complex64 rounding, |.|^2 in float32: bins lane + 32 i (own slot i) and, on lane 0, bin 1024
    pair_sync(both, bar);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 32; ++k) buf[32 * k + (lane ^ k)] = make_double2(re[k], im[k]);
    __syncwarp();
    float pw[33];
#pragma unroll
    for (int i = 0; i <= 32; ++i) {
        const int kb = i < 32 ? lane + 32 * i : kHalf;
        const int k2 = (kHalf - kb) & (kHalf - 1);
        const double2 w = s_tw2[kb];
        const double c_re = w.y, c_im = -w.x;                                   // -i * W
        const double2 mz = buf[(k2 & ~31) | ((k2 ^ (k2 >> 5)) & 31)];
        const double zr = i < 32 ? re[i] : re[0], zi = i < 32 ? im[i] : im[0], cr = mz.x, ci = -mz.y;
        const double ar = __dmul_rn(0.5, __dadd_rn(zr, cr)), ai = __dmul_rn(0.5, __dadd_rn(zi, ci));
        const double br = __dmul_rn(0.5, __dsub_rn(zr, cr)), bi = __dmul_rn(0.5, __dsub_rn(zi, ci));
        const double xr = __dadd_rn(ar, __dsub_rn(__dmul_rn(c_re, br), __dmul_rn(c_im, bi)));
        const double xi = __dadd_rn(ai, __dadd_rn(__dmul_rn(c_re, bi), __dmul_rn(c_im, br)));
        const float r32 = (float)xr, i32 = (float)xi;
        const double r64 = (double)r32, i64 = (double)i32;
        const float mag = (float)sqrt_rn_inline(__dadd_rn(__dmul_rn(r64, r64), __dmul_rn(i64, i64)));
        pw[i] = __fmul_rn(mag, mag);
    }
    pair_sync(both, bar);
    __syncwarp();                                                   // every lane has read its mirrored bins: the spectrum takes the buffer
    float *S = reinterpret_cast<float *>(buf);
#pragma unroll
    for (int i = 0; i < 32; ++i) S[lane + 32 * i] = pw[i];
    if (lane == 0) S[kHalf] = pw[32];
    if (lane == 1) S[kZeroBin] = 0.0f;
    __syncwarp();
    // ---- mel projection: ascending bins, float32 multiply then add.  The host deals the bands out to the 32 lanes so that every
    //      lane has the same number of terms (lsm_mel_create): step k of a lane is one term of one of its bands, a band's last term
    //      carries the band's index and the sum is stored; lanes that run out of terms add 0 x 0 to a sum nobody reads.
    float acc = 0.0f;
#pragma unroll 4
    for (int k = 0; k < a.sched_len; ++k) {
        const uint2 e = s_sched[k * 32 + lane];
        const float nacc = __fadd_rn(acc, __fmul_rn(__uint_as_float(e.x), S[e.y & 0xffffu]));
        const unsigned out = e.y >> 16;
        if (out) plane_row[out - 1] = nacc;
        acc = out ? 0.0f : nacc;
    }
    __syncwarp();                                                   // the next frame's transpose overwrites S
}

__global__ void __launch_bounds__(kPT, 1) mel_power_kernel(const __grid_constant__ MelArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *s_win = reinterpret_cast<double2 *>(smem_raw);          // [1024] (win[2j], win[2j+1])
    double2 *s_tw = s_win + kHalf;                                   // per-stage tables as in mel_encode_kernel
    double2 *s_tw2 = s_tw + kHalf;                                   // [1025] (+ 1 pad)
    double2 *s_x = s_tw2 + kHalf + 2;                                // [kPW][1024]
    uint2 *s_sched = reinterpret_cast<uint2 *>(s_x + (size_t)kPW * kHalf);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < kHalf; j += kPT) s_win[j] = make_double2(__ldg(a.win + 2 * j), __ldg(a.win + 2 * j + 1));
    for (int q = tid; q < kHalf - 1; q += kPT) {
        const int s = 32 - __clz(q + 1);
        s_tw[q] = __ldg(a.tw + (q + 1 - (1 << (s - 1))) * (kHalf >> s));
    }
    for (int k = tid; k <= kHalf; k += kPT) s_tw2[k] = __ldg(a.tw2 + k);
    for (int q = tid; q < a.sched_len * 32; q += kPT) s_sched[q] = __ldg(a.sched + q);
    __syncthreads();
    // a CTA owns a contiguous run of frames and its warps walk it side by side: consecutive frames share 92 % of their samples, so
    // the run streams through L1 once; each warp prefetches the hop its next frame adds
    const long long frames = (long long)a.B * a.ncols;
    const long long f0 = frames * blockIdx.x / gridDim.x, f1 = frames * (blockIdx.x + 1) / gridDim.x;
    for (long long f = f0 + warp; f < f1; f += kPW) {
        const int utt = (int)(f / a.ncols), t = (int)(f - (long long)utt * a.ncols);
        if (f + kPW < f1 && lane < 8) {
            const long long fn = f + kPW;
            const int un = (int)(fn / a.ncols), tn = (int)(fn - (long long)un * a.ncols);
            long long i = (long long)tn * a.hop + kHalf - a.hop + 32 * lane;        // the last hop of the next frame, 128 bytes per lane
            i = i < 0 ? 0 : (i > a.L - 1 ? a.L - 1 : i);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.pcm + (size_t)un * a.L + i));
        }
        warp_frame(a, a.pcm + (size_t)utt * a.L, t, s_win, s_tw, s_tw2, s_sched, s_x + (size_t)warp * kHalf,
                   a.power + (size_t)f * a.C, lane, f - warp + (warp | 4) < f1, 1 + (warp & 3));
    }
}

// second kernel: one CTA per utterance in flight; FUSED as in mel_encode_kernel (the reservoir plan is the whole dynamic allocation).
// NPT = neurons per thread of the fused reservoir: 8 with 128 threads when the channels fit (every thread has a channel in the
// epilogue and 8 neurons afterwards, as in the fused gammatone kernel), 4 with 256 threads for wider filter banks.
template <int FUSED, int NPT>
__global__ void __launch_bounds__(NPT == 8 ? 128 : kThreads, FUSED == 1 ? (NPT == 8 ? 6 : 4) : FUSED == 2 ? (NPT == 8 ? 5 : 3) : 5) mel_finish_kernel(const MelArgs a, int *next_utt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_red[kThreads / 32];
    __shared__ int s_utt;
    __shared__ int s_cnt[5];
    const int tid = threadIdx.x;
    for (;;) {
        if (tid == 0) s_utt = atomicAdd(next_utt, 1);
        __syncthreads();
        const int utt = s_utt;
        if (utt >= a.B) break;
        mel_epilogue<FUSED>(a, utt, a.power + (size_t)utt * a.ncols * a.C, smem_raw, s_red, tid, (int)blockDim.x);
        if (FUSED) {
            __syncthreads();                                 // the bit plane is complete
            reservoir_simulate<NPT, FUSED == 1, false, 0>(a.res, utt, smem_raw, s_cnt, tid, (int)blockDim.x);
        }
        __syncthreads();
    }
}

template <int FUSED, int NPT>
static int mel_finish_fused_launch(lsm_ctx *ctx, const MelArgs &a, int *counter, size_t smem, int B, cudaStream_t st)
{
    const int threads = NPT == 8 ? 128 : kThreads;
    int per_sm = 0;
    LSM_CUDA(ctx, cudaFuncSetAttribute(mel_finish_kernel<FUSED, NPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mel_finish_kernel<FUSED, NPT>, threads, smem));
    if (per_sm < 1) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "fused mel kernel does not fit on an SM");
    int grid = per_sm * ctx->sm_count;
    if (grid > B) grid = B;
    mel_finish_kernel<FUSED, NPT><<<grid, threads, smem, st>>>(a, counter);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return LSM_OK;
}

template <typename T>
int up(lsm_ctx *ctx, T **dst, const T *src, size_t n)
{
    *dst = nullptr;
    LSM_CUDA(ctx, cudaMalloc((void **)dst, (n ? n : 1) * sizeof(T)));
    if (src && n) LSM_CUDA(ctx, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return LSM_OK;
}

}  // namespace

// stages 1-5 of the FFT use twiddles tw[j * (1024 >> s)], j < 2^(s-1): 31 values, kept on the host for the kernel parameters
static void mel_host_tw1(lsm_frontend *fe, const double *h_tw)
{
    for (int s = 1; s <= 5; ++s)
        for (int j = 0; j < (1 << (s - 1)); ++j) {
            const int q = j * (kHalf >> s), o = (1 << (s - 1)) - 1 + j;
            fe->mel_tw1[2 * o] = h_tw[2 * q];
            fe->mel_tw1[2 * o + 1] = h_tw[2 * q + 1];
        }
}

int lsm_mel_set_tables(lsm_ctx *ctx, lsm_frontend *fe, const double *h_win, const double *h_tw, const double *h_tw2)
{
    LSM_CUDA(ctx, cudaMemcpy(fe->d_window, h_win, sizeof(double) * kFft, cudaMemcpyHostToDevice));
    LSM_CUDA(ctx, cudaMemcpy(fe->d_twiddle, h_tw, sizeof(double) * 2 * (kHalf / 2), cudaMemcpyHostToDevice));
    LSM_CUDA(ctx, cudaMemcpy(fe->d_twiddle2, h_tw2, sizeof(double) * 2 * (kHalf + 1), cudaMemcpyHostToDevice));
    mel_host_tw1(fe, h_tw);
    return LSM_OK;
}

int lsm_mel_create(lsm_ctx *ctx, lsm_frontend *fe, const float *h_basis)
{
    const lsm_frontend_params &p = fe->p;
    if (p.n_fft != kFft) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "mel n_fft %d: only 2048 (librosa's default) is built", p.n_fft);
    const int C = p.channels, nb = 1 + p.n_fft / 2;
    // pack the non-zero run of every triangle (zeros add exactly nothing)
    std::vector<float> w;
    std::vector<int32_t> lo(C, 0), n(C, 0), off(C, 0);
    for (int m = 0; m < C; ++m) {
        int a = -1, b = -1;
        for (int f = 0; f < nb; ++f)
            if (h_basis[(size_t)m * nb + f] != 0.0f) { if (a < 0) a = f; b = f + 1; }
        off[m] = (int32_t)w.size();
        if (a >= 0) { lo[m] = a; n[m] = b - a; w.insert(w.end(), h_basis + (size_t)m * nb + a, h_basis + (size_t)m * nb + b); }
    }
    int rc = up(ctx, &fe->d_mel_w, w.data(), w.size());
    if (rc == LSM_OK) rc = up(ctx, &fe->d_mel_lo, lo.data(), (size_t)C);
    if (rc == LSM_OK) rc = up(ctx, &fe->d_mel_n, n.data(), (size_t)C);
    if (rc == LSM_OK) rc = up(ctx, &fe->d_mel_off, off.data(), (size_t)C);
    // default tables from the host libm; callers that need bit parity with another implementation override them
    std::vector<double> win(kFft), tw(kHalf), tw2(2 * (kHalf + 1));
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < kFft; ++i) win[i] = 0.5 - 0.5 * cos(2 * pi * i / kFft);
    for (int q = 0; q < kHalf / 2; ++q) { tw[2 * q] = cos(2 * pi * q / kHalf); tw[2 * q + 1] = -sin(2 * pi * q / kHalf); }
    for (int k = 0; k <= kHalf; ++k) { tw2[2 * k] = cos(2 * pi * k / kFft); tw2[2 * k + 1] = -sin(2 * pi * k / kFft); }
    if (rc == LSM_OK) rc = up(ctx, &fe->d_window, win.data(), win.size());
    if (rc == LSM_OK) rc = up(ctx, (double **)&fe->d_twiddle, tw.data(), tw.size());
    if (rc == LSM_OK) rc = up(ctx, (double **)&fe->d_twiddle2, tw2.data(), tw2.size());
    if (rc != LSM_OK) return rc;
    mel_host_tw1(fe, tw.data());
    int per_sm = 0;
    // projection schedule of the warp-per-frame kernel: bands dealt to 32 lanes, longest first, each to the lane with the fewest terms
    {
        std::vector<int> order(C), total(32, 0);
        std::vector<std::vector<int>> mine(32);
        for (int m = 0; m < C; ++m) order[m] = m;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return n[x] > n[y]; });
        for (int m : order) {
            int best = 0;
            for (int l = 1; l < 32; ++l) if (total[l] < total[best]) best = l;
            mine[best].push_back(m);
            total[best] += n[m] > 0 ? n[m] : 1;
        }
        int len = 0;
        for (int l = 0; l < 32; ++l) len = total[l] > len ? total[l] : len;
        fe->mel_sched_len = len;
        if (len <= kSchedMax) {
            std::vector<uint32_t> sched((size_t)len * 32 * 2);
            for (int l = 0; l < 32; ++l) {
                int k = 0;
                auto put = [&](float wv, uint32_t bin, uint32_t out) {
                    uint32_t bits; memcpy(&bits, &wv, 4);
                    sched[((size_t)k * 32 + l) * 2] = bits; sched[((size_t)k * 32 + l) * 2 + 1] = bin | (out << 16); ++k;
                };
                for (int m : mine[l]) {
                    if (n[m] == 0) put(0.0f, kZeroBin, (uint32_t)m + 1);               // an empty triangle: the sum is 0
                    for (int q = 0; q < n[m]; ++q) put(w[off[m] + q], (uint32_t)(lo[m] + q), q == n[m] - 1 ? (uint32_t)m + 1 : 0u);
                }
                while (k < len) put(0.0f, kZeroBin, 0u);
            }
            if ((rc = up(ctx, &fe->d_mel_sched, sched.data(), sched.size())) != LSM_OK) return rc;
            const size_t smem = kMelPowerSmemBase + sizeof(uint2) * 32 * (size_t)len;
            LSM_CUDA(ctx, cudaFuncSetAttribute(mel_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fe->finish_threads = C <= 64 ? 64 : (C + 31) & ~31;
            if (fe->finish_threads > kThreads) fe->finish_threads = kThreads;
            LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mel_finish_kernel<0, 4>, fe->finish_threads, 0));
            fe->grid_warp = per_sm * ctx->sm_count;
        }
    }
    LSM_CUDA(ctx, cudaFuncSetAttribute(mel_encode_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMelSmemBytes));
    LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mel_encode_kernel<0>, kThreads, kMelSmemBytes));
    if (per_sm < 1) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "mel kernel does not fit on an SM");
    fe->grid = per_sm * ctx->sm_count;
    rc = up<float>(ctx, &fe->d_mel_scratch, nullptr, (size_t)fe->grid * fe->ncols * C);
    if (rc == LSM_OK) rc = up<int>(ctx, &fe->d_counters, nullptr, 64);
    return rc;
}

void lsm_mel_destroy(lsm_frontend *fe)
{
    cudaFree(fe->d_mel_w); cudaFree(fe->d_mel_lo); cudaFree(fe->d_mel_n); cudaFree(fe->d_mel_off);
    cudaFree(fe->d_window); cudaFree(fe->d_twiddle); cudaFree(fe->d_twiddle2); cudaFree(fe->d_mel_scratch);
    cudaFree(fe->d_mel_power); cudaFree(fe->d_mel_sched);
}

// Can this mel front end hand its spike trains to this reservoir inside one kernel?  Whole warps of channels, no redundancy,
// a reservoir of at most 1024 neurons (4 per thread of the 256-thread CTA) whose shared-memory plan fits beside nothing else.
bool lsm_mel_fused_ok(const lsm_frontend *fe, const lsm_reservoir *res)
{
    const lsm_frontend_params &p = fe->p;
    if (getenv("LSM_NO_FUSE") || res->w64) return false;
    if (p.kind != LSM_FILTERBANK_MEL || p.redundancy != 1 || (p.channels & 31) || p.channels > 256) return false;
    if (res->p.num_inputs != p.channels || res->p.num_steps != p.n_bins * p.n_thresholds) return false;
    if (res->n_pad != 4 * kThreads) return false;
    return lsm_res_smem_bytes(res->p.num_steps, (p.channels + 31) / 32, 4 * kThreads, res->p.num_neurons) <= 100 * 1024;
}

static void mel_fill(const lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes, double *d_spec_norm, MelArgs *out);

// Warp-per-frame arrangement unless LSM_MEL_BLOCK is set (the block-per-frame kernels of round 1, kept selectable) or the packed
// weights do not fit in shared memory.
static bool mel_use_warp(const lsm_frontend *fe)
{
    static const bool block_fft = getenv("LSM_MEL_BLOCK") != nullptr;
    return !block_fft && fe->grid_warp > 0;
}

static int mel_power_reserve(lsm_ctx *ctx, lsm_frontend *fe, int B)
{
    if (B <= fe->mel_power_cap) return LSM_OK;
    const int cap = (B + 255) & ~255;
    LSM_CUDA(ctx, cudaDeviceSynchronize());
    cudaFree(fe->d_mel_power);
    fe->d_mel_power = nullptr; fe->mel_power_cap = 0;
    LSM_CUDA(ctx, cudaMalloc((void **)&fe->d_mel_power, sizeof(float) * (size_t)cap * fe->ncols * fe->p.channels));
    fe->mel_power_cap = cap;
    return LSM_OK;
}

static int mel_power_launch(lsm_ctx *ctx, const MelArgs &a, cudaStream_t st)
{
    const long long warps = (long long)a.B * a.ncols;
    long long grid = (warps + kPW - 1) / kPW;
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    mel_power_kernel<<<(int)grid, kPT, kMelPowerSmemBase + sizeof(uint2) * 32 * (size_t)a.sched_len, st>>>(a);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return LSM_OK;
}

// audio -> features in one launch (mel): features as lsm_launch_reservoir writes them, spike trains optional
int lsm_launch_mel_fused(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B, uint8_t *d_spikes_or_null,
                         uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st, long long row0)
{
    if (B <= 0) return LSM_OK;
    MelArgs a;
    mel_fill(fe, d_pcm, B, d_spikes_or_null, nullptr, &a);
    lsm_reservoir_fill_args(res, nullptr, B, feature_mask, nan_to_num, d_features, nullptr, &a.res);
    a.res.gather_row0 += row0;
    size_t smem = lsm_res_smem_bytes(a.res.T, a.res.CW, 4 * kThreads, a.res.N);
    const bool warp = mel_use_warp(fe);
    int rc = warp ? mel_power_reserve(ctx, fe, B) : LSM_OK;
    if (rc != LSM_OK) return rc;
    a.power = fe->d_mel_power;
    if (!warp && smem < kMelSmemBytes) smem = kMelSmemBytes;
    if ((rc = lsm_frontend_order_before(ctx, fe, st)) != LSM_OK) return rc;
    int *counter = fe->d_counters + (fe->counter_next++ % 64);
    LSM_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(int), st));
    int per_sm = 0;
    if (warp) {
        if ((rc = mel_power_launch(ctx, a, st)) != LSM_OK) return rc;
        const bool narrow = a.C <= 128;                       // 128 threads x 8 neurons, else 256 x 4
        if (res->lean) rc = narrow ? mel_finish_fused_launch<1, 8>(ctx, a, counter, smem, B, st) : mel_finish_fused_launch<1, 4>(ctx, a, counter, smem, B, st);
        else rc = narrow ? mel_finish_fused_launch<2, 8>(ctx, a, counter, smem, B, st) : mel_finish_fused_launch<2, 4>(ctx, a, counter, smem, B, st);
        if (rc != LSM_OK) return rc;
        return lsm_frontend_order_after(ctx, fe, st);
    }
    if (res->lean) {
        LSM_CUDA(ctx, cudaFuncSetAttribute(mel_encode_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mel_encode_kernel<1>, kThreads, smem));
    } else {
        LSM_CUDA(ctx, cudaFuncSetAttribute(mel_encode_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LSM_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mel_encode_kernel<2>, kThreads, smem));
    }
    if (per_sm < 1) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "fused mel kernel does not fit on an SM");
    int grid = per_sm * ctx->sm_count;
    if (grid > fe->grid) grid = fe->grid;                    // the per-CTA dB planes were sized for the front end's own grid
    if (grid > B) grid = B;
    if (res->lean) mel_encode_kernel<1><<<grid, kThreads, smem, st>>>(a, counter);
    else mel_encode_kernel<2><<<grid, kThreads, smem, st>>>(a, counter);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st);
}

static void mel_fill(const lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes, double *d_spec_norm, MelArgs *out)
{
    const lsm_frontend_params &p = fe->p;
    MelArgs &a = *out;
    a.pcm = d_pcm; a.win = fe->d_window; a.tw = fe->d_twiddle; a.tw2 = fe->d_twiddle2;
    a.mel_w = fe->d_mel_w; a.mel_lo = fe->d_mel_lo; a.mel_n = fe->d_mel_n; a.mel_off = fe->d_mel_off;
    a.zoom_i0 = fe->d_zoom_i0; a.zoom_f = fe->d_zoom_f; a.scratch = fe->d_mel_scratch;
    a.spikes = d_spikes; a.spec_norm = d_spec_norm; a.power = fe->d_mel_power;
    a.sched = reinterpret_cast<const uint2 *>(fe->d_mel_sched); a.sched_len = fe->mel_sched_len;
    a.B = B; a.L = p.n_samples; a.C = p.channels; a.hop = p.mel_hop; a.ncols = fe->ncols; a.nbins = p.n_bins;
    a.K = p.n_thresholds; a.R = p.redundancy;
    for (int k = 0; k < 8; ++k) { a.thr[k] = p.thresholds_desc[k]; a.lower[k] = p.lower_bounds[k]; }
    for (int o = 0; o < 31; ++o) a.tw1[o] = make_double2(fe->mel_tw1[2 * o], fe->mel_tw1[2 * o + 1]);
    memset(&a.res, 0, sizeof(a.res));
}

int lsm_launch_mel(lsm_ctx *ctx, lsm_frontend *fe, const float *d_pcm, int B, uint8_t *d_spikes, double *d_spec_norm,
                   cudaStream_t st)
{
    if (B <= 0) return LSM_OK;
    MelArgs a;
    int rc = mel_use_warp(fe) ? mel_power_reserve(ctx, fe, B) : LSM_OK;
    if (rc != LSM_OK) return rc;
    mel_fill(fe, d_pcm, B, d_spikes, d_spec_norm, &a);
    if ((rc = lsm_frontend_order_before(ctx, fe, st)) != LSM_OK) return rc;
    int *counter = fe->d_counters + (fe->counter_next++ % 64);
    LSM_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(int), st));
    if (mel_use_warp(fe)) {
        if ((rc = mel_power_launch(ctx, a, st)) != LSM_OK) return rc;
        const int grid = B < fe->grid_warp ? B : fe->grid_warp;
        mel_finish_kernel<0, 4><<<grid, fe->finish_threads, 0, st>>>(a, counter);
    } else {
        const int grid = B < fe->grid ? B : fe->grid;
        mel_encode_kernel<0><<<grid, kThreads, kMelSmemBytes, st>>>(a, counter);
    }
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return lsm_frontend_order_after(ctx, fe, st);
}
