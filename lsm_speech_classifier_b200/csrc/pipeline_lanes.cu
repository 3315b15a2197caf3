// Audio -> features for the reference's default shape as ONE warp-specialised persistent kernel.
//
// Replaces the loop bodies of /root/reference/create_dataset.py:143-162 (gammatone branch) and
// extract_lsm_features.py:78-87 for a batch of utterances; same results as gammatone_encode_kernel (frontend_gammatone.cu),
// which stays the kernel for every other shape and for zero-copy host buffers.
//
// Two resources bound this path on a B200: the fp64 pipe (13 DFMA per channel-sample for the filter bank) and the warp
// schedulers' issue slots (the event-driven reservoir).  A CTA therefore carries both kinds of work side by side:
//
//   warps 0-7   FILTER   lane = utterance, warp = one channel of a unit (32 utterances x 8 channels; LSM_PIPE_J = 2: four warps of two).  The coefficients are the
//                        same for all lanes of a warp and reach the DFMAs through uniform registers / the constant bank, which
//                        is what lets a DFMA issue at the pipe's full rate (with three register operands it is capped at 75 %).
//                        PCM arrives by TMA: one cp.async.bulk.tensor.2d per 32-utterance x 32-sample tile, 128-byte hardware
//                        swizzle so that lane = row reads are conflict-free, eight tiles in the ring, full / empty mbarriers;
//                        all four warps of the CTA read the same tiles, so no filter warp ever waits on global memory and the
//                        warps never meet at a CTA barrier.  Output: raw window energy sums to the energy planes (L2), then
//                        one release-increment of the unit's group counter.
//   the rest    ENCODER + RESERVOIR + READOUT, two units of 128 threads.  A unit takes the next utterance whose group is complete
//                        (acquire-poll of the group counter), runs the speculative encoder epilogue with the derived error bound
//                        (gammatone_core.cuh), hands the spikes over as bits in shared memory, simulates the reservoir and
//                        writes the feature row (reservoir_core.cuh).  Utterances the bound cannot settle go to a work list
//                        that the host follows up with the exact kernel.
//
// Filter warps never wait for the other role, so every resident CTA makes progress; units are dealt statically (blockIdx +
// k * gridDim), which keeps the channel index provably uniform for the compiler.  A launch has one CTA per SM; two launches
// (the two lanes of a ctx) share an SM, so the fill and drain phases of one overlap the steady state of the other.
#include <stdlib.h>

#include <memory>
#include <new>

#include <cuda.h>      // CUtensorMap and its enums (the encode function itself is fetched through the runtime)

#include "gammatone_core.cuh"

namespace {

#ifndef LSM_PIPE_J
#define LSM_PIPE_J 1
#endif
constexpr int kJ = LSM_PIPE_J;            // channels per filter warp (a unit = 32 utterances x kFW * kJ = 8 channels)
constexpr int kFW = 8 / kJ;               // filter warps per CTA
constexpr int kRegsFilter = 56, kRegsUnit = 72;   // kJ = 1: registers per thread after setmaxnreg (256 x 56 + 256 x 72 = 512 x 64)
constexpr int kEU = 2;                    // encoder / reservoir units per CTA
constexpr int kEThreads = 128;            // threads per unit = channels (one thread per channel), 8 neurons per thread
constexpr int kThreads = kFW * 32 + kEU * kEThreads;
constexpr int kUPGShift = 4;              // units per 32-utterance group = 128 channels / 8 = 16 (shifts keep the unit -> channel
constexpr int kUPG = 1 << kUPGShift;      // arithmetic on the uniform datapath; a division would move it to vector registers)
constexpr int kChunk = 32;                // samples per staged tile (x 32 utterances)
constexpr int kRawStages = 8;             // tiles in the ring
constexpr int kLookAhead = 6;             // tiles requested ahead of the consumers (the other two are being read)


struct PipeArgs {
    GtArgs gt;                // front end + reservoir (gt.res); gt.pcm / gt.pcm16 = the batch
    double *energy;           // [B][ncols][C] raw window energy sums
    float *xmax;              // [B] max |sample|
    int *done;                // [groups] units completed per 32-utterance group (zero at launch)
    int *unit_next;           // encoder units' utterance counter (zero at launch)
    int *err_flag;            // set if a unit gave up waiting (should never happen)
    int n_units;              // filter units in this launch (kUPG per 32-utterance group)
    int n_chunks, n_groups8;  // chunks per unit, last 8-sample group index
    int debug;                // LSM_PIPE_DEBUG: 1 = filter warps only, 2 = encoder/reservoir units only (timing experiments)
    double coef[128][6];      // per channel: c1..c4 (numerator zeros / A0), -a1, -a2
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok = 0, tries = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++tries > (1u << 24)) __trap();       // a copy that never lands: abort the launch instead of hanging the GPU
    }
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA: one 2-D tile (x = first sample, y = first utterance) of the [B][L] PCM tensor into shared memory, swizzled by the
// hardware, completion counted in bytes on an mbarrier; rows / samples outside the tensor arrive as zeros
__device__ __forceinline__ void tma_tile_g2s(void *dst, const CUtensorMap *map, int x, int y, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <int BAR>
__device__ __forceinline__ bool group_or(bool v, int nthr)
{
    unsigned r;
    asm volatile("{ .reg .pred p, q; setp.ne.u32 q, %1, 0; bar.red.or.pred p, %2, %3, q; selp.u32 %0, 1, 0, p; }"
                 : "=r"(r) : "r"((unsigned)v), "r"(BAR), "r"(nthr) : "memory");
    return r != 0;
}

struct FilterSmem {
    unsigned char *raw;               // [kRawStages][32 rows][32 samples] PCM tiles as the TMA engine lays them out (swizzled)
    unsigned long long *full;         // [kRawStages] mbarriers: tile landed
    unsigned long long *empty;        // [kRawStages] mbarriers: all kFW warps have read the tile
};

// ---- FILTER role, warp W of the filter group: channels 8 * (unit's channel block) + 2 W, + 1 for the unit's 32 utterances.
// Everything that selects the channels (W, u, cb, ch) is built from a warp-uniform W (REDUX result), blockIdx, gridDim and loop
// counters with adds, shifts and masks only, so that it stays on the uniform datapath and the coefficients reach the
// DFMAs as uniform registers.
template <bool I16>
__device__ __forceinline__ void filter_role(const PipeArgs &a, const CUtensorMap *tmap, const FilterSmem &sm, const int W)
{
    const int lane = threadIdx.x & 31;
    const int hop = a.gt.hop, ncols = a.gt.ncols, C = a.gt.C;
    const int r_old = a.gt.nwin - 2 * hop;
    constexpr int kTileBytes = 32 * kChunk * (I16 ? 2 : 4);
    // this lane's row inside a tile, with the hardware swizzle of its 16-byte chunks folded in: 128-byte rows XOR the chunk
    // index with (row & 7), 64-byte rows with ((row >> 1) & 3)
    const unsigned row_off = (unsigned)lane * (I16 ? 64u : 128u);
    const unsigned sw = I16 ? (((unsigned)lane >> 1) & 3u) : ((unsigned)lane & 7u);

    // producer cursor (lane 0 of warp 0): the next tile to request, kLookAhead tiles ahead of the consumers, across unit boundaries
    int pu = blockIdx.x, pc = 0;
    unsigned pseq = 0;
#define LSM_PIPE_ISSUE()                                                                                          \
    if (W == 0 && pu < a.n_units) {                                                                               \
        const unsigned st_ = pseq & (kRawStages - 1);                                                             \
        if (pseq >= kRawStages) mbar_wait(sm.empty + st_, ((pseq >> 3) + 1u) & 1u);   /* previous tenant read by all */ \
        if (lane == 0) {                                                                                          \
            mbar_expect_tx(sm.full + st_, kTileBytes);                                                            \
            tma_tile_g2s(sm.raw + (size_t)st_ * kTileBytes, tmap, pc * kChunk, (pu >> kUPGShift) * 32, sm.full + st_); \
        }                                                                                                         \
        ++pseq;                                                                                                   \
        if (++pc == a.n_chunks) { pc = 0; pu += gridDim.x; }                                                      \
    }
#pragma unroll 1
    for (int s = 0; s < kLookAhead; ++s) { LSM_PIPE_ISSUE() }

    unsigned seq = 0;
#pragma unroll 1
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        const int g = u >> kUPGShift, cb = u & (kUPG - 1);
        const int ch = cb * (kFW * kJ) + W * kJ;
        // the twelve coefficients of the two channels: loaded once per unit, uniform-register operands of every DFMA below
        const double a0 = a.coef[ch][0], a1 = a.coef[ch][1], a2 = a.coef[ch][2], a3 = a.coef[ch][3], a4 = a.coef[ch][4], a5 = a.coef[ch][5];
        const int chb = kJ == 2 ? ch + 1 : ch;
        const double b0 = a.coef[chb][0], b1 = a.coef[chb][1], b2 = a.coef[chb][2], b3 = a.coef[chb][3],
                     b4 = a.coef[chb][4], b5 = a.coef[chb][5];
        const int utt = g * 32 + lane;
        const bool valid = utt < a.gt.B;
        double *dst = a.energy + (size_t)(valid ? utt : 0) * ncols * C + ch;

        // channel A / channel B state: stage k's previous two outputs (p, q), running energy sum and the two previous block sums
        double xp = 0.0;
        double ap1 = 0.0, aq1 = 0.0, ap2 = 0.0, aq2 = 0.0, ap3 = 0.0, aq3 = 0.0, ap4 = 0.0, aq4 = 0.0, aacc = 0.0, af1 = 0.0, af2 = 0.0;
        double bp1 = 0.0, bq1 = 0.0, bp2 = 0.0, bq2 = 0.0, bp3 = 0.0, bq3 = 0.0, bp4 = 0.0, bq4 = 0.0, bacc = 0.0, bf1 = 0.0, bf2 = 0.0;
        int next_b = r_old, m = 0, gi = 0;
        bool head = true;

        // one sample through both channels: stage k works one iteration behind stage k-1 (software skew: four independent
        // 3-FMA chains per channel), so iteration i completes output i - 3
#define LSM_PIPE_SAMPLE(xv)                                                                     \
        {                                                                                       \
            const double x_ = (xv);                                                             \
            const double ya1 = fma(a4, ap1, fma(a5, aq1, fma(a0, xp, x_)));                     \
            const double ya2 = fma(a4, ap2, fma(a5, aq2, fma(a1, aq1, ap1)));                   \
            const double ya3 = fma(a4, ap3, fma(a5, aq3, fma(a2, aq2, ap2)));                   \
            const double ya4 = fma(a4, ap4, fma(a5, aq4, fma(a3, aq3, ap3)));                   \
            if (kJ == 2) {                                                                      \
                const double yb1 = fma(b4, bp1, fma(b5, bq1, fma(b0, xp, x_)));                 \
                const double yb2 = fma(b4, bp2, fma(b5, bq2, fma(b1, bq1, bp1)));               \
                const double yb3 = fma(b4, bp3, fma(b5, bq3, fma(b2, bq2, bp2)));               \
                const double yb4 = fma(b4, bp4, fma(b5, bq4, fma(b3, bq3, bp3)));               \
                bq1 = bp1; bp1 = yb1; bq2 = bp2; bp2 = yb2; bq3 = bp3; bp3 = yb3; bq4 = bp4; bp4 = yb4; \
                bacc = fma(yb4, yb4, bacc);                                                     \
            }                                                                                   \
            xp = x_;                                                                            \
            aq1 = ap1; ap1 = ya1; aq2 = ap2; ap2 = ya2; aq3 = ap3; ap3 = ya3; aq4 = ap4; ap4 = ya4; \
            aacc = fma(ya4, ya4, aacc);                                                         \
        }
        // the window-phase boundary at output sample 8 gi: reached after the third iteration of group gi
#define LSM_PIPE_BOUNDARY()                                                                                     \
        if (8 * gi == next_b) {                                                                                 \
            if (head) {                                                                                         \
                if (m >= 2 && valid) {                                /* window m-2 = full(m-2) + full(m-1) + head(m) */ \
                    if (kJ == 2) __stcg(reinterpret_cast<double2 *>(dst + (size_t)(m - 2) * C), make_double2((af2 + af1) + aacc, (bf2 + bf1) + bacc)); \
                    else __stcg(dst + (size_t)(m - 2) * C, (af2 + af1) + aacc);                                 \
                }                                                                                               \
                next_b += hop - r_old;                                                                          \
                head = false;                                                                                   \
            } else {                                                                                            \
                af2 = af1; af1 = aacc; aacc = 0.0;                                                              \
                bf2 = bf1; bf1 = bacc; bacc = 0.0;                                                              \
                next_b += r_old;                                                                                \
                head = true;                                                                                    \
                ++m;                                                                                            \
            }                                                                                                   \
        }
        // Loop body = samples 3..7 of the previous group (carried in registers; zeros before the first group, which leave the
        // zero state untouched), samples 0..2 of this one, then the boundary: the branch sits at the end of a straight-line
        // body of 208 DFMAs that stays in the instruction cache (the group loop is deliberately not unrolled).
        float c3 = 0.0f, c4 = 0.0f, c5 = 0.0f, c6 = 0.0f, c7 = 0.0f;
#define LSM_PIPE_GROUP(x0, x1, x2, x3, x4, x5, x6, x7)                                                          \
        if (gi <= a.n_groups8) {                                                                                \
            LSM_PIPE_SAMPLE((double)c3) LSM_PIPE_SAMPLE((double)c4) LSM_PIPE_SAMPLE((double)c5)                 \
            LSM_PIPE_SAMPLE((double)c6) LSM_PIPE_SAMPLE((double)c7)                                             \
            LSM_PIPE_SAMPLE((double)(x0)) LSM_PIPE_SAMPLE((double)(x1)) LSM_PIPE_SAMPLE((double)(x2))           \
            LSM_PIPE_BOUNDARY()                                                                                 \
            c3 = (x3); c4 = (x4); c5 = (x5); c6 = (x6); c7 = (x7);                                              \
            ++gi;                                                                                               \
        }

#pragma unroll 1
        for (int c = 0; c < a.n_chunks; ++c, ++seq) {
            LSM_PIPE_ISSUE()
            const unsigned stage = seq & (kRawStages - 1);
            mbar_wait(sm.full + stage, (seq >> 3) & 1u);
            const unsigned char *rp = sm.raw + (size_t)stage * kTileBytes + row_off;
            if (I16) {
                // 64-byte rows: four 16-byte chunks of 8 samples each; (float)int16 * 2^-15 is exact, as is its double
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(rp + (((unsigned)k ^ sw) << 4));
                    if (k == 3) { __syncwarp(); if (lane == 0) mbar_arrive(sm.empty + stage); }      // the tile is in registers
                    LSM_PIPE_GROUP((float)(short)(v.x & 0xffff) * 0x1p-15f, (float)(short)(v.x >> 16) * 0x1p-15f,
                                   (float)(short)(v.y & 0xffff) * 0x1p-15f, (float)(short)(v.y >> 16) * 0x1p-15f,
                                   (float)(short)(v.z & 0xffff) * 0x1p-15f, (float)(short)(v.z >> 16) * 0x1p-15f,
                                   (float)(short)(v.w & 0xffff) * 0x1p-15f, (float)(short)(v.w >> 16) * 0x1p-15f)
                }
            } else {
                // 128-byte rows: eight 16-byte chunks of 4 samples each
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const float4 f0 = *reinterpret_cast<const float4 *>(rp + (((unsigned)(2 * k) ^ sw) << 4));
                    const float4 f1 = *reinterpret_cast<const float4 *>(rp + (((unsigned)(2 * k + 1) ^ sw) << 4));
                    if (k == 3) { __syncwarp(); if (lane == 0) mbar_arrive(sm.empty + stage); }      // the tile is in registers
                    LSM_PIPE_GROUP(f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w)
                }
            }
        }
#undef LSM_PIPE_GROUP
#undef LSM_PIPE_BOUNDARY
#undef LSM_PIPE_SAMPLE
        // publish the unit: each lane wrote its own utterance's sums; fence, then one release-increment per warp
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(a.done + g, 1);
    }
#undef LSM_PIPE_ISSUE
}

// Peak level of every utterance (the error bound scales with it): one warp per utterance, 16-byte loads.
template <bool I16>
__global__ void __launch_bounds__(256) peak_kernel(const float *pcm, const int16_t *pcm16, int B, int L, float *xmax)
{
    const int utt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (utt >= B) return;
    const int lane = threadIdx.x & 31;
    float m = 0.0f;
    if (I16) {
        const uint4 *src = reinterpret_cast<const uint4 *>(pcm16 + (size_t)utt * L);
        for (int i = lane; i < L / 8; i += 32) {
            const uint4 v = __ldg(src + i);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                m = fmaxf(m, fabsf((float)(short)(w[k] & 0xffff)));
                m = fmaxf(m, fabsf((float)(short)(w[k] >> 16)));
            }
        }
        m *= 0x1p-15f;
    } else {
        const float4 *src = reinterpret_cast<const float4 *>(pcm + (size_t)utt * L);
        for (int i = lane; i < L / 4; i += 32) {
            const float4 v = __ldg(src + i);
            m = fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fmaxf(fabsf(v.y), fabsf(v.z)), fabsf(v.w)));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) xmax[utt] = m;
}

// ---- ENCODER + RESERVOIR role: unit e (threads 0..127 of the unit, named barrier 1 + e)
template <int BAR>
__device__ __forceinline__ void unit_role(const PipeArgs &a, unsigned char *usmem, double *s_red, double *s_out, int *s_cnt,
                                          int *s_utt, const int tid)
{
    const GtArgs &gt = a.gt;
    for (;;) {
        if (tid == 0) {
            const int i = atomicAdd(a.unit_next, 1);
            int ok = i < gt.B ? i : -1;
            if (ok >= 0) {
                // wait until the filter warps (of any CTA) have completed every unit of this utterance's group
                const int *flag = a.done + (i >> 5);
                long long spins = 0;
                while (ld_acquire(flag) < kUPG * kFW) {          // every warp of every unit of the group counts once
                    __nanosleep(1000);
                    if (++spins > (1ll << 24)) { atomicExch(a.err_flag, 1); ok = -1; break; }     // several seconds: give up loudly
                }
            }
            *s_utt = ok;
        }
        res_sync<BAR>(kEThreads);
        const int utt = *s_utt;
        if (utt < 0) break;
        const float xm = __ldcg(a.xmax + utt);
        double *plane = a.energy + (size_t)utt * gt.ncols * gt.C;
        const bool near = spec_epilogue<8, BAR>(gt, utt, plane, tid, kEThreads, xm, s_red, s_out, usmem);
        // too close to call on the speculative plane: finish it like the others, and list it for the exact pass
        if (group_or<BAR>(near, kEThreads) && tid == 0) {
            gt.rerun_list[1 + atomicAdd(gt.rerun_list, 1)] = utt;
            atomicAdd(gt.reruns, 1);
        }
        reservoir_simulate<8, true, false, BAR>(gt.res, utt, usmem, s_cnt, tid, kEThreads);
        res_sync<BAR>(kEThreads);                       // shared memory and *s_utt are reused by the next utterance
    }
}

template <bool I16>
__global__ void __launch_bounds__(kThreads, 2) pipeline_kernel(const __grid_constant__ PipeArgs a, const __grid_constant__ CUtensorMap tmap,
                                                               const int unit_smem_bytes)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ double s_red[kEU][6 * 8];
    __shared__ double s_out[kEU][6];
    __shared__ int s_cnt[kEU][5];
    __shared__ int s_utt[kEU];
    __shared__ __align__(8) unsigned long long s_full[kRawStages], s_empty[kRawStages];

    // the warp index as a value the compiler knows to be warp-uniform (REDUX writes a uniform register): the filter warps'
    // channel index, and with it their coefficients, stay on the uniform datapath
    const int warp = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    FilterSmem fs;
    fs.raw = smem + (size_t)kEU * unit_smem_bytes;              // 1024-byte aligned: the swizzle pattern is relative to it
    fs.full = s_full;
    fs.empty = s_empty;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kRawStages; ++s) { mbar_init(s_full + s, 1); mbar_init(s_empty + s, kFW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < kFW) {
        if (kJ == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsFilter));
        if (a.debug == 2) {          // experiment: no filtering, the planes of an earlier launch stand in
            if (threadIdx.x == 0) for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) atomicAdd(a.done + (u >> kUPGShift), kFW);
            return;
        }
        filter_role<I16>(a, &tmap, fs, warp);
    } else {
        if (a.debug == 1) return;    // experiment: filter warps alone
        if (kJ == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsUnit));
        const int e = (warp - kFW) >> 2;
        const int tid = threadIdx.x - kFW * 32 - e * kEThreads;
        unsigned char *usmem = smem + (size_t)e * unit_smem_bytes;
        if (e == 0) unit_role<1>(a, usmem, s_red[0], s_out[0], s_cnt[0], &s_utt[0], tid);
        else unit_role<2>(a, usmem, s_red[1], s_out[1], s_cnt[1], &s_utt[1], tid);
    }
}


// ---- Second form of the same idea (LSM_PIPELINE=2): the filter role is the lane = utterance unit of gammatone_energy_kernel
//      (energy_unit: 8-sample groups fully expanded, coefficients resident in uniform registers, PCM by 16-byte loads through L1 -
//      0.17 other instructions per DFMA where the TMA-fed role above executes 0.97), units of two channels x 32 utterances handed
//      out by an atomic counter (single-channel units at the end of the queue balance the tail), and the same encoder / reservoir
//      units.  4 filter warps + 2 units per CTA, two CTAs per SM, so that consecutive launches overlap at half-SM granularity.
constexpr int kP2FW = 4, kP2EU = 2, kP2JB = 2;
constexpr int kP2Threads = kP2FW * 32 + kP2EU * kEThreads;

struct Pipe2Args {
    GtArgs gt;                // front end + reservoir (gt.res)
    EnergyArgs ea;            // the filter units' view: PCM, energy planes, peak levels, coefficient table
    int *done;                // [groups] channels completed per 32-utterance group (zero at launch)
    int *unit_next;           // encoder units' utterance counter (zero at launch)
    int *filter_next;         // filter warps' unit counter (zero at launch)
    int *err_flag;
    int n_units, n_big, big_groups;
};

template <int BAR>
__device__ __forceinline__ void unit2_role(const Pipe2Args &a, unsigned char *usmem, double *s_red, double *s_out, int *s_cnt,
                                           int *s_utt, const int tid)
{
    const GtArgs &gt = a.gt;
    for (;;) {
        if (tid == 0) {
            const int i = atomicAdd(a.unit_next, 1);
            int ok = i < gt.B ? i : -1;
            if (ok >= 0) {
                const int *flag = a.done + (i >> 5);
                long long spins = 0;
                while (ld_acquire(flag) < gt.C) {                 // every channel of the utterance's group has been filtered
                    __nanosleep(500);
                    if (++spins > (1ll << 24)) { atomicExch(a.err_flag, 1); ok = -1; break; }
                }
            }
            *s_utt = ok;
        }
        res_sync<BAR>(kEThreads);
        const int utt = *s_utt;
        if (utt < 0) break;
        const float xm = __ldcg(a.ea.xmax + utt);
        double *plane = a.ea.energy + (size_t)utt * gt.ncols * gt.C;
        const bool near = spec_epilogue<8, BAR>(gt, utt, plane, tid, kEThreads, xm, s_red, s_out, usmem);
        if (group_or<BAR>(near, kEThreads) && tid == 0) {
            gt.rerun_list[1 + atomicAdd(gt.rerun_list, 1)] = utt;
            atomicAdd(gt.reruns, 1);
        }
        reservoir_simulate<8, true, false, BAR>(gt.res, utt, usmem, s_cnt, tid, kEThreads);
        res_sync<BAR>(kEThreads);
    }
}

__global__ void __launch_bounds__(kP2Threads, 2) pipeline2_kernel(const __grid_constant__ Pipe2Args a, const int unit_smem_bytes)
{
    extern __shared__ __align__(1024) unsigned char smem[];      // one declaration per translation unit: pipeline_kernel needs 1024
    __shared__ double s_red[kP2EU][6 * 8];
    __shared__ double s_out[kP2EU][6];
    __shared__ int s_cnt[kP2EU][5];
    __shared__ int s_utt[kP2EU];

    const int warp = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    if (warp < kP2FW) {
        const int lane = threadIdx.x & 31;
        for (;;) {
            int u = 0;
            if (lane == 0) u = atomicAdd(a.filter_next, 1);
            // the unit index as a value the compiler knows to be warp-uniform: channel index and coefficients stay on the uniform datapath
            u = __reduce_max_sync(0xffffffffu, __shfl_sync(0xffffffffu, u, 0));
            if (u >= a.n_units) break;
            int g, add;
            if (u < a.n_big) {                                  // 64 two-channel units per group
                g = u >> 6;
                energy_unit<kP2JB>(a.ea, g, (u & 63) << 1);
                add = kP2JB;
            } else {                                            // 128 single-channel units per group
                const int s = u - a.n_big;
                g = a.big_groups + (s >> 7);
                energy_unit<1>(a.ea, g, s & 127);
                add = 1;
            }
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(a.done + g, add);
        }
    } else {
        const int e = (warp - kP2FW) >> 2;
        const int tid = threadIdx.x - kP2FW * 32 - e * kEThreads;
        unsigned char *usmem = smem + (size_t)e * unit_smem_bytes;
        if (e == 0) unit2_role<1>(a, usmem, s_red[0], s_out[0], s_cnt[0], &s_utt[0], tid);
        else unit2_role<2>(a, usmem, s_red[1], s_out[1], s_cnt[1], &s_utt[1], tid);
    }
}

}  // namespace

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*lsm_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static lsm_encode_tiled_fn encode_tiled()
{
    static lsm_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (lsm_encode_tiled_fn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// Can this pair run as the warp-specialised kernel?  The reference's default shape: 128 gammatone channels, no redundancy, a lean
// reservoir of <= 1024 neurons (8 per thread), window phases in whole 8-sample groups.
bool lsm_pipeline_lanes_eligible(const lsm_frontend *fe, const lsm_reservoir *res, const void *d_pcm, bool i16)
{
    const lsm_frontend_params &p = fe->p;
    // Opt-in (LSM_PIPELINE=1): measured on B200 at 7.1 ms per 2400 utterances against 5.8 ms for the lane = channel fused kernel
    // (the two roles together issue 4.4 G warp instructions per step, the lane = channel kernel 3.5 G).
    if (!getenv("LSM_PIPELINE") || !d_pcm) return false;
    if (lsm_fused_npt(fe, res) != 8 || p.channels != 128 || !res->lean || res->n_pad != 1024) return false;
    if (fe->mode != LSM_FILTER_SPECULATIVE) return false;
    const int r_old = p.nwin - 2 * p.hop;
    if (p.hop % 8 || r_old % 8 || r_old <= 0) return false;
    if (p.n_samples % (i16 ? 8 : 4) || (((uintptr_t)d_pcm) & 15)) return false;
    if (p.n_samples < kChunk) return false;
    return encode_tiled() != nullptr;
}


// LSM_PIPELINE=2: the energy-unit form (float32 PCM in device memory only)
static int launch_pipeline2(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B, uint8_t *d_spikes_or_null,
                            uint32_t feature_mask, int nan_to_num, double *d_features, cudaStream_t st, int lane, long long row0)
{
    const lsm_frontend_params &p = fe->p;
    int rc;
    if ((rc = lsm_frontend_ensure_energy(ctx, fe, B)) != LSM_OK) return rc;
    if ((rc = lsm_frontend_order_before(ctx, fe, st, lane)) != LSM_OK) return rc;
    std::unique_ptr<Pipe2Args> holder(new (std::nothrow) Pipe2Args);
    if (!holder) LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
    Pipe2Args &a = *holder;
    lsm_gammatone_fill_args(fe, d_pcm, B, d_spikes_or_null, nullptr, &a.gt);
    lsm_reservoir_fill_args(res, nullptr, B, feature_mask, nan_to_num, d_features, nullptr, &a.gt.res);
    a.gt.res.gather_row0 += row0;
    const int groups = (B + 31) / 32;
    const size_t plane = (size_t)fe->ncols * p.channels;
    EnergyArgs &ea = a.ea;
    ea.pcm = d_pcm; ea.energy = fe->d_energy + (size_t)lane * fe->energy_cap * plane; ea.xmax = fe->d_xmax + (size_t)lane * fe->energy_cap;
    ea.B = B; ea.L = p.n_samples; ea.C = p.channels; ea.nwin = p.nwin; ea.hop = p.hop; ea.ncols = fe->ncols;
    memcpy(ea.coef, fe->h_lane_coef, sizeof(double) * 6 * p.channels);
    int *sync = fe->d_pipe_sync + (size_t)lane * (fe->energy_cap / 32 + 8);
    a.unit_next = sync; a.filter_next = sync + 1; a.done = sync + 4;
    a.err_flag = fe->d_pipe_sync + 2 * (size_t)(fe->energy_cap / 32 + 8);
    a.gt.rerun_list = fe->d_rerun + (size_t)lane * (fe->rerun_cap + 1);
    int grid = 2 * ctx->sm_count;
    // whole rounds of two-channel units over all filter warps, the rest of the batch as single-channel units (they balance the tail)
    const int upg = p.channels / kP2JB, workers = grid * kP2FW;
    const long long big_total = (long long)groups * upg;
    long long rounds = big_total / workers;
    if (rounds * workers == big_total && rounds > 0) rounds -= 0;          // an exact fit needs no small units
    int big_groups = (int)((rounds * workers) / upg);
    if (big_groups > groups) big_groups = groups;
    a.big_groups = big_groups;
    a.n_big = big_groups * upg;
    a.n_units = a.n_big + (groups - big_groups) * p.channels;
    ea.big_groups = big_groups; ea.n_units = a.n_units;
    const int unit_smem = (int)((lsm_res_smem_bytes(a.gt.res.T, a.gt.res.CW, kEThreads * 8, a.gt.res.N) + 127) & ~(size_t)127);
    const size_t smem = (size_t)kP2EU * unit_smem;
    LSM_CUDA(ctx, cudaMemsetAsync(sync, 0, sizeof(int) * (4 + (size_t)groups), st));
    LSM_CUDA(ctx, cudaMemsetAsync(a.gt.rerun_list, 0, sizeof(int), st));
    LSM_CUDA(ctx, cudaFuncSetAttribute(pipeline2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (grid * kP2FW > a.n_units) grid = (a.n_units + kP2FW - 1) / kP2FW;
    pipeline2_kernel<<<grid, kP2Threads, smem, st>>>(a, unit_smem);
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    GtArgs x = a.gt;
    x.mode = 0; x.rerun_list = nullptr;
    x.utt_list = a.gt.rerun_list + 1; x.utt_count = a.gt.rerun_list;
    return lsm_launch_fused_args(ctx, fe, res, x, st, true, nullptr, 8, lane);
}

// One launch: utterances [0, B) of d_pcm (float32, or PCM16 when fe->next_pcm16 is set and d_pcm is null) on launch lane `lane`
// of the front end (its own energy planes, counters and work list), stream st.  Features (and optionally spike trains) as
// lsm_launch_fused; row0 offsets the fused all-gather's destination rows.  The exact pass over the flagged utterances follows
// on the same stream.
int lsm_launch_pipeline_lanes(lsm_ctx *ctx, lsm_frontend *fe, lsm_reservoir *res, const float *d_pcm, int B,
                              uint8_t *d_spikes_or_null, uint32_t feature_mask, int nan_to_num, double *d_features,
                              cudaStream_t st, int lane, long long row0)
{
    if (B <= 0) return LSM_OK;
    const lsm_frontend_params &p = fe->p;
    const bool i16 = d_pcm == nullptr;
    int rc;
    if (!i16 && getenv("LSM_PIPELINE") && atoi(getenv("LSM_PIPELINE")) == 2)
        return launch_pipeline2(ctx, fe, res, d_pcm, B, d_spikes_or_null, feature_mask, nan_to_num, d_features, st, lane, row0);
    if ((rc = lsm_frontend_ensure_energy(ctx, fe, B)) != LSM_OK) return rc;
    if ((rc = lsm_frontend_order_before(ctx, fe, st, lane)) != LSM_OK) return rc;      // earlier users of this lane's planes
    std::unique_ptr<PipeArgs> holder(new (std::nothrow) PipeArgs);
    if (!holder) LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
    PipeArgs &a = *holder;
    lsm_gammatone_fill_args(fe, d_pcm, B, d_spikes_or_null, nullptr, &a.gt);
    lsm_reservoir_fill_args(res, nullptr, B, feature_mask, nan_to_num, d_features, nullptr, &a.gt.res);
    a.gt.res.gather_row0 += row0;
    const int groups = (B + 31) / 32;
    const size_t plane = (size_t)fe->ncols * p.channels;
    a.energy = fe->d_energy + (size_t)lane * fe->energy_cap * plane;
    a.xmax = fe->d_xmax + (size_t)lane * fe->energy_cap;
    int *sync = fe->d_pipe_sync + (size_t)lane * (fe->energy_cap / 32 + 8);
    a.unit_next = sync; a.err_flag = fe->d_pipe_sync + 2 * (size_t)(fe->energy_cap / 32 + 8); a.done = sync + 4;
    a.gt.rerun_list = fe->d_rerun + (size_t)lane * (fe->rerun_cap + 1);
    a.n_units = groups * kUPG;
    { const char *e = getenv("LSM_PIPE_DEBUG"); a.debug = e ? atoi(e) : 0; }
    const int n_used = (fe->ncols - 1) * p.hop + p.nwin;
    a.n_groups8 = n_used / 8;
    a.n_chunks = (8 * (a.n_groups8 + 1) + kChunk - 1) / kChunk;
    memcpy(a.coef, fe->h_lane_coef, sizeof(double) * 6 * 128);
    const int unit_smem = (int)((lsm_res_smem_bytes(a.gt.res.T, a.gt.res.CW, kEThreads * 8, a.gt.res.N) + 127) & ~(size_t)127);
    const int unit_smem_al = (kEU * unit_smem + 1023) / 1024 * 1024 / kEU;          // the tile ring behind the units starts 1024-aligned
    const size_t smem = (size_t)kEU * unit_smem_al + (size_t)kRawStages * 32 * kChunk * (i16 ? 2 : 4);
    if (smem > 113 * 1024) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "pipeline kernel: %zu bytes of shared memory per CTA (two per SM must fit)", smem);
    LSM_CUDA(ctx, cudaMemsetAsync(sync, 0, sizeof(int) * (4 + (size_t)groups), st));
    LSM_CUDA(ctx, cudaMemsetAsync(a.gt.rerun_list, 0, sizeof(int), st));
    // the batch as a 2-D tensor [B][L] for the TMA engine: tiles of 32 utterances x 32 samples, swizzled so that lane = row reads
    // of 16 bytes are conflict-free (128-byte rows: SWIZZLE_128B; PCM16's 64-byte rows: SWIZZLE_64B); out-of-range rows and
    // samples arrive as zeros
    alignas(64) CUtensorMap tmap;
    {
        const void *base = i16 ? (const void *)a.gt.pcm16 : (const void *)a.gt.pcm;
        const cuuint64_t dims[2] = {(cuuint64_t)p.n_samples, (cuuint64_t)B};
        const cuuint64_t strides[1] = {(cuuint64_t)p.n_samples * (i16 ? 2 : 4)};
        const cuuint32_t box[2] = {kChunk, 32}, estr[2] = {1, 1};
        const CUresult r = encode_tiled()(&tmap, i16 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                          const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          i16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) LSM_FAIL(ctx, LSM_ERR_CUDA, "cuTensorMapEncodeTiled -> %d", (int)r);
    }
    int grid = ctx->sm_count;
    if (const char *e = getenv("LSM_PIPE_GRID_MULT")) grid *= atoi(e) > 0 ? atoi(e) : 1;      // experiment: CTAs per SM of one launch
    if (grid > a.n_units) grid = a.n_units;
    const int pk_blocks = (B + 7) / 8;
    if (i16) {
        peak_kernel<true><<<pk_blocks, 256, 0, st>>>(nullptr, a.gt.pcm16, B, p.n_samples, a.xmax);
        LSM_CUDA(ctx, cudaFuncSetAttribute(pipeline_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pipeline_kernel<true><<<grid, kThreads, smem, st>>>(a, tmap, unit_smem_al);
    } else {
        peak_kernel<false><<<pk_blocks, 256, 0, st>>>(a.gt.pcm, nullptr, B, p.n_samples, a.xmax);
        LSM_CUDA(ctx, cudaFuncSetAttribute(pipeline_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pipeline_kernel<false><<<grid, kThreads, smem, st>>>(a, tmap, unit_smem_al);
    }
    ctx->launches += 2;
    LSM_CUDA(ctx, cudaGetLastError());
    // exact pass: the lane = channel kernel in exact mode over the device work list (typically empty; a few CTAs suffice)
    GtArgs x = a.gt;
    x.mode = 0; x.rerun_list = nullptr;
    x.utt_list = a.gt.rerun_list + 1; x.utt_count = a.gt.rerun_list;
    return lsm_launch_fused_args(ctx, fe, res, x, st, true, nullptr, 8, lane);
}
