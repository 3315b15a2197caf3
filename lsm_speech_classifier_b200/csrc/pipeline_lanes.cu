// Audio -> features for the reference's default shape as ONE warp-specialised persistent kernel.
//
// Replaces the loop bodies of /root/reference/create_dataset.py:143-162 (gammatone branch) and
// extract_lsm_features.py:78-87 for a batch of utterances; same results as gammatone_encode_kernel (frontend_gammatone.cu),
// which stays the kernel for every other shape and for zero-copy host buffers.
//
// Two resources bound this path on a B200: the fp64 pipe (13 DFMA per channel-sample for the filter bank) and the warp
// schedulers' issue slots (the event-driven reservoir).  A CTA therefore carries both kinds of work side by side:
//
//   warps 0-7   FILTER   lane = utterance, warp = one channel of a unit (32 utterances x 8 channels).  The coefficients are the
//                        same for all lanes of a warp and reach the DFMAs through uniform registers / the constant bank, which
//                        is what lets a DFMA issue at the pipe's full rate (with three register operands it is capped at 75 %).
//                        PCM arrives by cp.async.bulk (one 128-byte row per utterance and 32-sample chunk, four chunks in
//                        flight, mbarrier completion), is converted to fp64 once per CTA into a double-buffered ring and read by
//                        all eight warps: no filter warp ever waits on global memory.  Output: raw window energy sums to the
//                        energy planes (L2), then one release-increment of the unit's group counter.
//   warps 8-15  ENCODER + RESERVOIR + READOUT, two units of 128 threads.  A unit takes the next utterance whose group is complete
//                        (acquire-poll of the group counter), runs the speculative encoder epilogue with the derived error bound
//                        (gammatone_core.cuh), hands the spikes over as bits in shared memory, simulates the reservoir and
//                        writes the feature row (reservoir_core.cuh).  Utterances the bound cannot settle go to a work list
//                        that the host follows up with the exact kernel.
//
// Filter warps never wait for the other role, so every resident CTA makes progress; units are dealt statically (blockIdx +
// k * gridDim), which keeps the channel index provably uniform for the compiler.  setmaxnreg moves registers from the filter
// warps (few live values) to the reservoir warps.  A launch has one CTA per SM; two launches (the two lanes of a ctx) share an
// SM, so the fill and drain phases of one overlap the steady state of the other.
#include <stdlib.h>

#include <memory>
#include <new>

#include "gammatone_core.cuh"

namespace {

constexpr int kFW = 8;                    // filter warps per CTA = channels per unit
constexpr int kEU = 2;                    // encoder / reservoir units per CTA
constexpr int kEThreads = 128;            // threads per unit = channels (one thread per channel), 8 neurons per thread
constexpr int kThreads = kFW * 32 + kEU * kEThreads;
constexpr int kUPGShift = 4;              // units per 32-utterance group = 128 channels / 8 = 16 (shifts keep the unit -> channel
constexpr int kUPG = 1 << kUPGShift;      // arithmetic on the uniform datapath; a division would move it to vector registers)
constexpr int kChunk = 32;                // samples per staged chunk
constexpr int kRawStages = 4;             // bulk copies in flight
constexpr int kRowPitch = kChunk * 8 + 16;  // bytes per utterance row of the fp64 ring (16-byte loads of 8 lanes hit 8 bank groups)
constexpr int kBarFilter = 3;             // named barriers: 1, 2 = the units; 3 = the filter warps
constexpr int kRegsFilter = 56, kRegsUnit = 72;   // per-thread registers after setmaxnreg (256 x 56 + 256 x 72 = 512 x 64)

struct PipeArgs {
    GtArgs gt;                // front end + reservoir (gt.res); gt.pcm / gt.pcm16 = the batch
    double *energy;           // [B][ncols][C] raw window energy sums
    float *xmax;              // [B] max |sample|
    int *done;                // [groups] units completed per 32-utterance group (zero at launch)
    int *unit_next;           // encoder units' utterance counter (zero at launch)
    int *err_flag;            // set if a unit gave up waiting (should never happen)
    int n_units;              // filter units in this launch (kUPG per 32-utterance group)
    int n_chunks, n_groups8;  // chunks per unit, last 8-sample group index
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
// global -> shared bulk copy (the TMA engine's linear form), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
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
    unsigned char *ring;              // [2][32][kRowPitch] fp64 samples, row = utterance of the group
    unsigned char *raw;               // [kRawStages][32][row_bytes] PCM as it sits in memory
    unsigned long long *full;         // [kRawStages] mbarriers: chunk landed
};

// Issue the bulk copies of chunk c of unit u into raw stage `stage`.  Called by warp 0 of the filter group with all 32
// lanes: lane r copies utterance row r of the unit's group.
template <bool I16>
__device__ __forceinline__ void issue_chunk(const PipeArgs &a, const FilterSmem &sm, int u, int c, int stage, int lane)
{
    const int g = u >> kUPGShift;
    const int row_bytes = kChunk * (I16 ? 2 : 4);
    int n = a.gt.L - c * kChunk;                       // samples of this chunk that exist (> 0 for every chunk that is needed)
    n = n > kChunk ? kChunk : n;
    const unsigned bytes = (unsigned)n * (I16 ? 2u : 4u);
    if (lane == 0) mbar_expect_tx(sm.full + stage, bytes * 32u);
    __syncwarp();
    int utt = g * 32 + lane;
    utt = utt < a.gt.B ? utt : a.gt.B - 1;              // rows past the batch re-read the last utterance (results discarded)
    const unsigned char *src = I16 ? reinterpret_cast<const unsigned char *>(a.gt.pcm16 + (size_t)utt * a.gt.L + (size_t)c * kChunk)
                                   : reinterpret_cast<const unsigned char *>(a.gt.pcm + (size_t)utt * a.gt.L + (size_t)c * kChunk);
    bulk_g2s(sm.raw + ((size_t)stage * 32 + lane) * row_bytes, src, bytes, sm.full + stage);
}

// ---- FILTER role, warp W of the filter group: channel = 8 * (unit's channel block) + W for the unit's 32 utterances.
// Everything that selects the channel (u, cb, ch) is built from blockIdx, gridDim and loop counters with adds, shifts and
// masks only, so that it stays on the uniform datapath and the six coefficients reach the DFMAs as uniform registers.
template <bool I16>
__device__ __forceinline__ void filter_role(const PipeArgs &a, const FilterSmem &sm, const int W)
{
    const int lane = threadIdx.x & 31;
    const int ftid = threadIdx.x;                       // 0..255 within the filter group
    const int hop = a.gt.hop, ncols = a.gt.ncols, C = a.gt.C;
    const int r_old = a.gt.nwin - 2 * hop;
    const int row_bytes = kChunk * (I16 ? 2 : 4);
    // conversion slice of this thread: 4 consecutive samples of row crow
    const int crow = ftid >> 3, cseg = ftid & 7;

    // producer cursor (warp 0): the next chunk to issue, kRawStages chunks ahead of the consumers, across unit boundaries
    int pu = blockIdx.x, pc = 0;
    unsigned pseq = 0;
#define LSM_PIPE_ISSUE()                                                                \
    if (W == 0 && pu < a.n_units) {                                                     \
        issue_chunk<I16>(a, sm, pu, pc, (int)(pseq & (kRawStages - 1)), lane);          \
        ++pseq;                                                                         \
        if (++pc == a.n_chunks) { pc = 0; pu += gridDim.x; }                            \
    }
#pragma unroll 1
    for (int s = 0; s < kRawStages; ++s) { LSM_PIPE_ISSUE() }

    unsigned seq = 0;
#pragma unroll 1
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        const int g = u >> kUPGShift, cb = u & (kUPG - 1);
        const int ch = cb * kFW + W;
        const double *cf = a.coef[ch];                  // read where they are used: ptxas keeps them as uniform-register operands
        const int utt = g * 32 + lane;
        const bool valid = utt < a.gt.B;
        double *dst = a.energy + (size_t)(valid ? utt : 0) * ncols * C + ch;

        double xp = 0.0, p1 = 0.0, q1 = 0.0, p2 = 0.0, q2 = 0.0, p3 = 0.0, q3 = 0.0, p4 = 0.0, q4 = 0.0;
        double acc = 0.0, full1 = 0.0, full2 = 0.0;
        float xm = 0.0f;
        int next_b = r_old, m = 0, gi = 0;
        bool head = true;

#define LSM_PIPE_SAMPLE(xv)                                                                   \
        {                                                                                     \
            const double x_ = (xv);                                                           \
            const double y1 = fma(cf[4], p1, fma(cf[5], q1, fma(cf[0], xp, x_)));             \
            const double y2 = fma(cf[4], p2, fma(cf[5], q2, fma(cf[1], q1, p1)));             \
            const double y3 = fma(cf[4], p3, fma(cf[5], q3, fma(cf[2], q2, p2)));             \
            const double y4 = fma(cf[4], p4, fma(cf[5], q4, fma(cf[3], q3, p3)));             \
            xp = x_;                                                                          \
            q1 = p1; p1 = y1; q2 = p2; p2 = y2; q3 = p3; p3 = y3; q4 = p4; p4 = y4;           \
            acc = fma(y4, y4, acc);                                                           \
        }

#pragma unroll 1
        for (int c = 0; c < a.n_chunks; ++c, ++seq) {
            const int stage = (int)(seq & (kRawStages - 1));
            mbar_wait(sm.full + stage, (seq >> 2) & 1u);
            // PCM -> fp64, once per CTA: this thread's 4 samples of row crow into ring buffer c & 1
            {
                const unsigned char *rp = sm.raw + ((size_t)stage * 32 + crow) * row_bytes;
                double v0, v1, v2, v3;
                if (I16) {
                    const short4 h = *reinterpret_cast<const short4 *>(rp + cseg * 8);
                    v0 = __dmul_rn((double)h.x, 0x1p-15); v1 = __dmul_rn((double)h.y, 0x1p-15);
                    v2 = __dmul_rn((double)h.z, 0x1p-15); v3 = __dmul_rn((double)h.w, 0x1p-15);
                    xm = fmaxf(fmaxf(xm, fabsf((float)v0)), fmaxf(fmaxf(fabsf((float)v1), fabsf((float)v2)), fabsf((float)v3)));
                } else {
                    const float4 f = *reinterpret_cast<const float4 *>(rp + cseg * 16);
                    xm = fmaxf(fmaxf(xm, fabsf(f.x)), fmaxf(fmaxf(fabsf(f.y), fabsf(f.z)), fabsf(f.w)));
                    v0 = (double)f.x; v1 = (double)f.y; v2 = (double)f.z; v3 = (double)f.w;
                }
                const int s0 = c * kChunk + cseg * 4;             // samples past the end of the utterance are zeros
                if (s0 + 3 >= a.gt.L) {
                    if (s0 + 0 >= a.gt.L) v0 = 0.0;
                    if (s0 + 1 >= a.gt.L) v1 = 0.0;
                    if (s0 + 2 >= a.gt.L) v2 = 0.0;
                    if (s0 + 3 >= a.gt.L) v3 = 0.0;
                }
                double2 *wp = reinterpret_cast<double2 *>(sm.ring + (size_t)(c & 1) * 32 * kRowPitch + (size_t)crow * kRowPitch + cseg * 32);
                wp[0] = make_double2(v0, v1);
                wp[1] = make_double2(v2, v3);
            }
            res_sync<kBarFilter>(kFW * 32);                       // ring buffer c & 1 complete; raw stage free
            LSM_PIPE_ISSUE()

            const double2 *row = reinterpret_cast<const double2 *>(sm.ring + (size_t)(c & 1) * 32 * kRowPitch + (size_t)lane * kRowPitch);
#pragma unroll
            for (int s = 0; s < kChunk / 8; ++s) {
                if (gi > a.n_groups8) break;
                // group gi completes outputs 8gi-3 .. 8gi+4 (software skew: stage k works one iteration behind stage k-1), so a
                // window-phase boundary at sample 8gi falls after the group's third iteration
                const double2 d0 = row[4 * s + 0], d1 = row[4 * s + 1];
                LSM_PIPE_SAMPLE(d0.x) LSM_PIPE_SAMPLE(d0.y) LSM_PIPE_SAMPLE(d1.x)
                if (8 * gi == next_b) {
                    if (head) {
                        if (m >= 2 && valid) __stcg(dst + (size_t)(m - 2) * C, (full2 + full1) + acc);   // window m-2 complete
                        next_b += hop - r_old;
                        head = false;
                    } else {
                        full2 = full1; full1 = acc; acc = 0.0;
                        next_b += r_old;
                        head = true;
                        ++m;
                    }
                }
                const double2 d2 = row[4 * s + 2], d3 = row[4 * s + 3];
                LSM_PIPE_SAMPLE(d1.y) LSM_PIPE_SAMPLE(d2.x) LSM_PIPE_SAMPLE(d2.y) LSM_PIPE_SAMPLE(d3.x) LSM_PIPE_SAMPLE(d3.y)
                ++gi;
            }
        }
#undef LSM_PIPE_SAMPLE
        // peak level of each utterance (the bound scales with it): the eight threads that converted a row share it
        if (cb == 0) {
            xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, 1));
            xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, 2));
            xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, 4));
            const int xu = g * 32 + crow;
            if (cseg == 0 && xu < a.gt.B) __stcg(a.xmax + xu, xm);
        }
        // publish the unit: every writer fences, the group meets, one thread counts the unit as done (release)
        __threadfence();
        res_sync<kBarFilter>(kFW * 32);
        if (ftid == 0) {
            __threadfence();
            atomicAdd(a.done + g, 1);
        }
    }
#undef LSM_PIPE_ISSUE
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
                while (ld_acquire(flag) < kUPG) {
                    __nanosleep(256);
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
__global__ void __launch_bounds__(kThreads, 2) pipeline_kernel(const __grid_constant__ PipeArgs a, const int unit_smem_bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ double s_red[kEU][6 * 8];
    __shared__ double s_out[kEU][6];
    __shared__ int s_cnt[kEU][5];
    __shared__ int s_utt[kEU];
    __shared__ __align__(8) unsigned long long s_full[kRawStages];

    // the warp index as a value the compiler knows to be warp-uniform (REDUX writes a uniform register): the filter warps'
    // channel index, and with it their coefficients, stay on the uniform datapath
    const int warp = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5));
    FilterSmem fs;
    fs.ring = smem + (size_t)kEU * unit_smem_bytes;
    fs.raw = fs.ring + 2 * 32 * kRowPitch;
    fs.full = s_full;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kRawStages; ++s) mbar_init(s_full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < kFW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsFilter));
        filter_role<I16>(a, fs, warp);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsUnit));
        const int e = (warp - kFW) >> 2;
        const int tid = threadIdx.x - kFW * 32 - e * kEThreads;
        unsigned char *usmem = smem + (size_t)e * unit_smem_bytes;
        if (e == 0) unit_role<1>(a, usmem, s_red[0], s_out[0], s_cnt[0], &s_utt[0], tid);
        else unit_role<2>(a, usmem, s_red[1], s_out[1], s_cnt[1], &s_utt[1], tid);
    }
}

}  // namespace

// Can this pair run as the warp-specialised kernel?  The reference's default shape: 128 gammatone channels, no redundancy, a lean
// reservoir of <= 1024 neurons (8 per thread), window phases in whole 8-sample groups.
bool lsm_pipeline_lanes_eligible(const lsm_frontend *fe, const lsm_reservoir *res, const void *d_pcm, bool i16)
{
    const lsm_frontend_params &p = fe->p;
    if (getenv("LSM_NO_PIPELINE") || !d_pcm) return false;
    if (lsm_fused_npt(fe, res) != 8 || p.channels != 128 || !res->lean || res->n_pad != 1024) return false;
    if (fe->mode != LSM_FILTER_SPECULATIVE) return false;
    const int r_old = p.nwin - 2 * p.hop;
    if (p.hop % 8 || r_old % 8 || r_old <= 0) return false;
    if (p.n_samples % (i16 ? 8 : 4) || (((uintptr_t)d_pcm) & 15)) return false;
    if (p.n_samples < kChunk) return false;
    return true;
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
    const int n_used = (fe->ncols - 1) * p.hop + p.nwin;
    a.n_groups8 = n_used / 8;
    a.n_chunks = (8 * (a.n_groups8 + 1) + kChunk - 1) / kChunk;
    memcpy(a.coef, fe->h_lane_coef, sizeof(double) * 6 * 128);
    const int unit_smem = (int)((lsm_res_smem_bytes(a.gt.res.T, a.gt.res.CW, kEThreads * 8, a.gt.res.N) + 127) & ~(size_t)127);
    const size_t smem = (size_t)kEU * unit_smem + 2 * 32 * kRowPitch + (size_t)kRawStages * 32 * kChunk * 4;
    if (smem > 113 * 1024) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "pipeline kernel: %zu bytes of shared memory per CTA (two per SM must fit)", smem);
    LSM_CUDA(ctx, cudaMemsetAsync(sync, 0, sizeof(int) * (4 + (size_t)groups), st));
    LSM_CUDA(ctx, cudaMemsetAsync(a.gt.rerun_list, 0, sizeof(int), st));
    int grid = ctx->sm_count;
    if (grid > a.n_units) grid = a.n_units;
    if (i16) {
        LSM_CUDA(ctx, cudaFuncSetAttribute(pipeline_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pipeline_kernel<true><<<grid, kThreads, smem, st>>>(a, unit_smem);
    } else {
        LSM_CUDA(ctx, cudaFuncSetAttribute(pipeline_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pipeline_kernel<false><<<grid, kThreads, smem, st>>>(a, unit_smem);
    }
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    // exact pass: the lane = channel kernel in exact mode over the device work list (typically empty; a few CTAs suffice)
    GtArgs x = a.gt;
    x.mode = 0; x.rerun_list = nullptr;
    x.utt_list = a.gt.rerun_list + 1; x.utt_count = a.gt.rerun_list;
    return lsm_launch_fused_args(ctx, fe, res, x, st, true, nullptr, 8, lane);
}
