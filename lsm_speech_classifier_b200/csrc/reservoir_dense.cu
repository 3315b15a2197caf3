// K2 dense arm — the recurrent current of a whole batch as a tensor-core contraction (tcgen05.mma kind::i8, accumulators in TMEM).
//
// Replaces the same reference work as reservoir.cu (/root/reference/extract_lsm_features.py:79-87: reset / set_input_spike_times /
// simulate / extract_features_from_spikes for every utterance) and produces the same bits: BASELINE.json's north star names two
// ways to form  I_rec[b, i] = sum_j W[i, j] * s[b, j](t-1)  - the event-driven gather over the spiking set (reservoir_core.cuh)
// or, "when batching makes it a genuine dense contraction (spikes[B,N] . W[N,N])", a tensor-core tile - "choosing whichever ncu
// shows wins".  This file is the second arm, so that the choice is measured (profiles/r2_config4.md), not argued.
//
// Exactness.  Weights are int32 multiples of 2^-w_shift.  They are cut into 8-bit digit planes
//     w = d0 + d1 * 2^8 + d2 * 2^16 (+ d3 * 2^24, signed, only when some weight is negative or >= 2^24)
// and each plane is contracted with the 0/1 spike matrix by the integer tensor cores (u8 x u8 / u8 x s8 -> s32): every partial
// sum is an exact integer (<= 255 * K < 2^31), the planes are recombined with shifts in the epilogue, and the membrane update
// (spec R6: three fp64 roundings) runs on the exact integer sum - the same value the event-driven kernel and the CPU oracle form.
//
// Shape of the computation.  Time steps are sequential (the spikes of step t are the A operand of step t+1), so one launch
// advances the whole batch by one step:
//     grid = (Bp / 128 utterance tiles) x (NP / 64 neuron tiles), 320 threads per CTA, two CTAs per SM (the TMA / MMA phase of
//     one overlaps the epilogue of the other: 256 of the 512 TMEM columns and ~80 KB of shared memory each)
//     warp 8   producer: TMA 2-D tiles (rows x 128 bytes of K, SWIZZLE_128B) of the spike matrix S[Bp][NK] (A, K-major, 128
//              rows) and of each digit plane Wd[p][NP][NK] (B, K-major, 64 rows) into a two-slot shared-memory ring,
//              full / empty mbarriers
//     warp 9   one thread issues tcgen05.mma (M = 128, N = 64, K = 32 per instruction), one accumulator block of 64 TMEM
//              columns per digit plane; tcgen05.commit releases ring slots and finally signals the epilogue
//     warps 0-7 epilogue: thread = utterance (TMEM lane; lane quarter = warp & 3) x 32 of the tile's neurons (warp >> 2),
//              tcgen05.ld 8 columns x planes at a time, recombine, then the LIF update on V[n][b] / refractory[n][b]
//              (neuron-major planes: coalesced across the warp; the next block's state is requested while the current one is
//              computed, the first block's before the accumulators are complete), the spike statistics of the readout
//              (two 16-byte records per neuron-utterance, touched only on a spike) and the spike bytes of S_next[b][n]
// The readout (K3, spec R9) is a small kernel over the statistics after the last step.
#include <stdlib.h>

#include <new>
#include <vector>

#include <cuda.h>

#include "reservoir_core.cuh"

namespace {

constexpr int kTile = 128;                  // utterances per CTA (MMA M) = bytes of K per ring slot
constexpr int kNT = 64;                     // neurons per CTA (MMA N): 64 columns x <= 4 digit planes = 256 TMEM columns, two CTAs per SM
constexpr int kEpiWarps = 8;                // epilogue warps: TMEM lane quarter = warp & 3, column half = warp >> 2
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kATileBytes = kTile * kTile;  // 128 utterances x 128 bytes of K
constexpr int kBTileBytes = kNT * kTile;    // 64 neurons x 128 bytes of K
constexpr int kStages = 2;

struct DenseArgs {
    double *V;               // [NP][Bp]
    uint8_t *ref;            // [NP][Bp]
    uint8_t *s_next;         // [Bp][NK] spikes of this step (next step's A operand)
    const uint8_t *in_t;     // [T][C][Bp] input level bytes, time-major
    int4 *stat;              // [NP][Bp][2]: {count, sum t, first, last}, {sum isi^2, bursts, -, -}
    const int32_t *in_row;   // [NP] internal neuron -> its input row, -1 none
    const double *gain;      // [NP]
    const double *leak;      // [NP]
    const int32_t *ext_id;   // [NP] internal -> external neuron index (>= N: padding)
    uint8_t *raster;         // optional [B][T][N]
    int32_t *probe;          // optional [Bp][NP]: write the recombined integer sums and do nothing else (diagnostic)
    int B, Bp, N, NP, NK, C, T, t, refractory, planes, top_signed;
    double theta, scale;
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
        if (!ok && ++tries > (1u << 26)) __trap();       // never hang the GPU: abort the launch
    }
}
__device__ __forceinline__ void tma_tile_g2s(void *dst, const CUtensorMap *map, int x, int y, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// Shared-memory matrix descriptor, K-major operand tile of 128-byte rows laid out by TMA with SWIZZLE_128B: rows 128 bytes apart,
// groups of 8 rows 1024 bytes apart (stride byte offset), descriptor version 1 (sm_100), layout type 2 (128-byte swizzle).
// Advancing along K inside the swizzle atom is an advance of the start address (32 bytes per MMA of K = 32).
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr)
{
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::i8: D = s32, A = u8, B = u8 or s8, both K-major, N = kNT, M = 128.
__device__ __forceinline__ unsigned umma_idesc(bool b_signed)
{
    return (2u << 4) | (0u << 7) | ((b_signed ? 1u : 0u) << 10) | ((unsigned)(kNT >> 3) << 17) | ((unsigned)(kTile >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate)
{
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p; }"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, unsigned (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2) dense_step_kernel(const DenseArgs a, const __grid_constant__ CUtensorMap map_s,
                                                                 const __grid_constant__ CUtensorMap map_w)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages], s_acc;
    __shared__ unsigned s_tmem;

    // the swizzle pattern is a function of the shared-memory address: tiles start on 1024-byte boundaries
    unsigned char *const ring = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kTile, n0 = blockIdx.y * kNT;
    const int P = a.planes;
    const unsigned stage_bytes = (unsigned)kATileBytes + (unsigned)P * kBTileBytes;
    const int n_chunks = a.NK / kTile;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(s_full + s, 1); mbar_init(s_empty + s, 1); }
        mbar_init(&s_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kEpiWarps + 1) {
        // 256 columns = up to four accumulator blocks of 64; two CTAs per SM share the 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;

    if (warp == kEpiWarps) {
        // ---- producer
        if (lane == 0) {
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % kStages, round = c / kStages;
                if (round > 0) mbar_wait(s_empty + s, (unsigned)(round - 1) & 1u);
                unsigned char *slot = ring + (size_t)s * stage_bytes;
                mbar_expect_tx(s_full + s, stage_bytes);
                tma_tile_g2s(slot, &map_s, c * kTile, m0, s_full + s);
                for (int p = 0; p < P; ++p) tma_tile_g2s(slot + kATileBytes + (size_t)p * kBTileBytes, &map_w, c * kTile, p * a.NP + n0, s_full + s);
            }
        }
    } else if (warp == kEpiWarps + 1) {
        // ---- MMA issuer
        if (lane == 0) {
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % kStages, round = c / kStages;
                mbar_wait(s_full + s, (unsigned)round & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned slot = smem_u32(ring + (size_t)s * stage_bytes);
                for (int p = 0; p < P; ++p) {
                    const unsigned idesc = umma_idesc(a.top_signed && p == P - 1);
#pragma unroll
                    for (int kk = 0; kk < kTile / 32; ++kk)
                        umma_i8(tmem + (unsigned)p * kNT, umma_desc(slot + kk * 32), umma_desc(slot + kATileBytes + p * kBTileBytes + kk * 32),
                                idesc, (c > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(s_empty + s);                  // the slot may be refilled once these MMAs have read it
            }
            umma_commit(&s_acc);                           // all accumulators complete
        }
    } else {
        // ---- epilogue: thread = utterance row of the tile = TMEM lane (quarter = warp & 3); warps 0-3 take the first 32
        //      neuron columns of the tile, warps 4-7 the second
        const int q = warp & 3, half = warp >> 2;
        const int u = m0 + q * 32 + lane;
        const bool valid = u < a.B;
        const unsigned lane_base = tmem + ((unsigned)(q * 32) << 16);
        const size_t Bp = (size_t)a.Bp;
        const int nb = n0 + half * 32;                      // first neuron of this thread's 32
        // state of the first block, requested before the accumulators are complete (it does not depend on them)
        double v0[8];
        unsigned rf[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const size_t idx = (size_t)(nb + j) * Bp + u;
            v0[j] = (valid && !a.probe) ? __ldcg(a.V + idx) : 0.0;
            rf[j] = (valid && !a.probe) ? (unsigned)__ldcg(a.ref + idx) : 0u;
        }
        mbar_wait(&s_acc, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
            const int nc = nb + cb * 8;
            unsigned d[4][8];
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (p < P) tmem_ld8(lane_base + (unsigned)(p * kNT + half * 32 + cb * 8), d[p]);
            // next block's state while the tensor-memory loads are in flight
            double v1[8];
            unsigned rf1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool more = valid && !a.probe && cb < 3;
                const size_t idx = (size_t)(nc + 8 + j) * Bp + u;
                v1[j] = more ? __ldcg(a.V + idx) : 0.0;
                rf1[j] = more ? (unsigned)__ldcg(a.ref + idx) : 0u;
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            unsigned out_lo = 0u, out_hi = 0u, fired = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = nc + j;
                unsigned accu = d[0][j];
                if (P > 1) accu += d[1][j] << 8;
                if (P > 2) accu += d[2][j] << 16;
                if (P > 3) accu += d[3][j] << 24;
                const int acc = (int)accu;              // exact: the true sum fits in int32 (checked at create), the rest is mod 2^32
                if (a.probe) { if (valid) a.probe[(size_t)u * a.NP + n] = acc; continue; }
                if (!valid) continue;
                const int ext = __ldg(a.ext_id + n);
                const int r = __ldg(a.in_row + n);
                double i_in = 0.0;
                if (r >= 0 && __ldg(a.in_t + ((size_t)a.t * a.C + r) * Bp + u)) i_in = add64(0.0, __ldg(a.gain + n));
                const double cur = add64(i_in, mul64((double)acc, a.scale));
                const double v = add64(sub64(v0[j], mul64(__ldg(a.leak + n), v0[j])), cur);
                const bool active = rf[j] == 0u;
                const bool fire = active && (v >= a.theta) && (ext < a.N);
                const size_t idx = (size_t)n * Bp + u;
                __stcg(a.V + idx, (active && !fire) ? v : 0.0);
                a.ref[idx] = (uint8_t)(fire ? (unsigned)a.refractory : (active ? 0u : rf[j] - 1u));
                if (fire) { fired |= 1u << j; if (j < 4) out_lo |= 1u << (8 * j); else out_hi |= 1u << (8 * (j - 4)); }
                if (a.raster && ext < a.N) a.raster[((size_t)u * a.T + a.t) * a.N + ext] = fire ? 1 : 0;
            }
            // spikes are rare: statistics of the readout (spec R9) only for the neurons that fired
            while (fired) {
                const int j = __ffs(fired) - 1;
                fired &= fired - 1u;
                int4 *st = a.stat + ((size_t)(nc + j) * Bp + u) * 2;
                int4 s0 = st[0];                        // count, sum t, first, last
                if (s0.x > 0) {
                    int4 s1 = st[1];                    // sum isi^2, bursts
                    const int isi = a.t - s0.w;
                    s1.x += isi * isi;
                    if (isi <= a.refractory + 1) s1.y += 1;
                    st[1] = s1;
                } else s0.z = a.t;
                s0.x += 1; s0.y += a.t; s0.w = a.t;
                st[0] = s0;
            }
            if (!a.probe && valid) *reinterpret_cast<uint2 *>(a.s_next + (size_t)u * a.NK + nc) = make_uint2(out_lo, out_hi);
#pragma unroll
            for (int j = 0; j < 8; ++j) { v0[j] = v1[j]; rf[j] = rf1[j]; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kEpiWarps + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// Digit planes of the weights, K-major: Wd[p][n][k] = digit p of wt[k][n] (wt = presynaptic-major int32 plane of the reservoir)
__global__ void dense_planes_kernel(const int32_t *wt, int n_pad, int n_rows, uint8_t *wd, int NP, int NK, int planes)
{
    __shared__ int32_t tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, n = n0 + threadIdx.x;
        tile[r][threadIdx.x] = (k < n_rows && n < n_pad) ? wt[(size_t)k * n_pad + n] : 0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int n = n0 + r, k = k0 + threadIdx.x;
        if (n < NP && k < NK) {
            const int32_t w = tile[threadIdx.x][r];
            for (int p = 0; p < planes; ++p)
                wd[((size_t)p * NP + n) * NK + k] = (uint8_t)((p == 3 ? (w >> 24) : (w >> (8 * p))) & 0xff);
        }
    }
}

// Input spike trains [B][C][T] -> time-major level bytes [T][C][Bp]
__global__ void dense_input_kernel(const uint8_t *spikes, int B, int C, int T, int Bp, uint8_t *in_t)
{
    __shared__ uint8_t tile[32][33];
    const int c = blockIdx.z;
    const int t0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int b = b0 + r, t = t0 + threadIdx.x;
        tile[r][threadIdx.x] = (b < B && t < T) ? spikes[((size_t)b * C + c) * T + t] : 0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int t = t0 + r, b = b0 + threadIdx.x;
        if (t < T && b < Bp) in_t[((size_t)t * C + c) * Bp + b] = tile[threadIdx.x][r];
    }
}

__global__ void dense_init_stat_kernel(int4 *stat, size_t plane)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane) return;
    stat[2 * i] = make_int4(0, 0, -1, -1);
    stat[2 * i + 1] = make_int4(0, 0, 0, 0);
}

// K3 (spec R9; the same formulas as the event-driven kernel's epilogue in reservoir_core.cuh): one thread per (utterance, output neuron)
__global__ void dense_readout_kernel(const int4 *stat, size_t plane, int Bp, int B, const int32_t *out_int, int n_out, int T,
                                     unsigned feature_mask, int nkeys, int nan_to_num, double *features, int *diag, int N,
                                     const int32_t *ext_id, int NP)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x, u = blockIdx.y;
    if (o >= n_out || u >= B) return;
    const size_t idx = (size_t)__ldg(out_int + o) * Bp + u;
    const int4 s0 = stat[2 * idx], s1 = stat[2 * idx + 1];
    const int cnt = s0.x, sumt = s0.y, first = s0.z, last = s0.w, s2 = s1.x, burst = s1.y;
    const double c = (double)cnt, nan = __longlong_as_double(0x7ff8000000000000LL);
    double *f = features + (size_t)u * nkeys * n_out;
    int slot = 0;
    for (int key = 0; key < 8; ++key) {
        if (!(feature_mask & (1u << key))) continue;
        double v = nan;
        switch (key) {
        case 0: v = c; break;
        case 1: { const double p = __ddiv_rn(c, (double)T); v = mul64(p, sub64(1.0, p)); } break;
        case 2: if (cnt >= 1) v = __ddiv_rn((double)sumt, c); break;
        case 3: if (cnt >= 1) v = (double)first; break;
        case 4: if (cnt >= 1) v = (double)last; break;
        case 5: if (cnt >= 2) v = __ddiv_rn((double)(last - first), (double)(cnt - 1)); break;
        case 6: if (cnt >= 2) {
                    const long long n = cnt - 1, s1 = last - first;
                    v = __ddiv_rn((double)(n * (long long)s2 - s1 * s1), (double)(n * n));
                } break;
        default: v = (double)burst; break;
        }
        if (nan_to_num && v != v) v = 0.0;
        f[(size_t)slot * n_out + o] = v;
        ++slot;
    }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn encode_tiled()
{
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// u8 matrix [rows][cols] (cols contiguous) as a 2-D tensor with box_rows x 128-byte boxes, 128-byte swizzle
int make_map(lsm_ctx *ctx, CUtensorMap *map, const void *base, size_t rows, size_t cols, int box_rows)
{
    encode_tiled_fn fn = encode_tiled();
    if (!fn) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols};
    const cuuint32_t box[2] = {kTile, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LSM_FAIL(ctx, LSM_ERR_CUDA, "cuTensorMapEncodeTiled -> %d", (int)r);
    return LSM_OK;
}

}  // namespace

// Workspace of the dense arm, owned by the reservoir (created on first use, grown on demand)
struct lsm_dense_ws {
    int NP = 0, NK = 0, planes = 0, top_signed = 0;
    uint8_t *d_wd = nullptr;        // [planes][NP][NK]
    int32_t *d_in_row = nullptr, *d_out_int = nullptr;
    double *d_gain = nullptr, *d_leak = nullptr;
    int cap = 0;                    // utterances the state buffers hold (multiple of 128)
    double *d_V = nullptr;
    uint8_t *d_ref = nullptr, *d_S = nullptr /* [2][cap][NK] */, *d_in_t = nullptr;
    int4 *d_stat = nullptr;
};

void lsm_dense_ws_free(lsm_dense_ws *w)
{
    if (!w) return;
    cudaFree(w->d_wd); cudaFree(w->d_in_row); cudaFree(w->d_out_int); cudaFree(w->d_gain); cudaFree(w->d_leak);
    cudaFree(w->d_V); cudaFree(w->d_ref); cudaFree(w->d_S); cudaFree(w->d_in_t); cudaFree(w->d_stat);
    delete w;
}

// Why the dense arm cannot serve this reservoir, or nullptr.
const char *lsm_dense_unsupported(const lsm_reservoir *res)
{
    if (res->w64) return "strict reservoirs (fp64 weights, ordered sums) have no integer digit planes";
    if (res->max_in_per_neuron > 1) return "a neuron is driven by several input rows";
    if (res->p.refractory > 255) return "refractory period > 255";
    if (res->n_gather > 0) return "the fused all-gather is an epilogue of the event-driven kernel";
    if (!res->h_dense_in_row) return "reservoir created without its host-side dense tables";
    return nullptr;
}

static int dense_prepare(lsm_ctx *ctx, lsm_reservoir *res, int B, cudaStream_t st)
{
    const lsm_reservoir_params &p = res->p;
    if (!res->dense) {
        lsm_dense_ws *w = new (std::nothrow) lsm_dense_ws();
        if (!w) LSM_FAIL(ctx, LSM_ERR_NOMEM, "out of host memory");
        res->dense = w;
        w->NP = (res->n_pad + kTile - 1) / kTile * kTile;
        w->NK = w->NP;
        w->planes = res->dense_planes;
        w->top_signed = res->dense_top_signed;
        const size_t np = (size_t)w->NP;
        LSM_CUDA(ctx, cudaMalloc(&w->d_wd, (size_t)w->planes * np * w->NK));
        LSM_CUDA(ctx, cudaMalloc(&w->d_in_row, np * sizeof(int32_t)));
        LSM_CUDA(ctx, cudaMalloc(&w->d_gain, np * sizeof(double)));
        LSM_CUDA(ctx, cudaMalloc(&w->d_leak, np * sizeof(double)));
        LSM_CUDA(ctx, cudaMalloc(&w->d_out_int, (size_t)(p.n_out > 0 ? p.n_out : 1) * sizeof(int32_t)));
        // host tables are sized n_pad; pad to NP with "no input, no leak"
        std::vector<int32_t> in_row(np, -1);
        std::vector<double> gain(np, 0.0), leak(np, 0.0);
        for (int i = 0; i < res->n_pad; ++i) { in_row[i] = res->h_dense_in_row[i]; gain[i] = res->h_dense_gain[i]; leak[i] = res->h_dense_leak[i]; }
        LSM_CUDA(ctx, cudaMemcpyAsync(w->d_in_row, in_row.data(), np * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        LSM_CUDA(ctx, cudaMemcpyAsync(w->d_gain, gain.data(), np * sizeof(double), cudaMemcpyHostToDevice, st));
        LSM_CUDA(ctx, cudaMemcpyAsync(w->d_leak, leak.data(), np * sizeof(double), cudaMemcpyHostToDevice, st));
        LSM_CUDA(ctx, cudaMemcpyAsync(w->d_out_int, res->h_dense_out_int, (size_t)p.n_out * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        LSM_CUDA(ctx, cudaStreamSynchronize(st));       // the vectors above go out of scope
        const dim3 grid((w->NK + 31) / 32, (w->NP + 31) / 32), block(32, 8);
        dense_planes_kernel<<<grid, block, 0, st>>>(res->d_wt, res->n_pad, res->zero_row, w->d_wd, w->NP, w->NK, w->planes);
        ctx->launches += 1;
        LSM_CUDA(ctx, cudaGetLastError());
    }
    lsm_dense_ws *w = res->dense;
    const int need = (B + kTile - 1) / kTile * kTile;
    if (w->cap < need) {
        LSM_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(w->d_V); cudaFree(w->d_ref); cudaFree(w->d_S); cudaFree(w->d_in_t); cudaFree(w->d_stat);
        w->d_V = nullptr; w->d_ref = nullptr; w->d_S = nullptr; w->d_in_t = nullptr; w->d_stat = nullptr; w->cap = 0;
        const size_t plane = (size_t)w->NP * need;
        if (cudaMalloc(&w->d_V, plane * sizeof(double)) != cudaSuccess || cudaMalloc(&w->d_ref, plane) != cudaSuccess ||
            cudaMalloc(&w->d_S, 2 * (size_t)need * w->NK) != cudaSuccess ||
            cudaMalloc(&w->d_in_t, (size_t)p.num_steps * p.num_inputs * need) != cudaSuccess ||
            cudaMalloc(&w->d_stat, 2 * plane * sizeof(int4)) != cudaSuccess) {
            cudaGetLastError();
            LSM_FAIL(ctx, LSM_ERR_NOMEM, "dense reservoir arm: out of device memory for %d utterances x %d neurons", need, w->NP);
        }
        w->cap = need;
    }
    return LSM_OK;
}

static void dense_fill(const lsm_reservoir *res, const lsm_dense_ws *w, int B, int Bp, DenseArgs *a)
{
    const lsm_reservoir_params &p = res->p;
    a->V = w->d_V; a->ref = w->d_ref; a->in_t = w->d_in_t; a->stat = w->d_stat;
    a->in_row = w->d_in_row; a->gain = w->d_gain; a->leak = w->d_leak; a->ext_id = res->d_ext_id;
    a->raster = nullptr; a->probe = nullptr;
    a->B = B; a->Bp = Bp; a->N = p.num_neurons; a->NP = w->NP; a->NK = w->NK; a->C = p.num_inputs; a->T = p.num_steps; a->t = 0;
    a->refractory = p.refractory; a->planes = w->planes; a->top_signed = w->top_signed;
    a->theta = p.theta; a->scale = ldexp(1.0, -p.w_shift);
}

// All T steps for utterances [0, B) (B <= the workspace's capacity is arranged here), then the readout.
int lsm_launch_reservoir_dense(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int B, uint32_t feature_mask,
                               int nan_to_num, double *d_features, uint8_t *d_raster, cudaStream_t st)
{
    if (B <= 0) return LSM_OK;
    if (const char *why = lsm_dense_unsupported(res)) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "dense reservoir arm: %s", why);
    const lsm_reservoir_params &p = res->p;
    const int kMaxChunk = 8192;
    const int nkeys = __builtin_popcount(feature_mask & 0xFFu);
    for (int off = 0; off < B; off += kMaxChunk) {
        const int n = B - off < kMaxChunk ? B - off : kMaxChunk;
        int rc = dense_prepare(ctx, res, n, st);
        if (rc != LSM_OK) return rc;
        lsm_dense_ws *w = res->dense;
        const int Bp = w->cap;
        const size_t plane = (size_t)w->NP * Bp;
        DenseArgs a;
        dense_fill(res, w, n, Bp, &a);
        if (d_raster) a.raster = d_raster + (size_t)off * p.num_steps * p.num_neurons;
        LSM_CUDA(ctx, cudaMemsetAsync(w->d_V, 0, plane * sizeof(double), st));
        LSM_CUDA(ctx, cudaMemsetAsync(w->d_ref, 0, plane, st));
        LSM_CUDA(ctx, cudaMemsetAsync(w->d_S, 0, 2 * (size_t)Bp * w->NK, st));
        dense_init_stat_kernel<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(w->d_stat, plane);
        {
            const dim3 grid((p.num_steps + 31) / 32, (Bp + 31) / 32, p.num_inputs), block(32, 8);
            dense_input_kernel<<<grid, block, 0, st>>>(d_spikes + (size_t)off * p.num_inputs * p.num_steps, n, p.num_inputs, p.num_steps, Bp, w->d_in_t);
        }
        alignas(64) CUtensorMap map_s[2], map_w;
        for (int k = 0; k < 2; ++k)
            if ((rc = make_map(ctx, &map_s[k], w->d_S + (size_t)k * Bp * w->NK, (size_t)Bp, (size_t)w->NK, kTile)) != LSM_OK) return rc;
        if ((rc = make_map(ctx, &map_w, w->d_wd, (size_t)w->planes * w->NP, (size_t)w->NK, kNT)) != LSM_OK) return rc;
        const size_t smem = (size_t)kStages * (kATileBytes + w->planes * kBTileBytes) + 1024;
        LSM_CUDA(ctx, cudaFuncSetAttribute(dense_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const dim3 grid((n + kTile - 1) / kTile, w->NP / kNT);
        for (int t = 0; t < p.num_steps; ++t) {
            a.t = t;
            a.s_next = w->d_S + (size_t)((t + 1) & 1) * Bp * w->NK;          // step t reads S[t & 1] (spikes of t-1), writes S[(t+1) & 1]
            dense_step_kernel<<<grid, kThreads, smem, st>>>(a, map_s[t & 1], map_w);
        }
        ctx->launches += 2 + p.num_steps;
        if (d_features) {
            const dim3 rg((p.n_out + 127) / 128, n);
            dense_readout_kernel<<<rg, 128, 0, st>>>(w->d_stat, plane, Bp, n, w->d_out_int, p.n_out, p.num_steps, feature_mask & 0xFFu, nkeys,
                                                     nan_to_num, d_features + (size_t)off * nkeys * p.n_out, nullptr, p.num_neurons,
                                                     res->d_ext_id, w->NP);
            ctx->launches += 1;
        }
        LSM_CUDA(ctx, cudaGetLastError());
    }
    return LSM_OK;
}

// Diagnostic: one contraction only.  d_s: uint8[B][num_neurons] spike bytes (0/1) in the caller's (external) neuron order;
// d_acc: int32[B][num_neurons], d_acc[b][i] = sum_j Wq[i][j] * s[b][j] as the tensor cores and the recombination form it.
__global__ void dense_probe_in_kernel(const uint8_t *s, int B, int N, const int32_t *ext_id, int NK, uint8_t *S)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (k >= NK || b >= B) return;
    const int e = ext_id[k];
    S[(size_t)b * NK + k] = e < N ? (s[(size_t)b * N + e] ? 1 : 0) : 0;
}
__global__ void dense_probe_out_kernel(const int32_t *acc_int, int B, int N, const int32_t *ext_id, int NP, int32_t *acc)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (n >= NP || b >= B) return;
    const int e = ext_id[n];
    if (e < N) acc[(size_t)b * N + e] = acc_int[(size_t)b * NP + n];
}

int lsm_reservoir_dense_probe_launch(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_s, int B, int32_t *d_acc, cudaStream_t st)
{
    if (B <= 0) return LSM_OK;
    if (const char *why = lsm_dense_unsupported(res)) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "dense reservoir arm: %s", why);
    if (B > 8192) LSM_FAIL(ctx, LSM_ERR_INVALID, "lsm_reservoir_dense_probe: at most 8192 rows");
    int rc = dense_prepare(ctx, res, B, st);
    if (rc != LSM_OK) return rc;
    lsm_dense_ws *w = res->dense;
    const int Bp = w->cap;
    const lsm_reservoir_params &p = res->p;
    DenseArgs a;
    dense_fill(res, w, B, Bp, &a);
    void *d_tmp;
    if ((rc = lsm_stage_device(ctx, 7, (size_t)Bp * w->NP * sizeof(int32_t), &d_tmp)) != LSM_OK) return rc;
    a.probe = (int32_t *)d_tmp;
    a.s_next = w->d_S + (size_t)Bp * w->NK;
    LSM_CUDA(ctx, cudaMemsetAsync(w->d_S, 0, (size_t)Bp * w->NK, st));
    dense_probe_in_kernel<<<dim3((w->NK + 255) / 256, B), 256, 0, st>>>(d_s, B, p.num_neurons, res->d_ext_id, w->NK, w->d_S);
    alignas(64) CUtensorMap map_s, map_w;
    if ((rc = make_map(ctx, &map_s, w->d_S, (size_t)Bp, (size_t)w->NK, kTile)) != LSM_OK) return rc;
    if ((rc = make_map(ctx, &map_w, w->d_wd, (size_t)w->planes * w->NP, (size_t)w->NK, kNT)) != LSM_OK) return rc;
    const size_t smem = (size_t)kStages * (kATileBytes + w->planes * kBTileBytes) + 1024;
    LSM_CUDA(ctx, cudaFuncSetAttribute(dense_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_step_kernel<<<dim3((B + kTile - 1) / kTile, w->NP / kNT), kThreads, smem, st>>>(a, map_s, map_w);
    dense_probe_out_kernel<<<dim3((w->NP + 255) / 256, B), 256, 0, st>>>(a.probe, B, p.num_neurons, res->d_ext_id, w->NP, d_acc);
    ctx->launches += 3;
    LSM_CUDA(ctx, cudaGetLastError());
    return LSM_OK;
}
