// K2 + K3 — liquid-state-machine reservoir simulation with the feature readout fused in.
//
// Replaces, per utterance, /root/reference/extract_lsm_features.py:79-87:
//     lsm.reset(); lsm.set_input_spike_times(sample); lsm.simulate()      (snnpy, un-vendored)
//     lsm.extract_features_from_spikes(); key selection; np.nan_to_num
// Semantics = the frozen reservoir spec (oracle/lsm_oracle.c simulate_one).
//
// Mapping: one CTA per utterance, all T steps inside the kernel; membrane potential, refractory
// counter and the streaming feature statistics of every neuron live in registers for the whole
// simulation (thread t owns neurons t, t+blockDim, ...).  The recurrent current is event driven:
// the neurons that fired in step t-1 are compacted (warp ballot + popc) into a shared-memory list
// and every thread adds the matching rows of the dense, presynaptic-major int32 weight plane
// (coalesced 4-byte loads, L2 resident: 4 MB at N = 1000).  Weights are integers (multiples of
// 2^-24), so the sum is exact in any order and bit-identical to the oracle's.
#include "lsm_common.cuh"

namespace {

struct ResArgs {
    const uint8_t *spikes;    // [B][C][T]
    const int32_t *wt;        // [N][n_pad]  row = presynaptic
    const int32_t *in_rowptr; // [N+1]
    const int32_t *in_col;
    const double *in_val;
    const double *leak;       // [N]
    const int32_t *out_slot;  // [N]
    double *features;         // [B][nkeys][n_out]
    uint8_t *raster;          // optional [B][T][N]
    int B, N, n_pad, C, T, refractory, n_out, nkeys, nan_to_num, x_in_smem;
    unsigned feature_mask;
    double theta, scale;
};

template <int NPT>
__global__ void __launch_bounds__(1024) reservoir_kernel(const ResArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // [2][N] spike lists (uint16 when N <= 65536), then optionally the utterance's input spikes
    unsigned short *s_list = reinterpret_cast<unsigned short *>(smem_raw);
    uint8_t *s_x = smem_raw + ((2 * (size_t)a.N * sizeof(unsigned short) + 15) & ~(size_t)15);
    __shared__ int s_cnt[3];

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int lane = tid & 31;
    const int N = a.N, T = a.T;
    const int utt = blockIdx.x;
    const uint8_t *x = a.spikes + (size_t)utt * a.C * T;

    if (a.x_in_smem) {
        const int nwords = (a.C * T) >> 2;      // C*T is a multiple of 4 when this path is chosen
        const uint32_t *src = reinterpret_cast<const uint32_t *>(x);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_x);
        for (int i = tid; i < nwords; i += nthr) dst[i] = __ldg(src + i);
    }
    if (tid < 3) s_cnt[tid] = 0;

    // per-neuron state in registers
    double V[NPT], leak[NPT];
    int ref[NPT], cnt[NPT], sumt[NPT], first[NPT], last[NPT], s2[NPT], burst[NPT];
    int in_lo[NPT], in_hi[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
        const int i = tid + k * nthr;
        V[k] = 0.0; ref[k] = 0; cnt[k] = 0; sumt[k] = 0; first[k] = -1; last[k] = -1; s2[k] = 0; burst[k] = 0;
        leak[k] = (i < N) ? a.leak[i] : 0.0;
        in_lo[k] = (i < N) ? a.in_rowptr[i] : 0;
        in_hi[k] = (i < N) ? a.in_rowptr[i + 1] : 0;
    }
    __syncthreads();
    const uint8_t *xs = a.x_in_smem ? s_x : x;

    for (int t = 0; t < T; ++t) {
        const unsigned short *list = s_list + (t & 1) * N;
        unsigned short *list_next = s_list + ((t + 1) & 1) * N;
        const int n_prev = s_cnt[t % 3];
        if (tid == 0) s_cnt[(t + 2) % 3] = 0;

        // ---- recurrent current: exact integer sum over the neurons that fired at t-1
        int acc[NPT];
#pragma unroll
        for (int k = 0; k < NPT; ++k) acc[k] = 0;
        int q = 0;
        for (; q + 4 <= n_prev; q += 4) {
            const int32_t *r0 = a.wt + (size_t)list[q] * a.n_pad;
            const int32_t *r1 = a.wt + (size_t)list[q + 1] * a.n_pad;
            const int32_t *r2 = a.wt + (size_t)list[q + 2] * a.n_pad;
            const int32_t *r3 = a.wt + (size_t)list[q + 3] * a.n_pad;
#pragma unroll
            for (int k = 0; k < NPT; ++k) {
                const int i = tid + k * nthr;
                if (i < a.n_pad) acc[k] += (__ldg(r0 + i) + __ldg(r1 + i)) + (__ldg(r2 + i) + __ldg(r3 + i));
            }
        }
        for (; q < n_prev; ++q) {
            const int32_t *r0 = a.wt + (size_t)list[q] * a.n_pad;
#pragma unroll
            for (int k = 0; k < NPT; ++k) {
                const int i = tid + k * nthr;
                if (i < a.n_pad) acc[k] += __ldg(r0 + i);
            }
        }

        // ---- membrane update, threshold, reset, refractory (spec R6), streaming statistics (R9)
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const int i = tid + k * nthr;
            bool fire = false;
            if (i < N) {
                double i_in = 0.0;
                for (int p = in_lo[k]; p < in_hi[k]; ++p)
                    i_in = add64(i_in, mul64(a.in_val[p], (double)xs[(size_t)a.in_col[p] * T + t]));
                const double cur = add64(i_in, mul64((double)acc[k], a.scale));
                if (ref[k] == 0) {
                    double v = add64(sub64(V[k], mul64(leak[k], V[k])), cur);
                    if (v >= a.theta) { fire = true; v = 0.0; ref[k] = a.refractory; }
                    V[k] = v;
                } else {
                    V[k] = 0.0;
                    ref[k] -= 1;
                }
                if (a.raster) a.raster[((size_t)utt * T + t) * N + i] = fire ? 1 : 0;
                if (fire) {
                    if (cnt[k] > 0) {
                        const int isi = t - last[k];
                        s2[k] += isi * isi;
                        burst[k] += (isi <= a.refractory + 1) ? 1 : 0;
                    } else first[k] = t;
                    cnt[k] += 1; sumt[k] += t; last[k] = t;
                }
            }
            // ---- compaction of the firing set into next step's list (order is irrelevant: exact sums)
            const unsigned m = __ballot_sync(0xffffffffu, fire);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_cnt[(t + 1) % 3], __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (fire) list_next[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
            }
        }
        __syncthreads();
    }

    // ---- K3: feature readout, key-major [nkeys][n_out] (extract_lsm_features.py:85-87)
    if (a.features) {
        double *f = a.features + (size_t)utt * a.nkeys * a.n_out;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const int i = tid + k * nthr;
            if (i >= N) continue;
            const int o = a.out_slot[i];
            if (o < 0) continue;
            const double c = (double)cnt[k];
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            int slot = 0;
#pragma unroll
            for (int key = 0; key < 8; ++key) {
                if (!(a.feature_mask & (1u << key))) continue;
                double v = nan;
                switch (key) {
                case 0: v = c; break;
                case 1: { const double p = __ddiv_rn(c, (double)T); v = mul64(p, sub64(1.0, p)); } break;
                case 2: if (cnt[k] >= 1) v = __ddiv_rn((double)sumt[k], c); break;
                case 3: if (cnt[k] >= 1) v = (double)first[k]; break;
                case 4: if (cnt[k] >= 1) v = (double)last[k]; break;
                case 5: if (cnt[k] >= 2) v = __ddiv_rn((double)(last[k] - first[k]), (double)(cnt[k] - 1)); break;
                case 6: if (cnt[k] >= 2) {
                            const long long n = cnt[k] - 1, s1 = last[k] - first[k];
                            v = __ddiv_rn((double)(n * (long long)s2[k] - s1 * s1), (double)(n * n));
                        } break;
                case 7: v = (double)burst[k]; break;
                }
                if (a.nan_to_num && v != v) v = 0.0;
                f[(size_t)slot * a.n_out + o] = v;
                ++slot;
            }
        }
    }
}

}  // namespace

int lsm_launch_reservoir(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int B,
                         uint32_t feature_mask, int nan_to_num, double *d_features, uint8_t *d_raster,
                         cudaStream_t st)
{
    const lsm_reservoir_params &p = res->p;
    if (B <= 0) return LSM_OK;
    ResArgs a;
    a.spikes = d_spikes; a.wt = res->d_wt; a.in_rowptr = res->d_in_rowptr; a.in_col = res->d_in_col;
    a.in_val = res->d_in_val; a.leak = res->d_leak; a.out_slot = res->d_out_slot;
    a.features = d_features; a.raster = d_raster;
    a.B = B; a.N = p.num_neurons; a.n_pad = res->n_pad; a.C = p.num_inputs; a.T = p.num_steps;
    a.refractory = p.refractory; a.n_out = p.n_out; a.nan_to_num = nan_to_num;
    a.feature_mask = feature_mask & 0xFFu;
    a.nkeys = __builtin_popcount(a.feature_mask);
    a.theta = p.theta;
    a.scale = ldexp(1.0, -p.w_shift);

    const int N = p.num_neurons;
    if (N > 65536) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "num_neurons %d > 65536 not supported by the event-driven kernel", N);
    const int threads = N >= 1024 ? 1024 : ((N + 31) / 32) * 32;
    const int npt = (N + threads - 1) / threads;
    const size_t list_bytes = (2 * (size_t)N * sizeof(unsigned short) + 15) & ~(size_t)15;
    const size_t x_bytes = (size_t)p.num_inputs * p.num_steps;
    a.x_in_smem = (x_bytes % 4 == 0 && list_bytes + x_bytes <= 96 * 1024) ? 1 : 0;
    const size_t smem = list_bytes + (a.x_in_smem ? x_bytes : 0);

#define LSM_RES_LAUNCH(NPT)                                                                          \
    do {                                                                                             \
        if (smem > 48 * 1024)                                                                        \
            LSM_CUDA(ctx, cudaFuncSetAttribute(reservoir_kernel<NPT>,                                \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        reservoir_kernel<NPT><<<B, threads, smem, st>>>(a);                                          \
    } while (0)
    if (npt <= 1) LSM_RES_LAUNCH(1);
    else if (npt <= 2) LSM_RES_LAUNCH(2);
    else if (npt <= 4) LSM_RES_LAUNCH(4);
    else if (npt <= 8) LSM_RES_LAUNCH(8);
    else if (npt <= 16) LSM_RES_LAUNCH(16);
    else LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "num_neurons %d needs more than 16 neurons per thread", N);
#undef LSM_RES_LAUNCH
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return LSM_OK;
}
