// K2 + K3 — liquid-state-machine reservoir simulation with the feature readout fused in.
//
// Replaces, per utterance, /root/reference/extract_lsm_features.py:79-87:
//     lsm.reset(); lsm.set_input_spike_times(sample); lsm.simulate()      (snnpy, un-vendored)
//     lsm.extract_features_from_spikes(); key selection; np.nan_to_num
// Semantics = the frozen reservoir spec (oracle/lsm_oracle.c simulate_one).
//
// Mapping: one CTA per utterance, all T steps inside the kernel; membrane potential and refractory
// counter of every neuron live in registers for the whole simulation (thread t owns neurons t,
// t+blockDim, ...), the streaming feature statistics and the bit-packed input plane in shared
// memory, so several utterances are resident per SM.  The recurrent current is event driven:
// the neurons that fired in step t-1 are compacted (warp ballot + popc) into a shared-memory list
// and every thread adds the matching rows of the dense, presynaptic-major int32 weight plane
// (one coalesced 16-byte load per row for its 4 consecutive neurons; L2 resident: 4 MB at N = 1000).  Weights are integers (multiples of
// 2^-24), so the sum is exact in any order and bit-identical to the oracle's.
#include "reservoir_core.cuh"

namespace {

template <int NPT, bool LEAN, bool STAT_GLOBAL, bool W64 = false>
__global__ void __launch_bounds__(1024) reservoir_kernel(const ResArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_cnt[5];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int T = a.T, CW = a.CW;
    unsigned *s_bits = reinterpret_cast<unsigned *>(smem_raw);
    const int utt = blockIdx.x;
    const uint8_t *x = a.spikes + (size_t)utt * a.C * T;

    for (int i = tid; i < T * CW; i += nthr) s_bits[i] = 0u;
    __syncthreads();
    // transpose + bit-pack the input (coalesced 4-byte reads; spikes are sparse, so few atomics)
    {
        const long long nbytes = (long long)a.C * T;
        if (((nbytes & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 3) == 0)) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(x);
            for (int w = tid; w < (int)(nbytes >> 2); w += nthr) {
                const uint32_t v = __ldg(src + w);
                if (v) {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((v >> (8 * b)) & 0xffu) {
                            const int idx = 4 * w + b, c = idx / T, t = idx - c * T;
                            atomicOr(&s_bits[t * CW + (c >> 5)], 1u << (c & 31));
                        }
                }
            }
        } else {
            for (long long idx = tid; idx < nbytes; idx += nthr)
                if (x[idx]) {
                    const int c = (int)(idx / T), t = (int)(idx - (long long)c * T);
                    atomicOr(&s_bits[t * CW + (c >> 5)], 1u << (c & 31));
                }
        }
    }
    __syncthreads();
    reservoir_simulate<NPT, LEAN, STAT_GLOBAL, 0, W64>(a, utt, smem_raw, s_cnt, threadIdx.x, blockDim.x, blockIdx.x);
}

}  // namespace

void lsm_reservoir_geometry(int N, int *npt, int *threads, int *n_pad)
{
    // neurons per thread: 4 up to 4096 neurons (256-thread CTAs at N = 1000, several utterances per SM), then 8, 16
    *npt = N <= 4096 ? 4 : (N <= 8192 ? 8 : 16);
    int t = ((N + *npt - 1) / *npt + 31) / 32 * 32;
    *threads = t;
    *n_pad = t * *npt;
}

void lsm_reservoir_fill_args(const lsm_reservoir *res, const uint8_t *d_spikes, int B, uint32_t feature_mask,
                             int nan_to_num, double *d_features, uint8_t *d_raster, ResArgs *out)
{
    const lsm_reservoir_params &p = res->p;
    ResArgs &a = *out;
    a.spikes = d_spikes; a.wt = res->d_wt; a.wt64 = res->d_wt64; a.in_rowptr = res->d_in_rowptr; a.in_col = res->d_in_col;
    a.in_val = res->d_in_val; a.in_row = res->d_in_row; a.leak = res->d_leak; a.out_slot = res->d_out_slot;
    a.features = d_features; a.raster = d_raster; a.stat_global = nullptr; a.diag = nullptr;
    a.ext_id = res->d_ext_id; a.c_off = res->c_off; a.c_on = res->c_on; a.hi_magic = res->hi_magic; a.zero_row = res->zero_row;
    a.skip_dead_time = res->skip_dead_time;
    a.n_gather = res->n_gather; a.gather_row0 = res->gather_row0;
    for (int k = 0; k < 8; ++k) a.gather_out[k] = k < res->n_gather ? res->gather_out[k] : nullptr;
    a.B = B; a.N = p.num_neurons; a.n_pad = res->n_pad; a.C = p.num_inputs; a.CW = (p.num_inputs + 31) / 32; a.T = p.num_steps;
    a.refractory = p.refractory; a.n_out = p.n_out; a.nan_to_num = nan_to_num;
    a.leak0 = res->leak0; a.gain0 = res->gain0;
    a.feature_mask = feature_mask & 0xFFu;
    a.nkeys = __builtin_popcount(a.feature_mask);
    a.theta = p.theta;
    a.scale = ldexp(1.0, -p.w_shift);
}

int lsm_launch_reservoir(lsm_ctx *ctx, lsm_reservoir *res, const uint8_t *d_spikes, int B,
                         uint32_t feature_mask, int nan_to_num, double *d_features, uint8_t *d_raster,
                         cudaStream_t st, int *d_diag, long long row0)
{
    const lsm_reservoir_params &p = res->p;
    if (B <= 0) return LSM_OK;
    ResArgs a;
    lsm_reservoir_fill_args(res, d_spikes, B, feature_mask, nan_to_num, d_features, d_raster, &a);
    a.diag = d_diag;
    a.gather_row0 += row0;
    const int N = p.num_neurons;
    int npt, threads, n_pad;
    lsm_reservoir_geometry(N, &npt, &threads, &n_pad);
    if (threads > 1024) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "num_neurons %d > 16384 not supported by the event-driven kernel", N);
    size_t smem = lsm_res_smem_bytes(a.T, a.CW, threads * npt, N);
    if (smem > 200 * 1024) {
        // large reservoirs: the per-neuron statistics (touched only on a spike) move to a global scratch, one slab per CTA
        smem = lsm_res_smem_bytes(a.T, a.CW, threads * npt, N, false);
        if (npt != 16) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "reservoir with %d neurons per thread does not fit in shared memory", npt);
        if (smem > 227 * 1024) LSM_FAIL(ctx, LSM_ERR_UNSUPPORTED, "reservoir needs %zu bytes of shared memory per utterance", smem);
        void *slab;
        int rc = lsm_stage_device(ctx, 6, sizeof(int) * 6 * (size_t)threads * npt * B, &slab);
        if (rc != LSM_OK) return rc;
        a.stat_global = (int *)slab;
    }

#define LSM_RES_LAUNCH(NPT, LEAN, SG)                                                                \
    do {                                                                                             \
        if (smem > 48 * 1024)                                                                        \
            LSM_CUDA(ctx, cudaFuncSetAttribute(reservoir_kernel<NPT, LEAN, SG>,                      \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        reservoir_kernel<NPT, LEAN, SG><<<B, threads, smem, st>>>(a);                                \
    } while (0)
#define LSM_RES_LAUNCH64(NPT, SG)                                                                    \
    do {                                                                                             \
        if (smem > 48 * 1024)                                                                        \
            LSM_CUDA(ctx, cudaFuncSetAttribute(reservoir_kernel<NPT, false, SG, true>,               \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        reservoir_kernel<NPT, false, SG, true><<<B, threads, smem, st>>>(a);                         \
    } while (0)
    if (res->w64) {                 // strict reservoir: fp64 weights, ordered sums
        if (a.stat_global) LSM_RES_LAUNCH64(16, true);
        else if (npt == 4) LSM_RES_LAUNCH64(4, false);
        else if (npt == 8) LSM_RES_LAUNCH64(8, false);
        else LSM_RES_LAUNCH64(16, false);
    } else
    if (a.stat_global) {            // only the 16-neurons-per-thread shapes are large enough to need the slab
        if (res->lean) LSM_RES_LAUNCH(16, true, true);
        else LSM_RES_LAUNCH(16, false, true);
    } else if (res->lean) {
        if (npt == 4) LSM_RES_LAUNCH(4, true, false);
        else if (npt == 8) LSM_RES_LAUNCH(8, true, false);
        else LSM_RES_LAUNCH(16, true, false);
    } else {
        if (npt == 4) LSM_RES_LAUNCH(4, false, false);
        else if (npt == 8) LSM_RES_LAUNCH(8, false, false);
        else LSM_RES_LAUNCH(16, false, false);
    }
#undef LSM_RES_LAUNCH
#undef LSM_RES_LAUNCH64
    ctx->launches += 1;
    LSM_CUDA(ctx, cudaGetLastError());
    return LSM_OK;
}
