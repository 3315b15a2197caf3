// Reservoir step loop + feature readout shared by the stand-alone kernel (reservoir.cu) and the fused
// audio -> features kernel (frontend_gammatone.cu).  Semantics: DESIGN.md reservoir spec R6, R8-R10;
// oracle/lsm_oracle.c simulate_one.  Replaces /root/reference/extract_lsm_features.py:79-87 per utterance.
#pragma once

#include "lsm_common.cuh"

struct ResArgs {
    const uint8_t *spikes;    // [B][C][T]  level signal: any non-zero byte is "on"
    const int32_t *wt;        // [zero_row+1][n_pad]  row = presynaptic neuron, last row = zeros (list padding)
    const double *wt64;       // strict reservoirs: [N][n_pad] fp64 weights, same orientation
    const int32_t *in_rowptr; // [N+1]
    const int32_t *in_col;
    const double *in_val;
    const int32_t *in_row;    // [n_pad] the single input row of a neuron, -1 none, -2 several (generic CSR walk)
    const double *leak;       // [N]
    const int32_t *out_slot;  // [n_pad] position in the output list, -1 = not an output neuron / padding
    double *features;         // [B][nkeys][n_out]
    uint8_t *raster;          // optional [B][T][N]
    int *stat_global;         // per-CTA [6][slots] statistics when they do not fit in shared memory (large N), else null
    int *diag;                // optional [B][2]: neurons that fired at least once, total spikes (run_network_diagnostics)
    const int32_t *ext_id;    // [n_pad] LEAN layout: external neuron index of an internal slot (>= N for padding), else null
    double c_off, c_on;       // LEAN: cur = D(acc) - c_off (input off) / - c_on (input on), see lean_current()
    int hi_magic;             // LEAN: high word of the double whose low word holds acc ^ 0x80000000
    int zero_row;             // index of the all-zero weight row that pads a group of four list entries
    int skip_dead_time;       // 1: theta > 0 and every leak in [0, 1], so silent stretches may be skipped (exactly)
    // fused all-gather: the feature row of utterance u also goes to row gather_row0 + u of every rank's gather matrix, as
    // direct stores over NVLink into IPC-mapped peer memory (lsm_reservoir_set_gather); n_gather = 0: off
    double *gather_out[8];
    long long gather_row0;
    int n_gather;
    int B, N, n_pad, C, CW, T, refractory, n_out, nkeys, nan_to_num;
    unsigned feature_mask;
    double theta, scale, leak0, gain0;
};


// Shared-memory plan (per CTA = per utterance), S = blockDim.x * NPT neuron slots:
//   bits  uint32[T][CW]     input spikes, time-major, one bit per input row (filled by the caller)
//   stat  int32[6][S]       per-neuron count, sum t, first, last, sum isi^2, bursts
//   list  uint16[2][NL]     neurons that fired in the previous / current step (read 4 at a time; slots past
//                           the end select the all-zero weight row)
__host__ __device__ inline size_t lsm_res_smem_bytes(int T, int CW, int slots, int N, bool stat_in_smem = true)
{
    // the bit plane is rounded up to whole 8-byte words: the spike list behind it is read with 8-byte loads
    return sizeof(unsigned) * (((size_t)T * CW + 1) & ~(size_t)1) + (stat_in_smem ? sizeof(int) * 6 * (size_t)slots : 0) +
           sizeof(unsigned short) * 2 * (size_t)((N + 7) & ~3);
}

// LEAN layout and arithmetic (the reference's setup: uniform leak, one input row per driven neuron, theta > 0).
//  * Neurons are relabelled on the host so that input row r drives internal neuron 8r: the input test is then a static
//    property of a thread's slot (slot 0 of every thread at 8 neurons per thread), the other slots carry no input code.
//    Weight rows/columns, output slots and the raster map use the internal labels; padding neurons have no input and
//    all-zero weight columns, so they never leave V = 0 and need no bounds test.
//  * I = I_in + acc * 2^-w (spec R6, one rounding) without I2F and DMUL: the bits of acc ^ 0x80000000 are the low word of the
//    double D = 2^(52-w) + (acc + 2^31) * 2^-w (exact), and I = D - (2^(52-w) + 2^(31-w) - I_in) has the same exact value and
//    therefore the same single rounding; the host checks that the constant is exact for this gain.
//  * Refractory counters are 4-bit fields of one register per 8 neurons (decrement-if-nonzero in five integer operations).
__device__ __forceinline__ double lean_current(int acc, int hi_magic, double c)
{
    return __dsub_rn(__hiloint2double(hi_magic, acc ^ (int)0x80000000), c);
}

// Barrier of the thread group that simulates one utterance: the whole CTA (BAR = 0, __syncthreads) or, in the
// warp-specialised kernel, the nthr threads that own named barrier BAR.
template <int BAR>
__device__ __forceinline__ void res_sync(int nthr)
{
    if (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(BAR), "r"(nthr) : "memory");
}

// Caller contract: s_bits holds the utterance's input and a barrier of the group has made it visible.
// LEAN = uniform leak, uniform input gain, at most one input row per neuron, relabelled neurons (see above).
// STAT_GLOBAL = per-neuron statistics in a.stat_global instead of shared memory (reservoirs too large for it).
// tid / nthr: this thread's index in the group and the group's size (whole warps); slab: the group's index for STAT_GLOBAL.
// bits_ext: the input bit plane when it lives outside the group's own shared-memory plan (a buffer another thread group filled);
// the plan keeps its layout, its own bit-plane area is then simply unused.
// W64: strict reservoir (SURVEY.md 8c S3/S6 as written): fp64 weights, and the recurrent current of a neuron is their sum added one
// by one in ascending presynaptic index.  The spike list is then built in ascending neuron order (a block-wide prefix sum instead
// of per-warp atomics) and consumed one row at a time; generic layout (identity labelling) and whole-CTA groups only.
template <int NPT, bool LEAN, bool STAT_GLOBAL = false, int BAR = 0, bool W64 = false>
__device__ __forceinline__ void reservoir_simulate(const ResArgs &a, const int utt, unsigned char *smem_raw, int *s_cnt,
                                                   const int tid, const int nthr, const int slab = 0, unsigned *bits_ext = nullptr)
{
    static_assert(!W64 || (!LEAN && BAR == 0), "strict reservoirs use the generic layout and one group per CTA");
    __shared__ int s_wsum[W64 ? 32 : 1];               // W64: spikes per warp of the current step (ordered compaction)
    const int lane = tid & 31;
    const int N = a.N, T = a.T, CW = a.CW;
    const int S = nthr * NPT;
    const int NL = (N + 7) & ~3;                       // list capacity, multiple of 4
    unsigned *s_plan = reinterpret_cast<unsigned *>(smem_raw);
    unsigned *s_bits = bits_ext ? bits_ext : s_plan;
    const size_t bits_words = ((size_t)T * CW + 1) & ~(size_t)1;      // lsm_res_smem_bytes: 8-byte aligned list
    int *s_stat = STAT_GLOBAL ? a.stat_global + (size_t)slab * 6 * S : reinterpret_cast<int *>(s_plan + bits_words);
    unsigned short *s_list = reinterpret_cast<unsigned short *>(reinterpret_cast<int *>(s_plan + bits_words) +
                                                                (STAT_GLOBAL ? 0 : 6 * (size_t)S));

    for (int i = tid; i < S; i += nthr) {
        s_stat[i] = 0; s_stat[S + i] = 0; s_stat[2 * S + i] = -1; s_stat[3 * S + i] = -1; s_stat[4 * S + i] = 0; s_stat[5 * S + i] = 0;
    }
    if (tid < 3) s_cnt[tid] = 0;
    if (tid == 3) { s_cnt[3] = T; s_cnt[4] = -1; }     // first / last step with an input spike (see below)

    // per-neuron state in registers: thread tid owns the NPT consecutive neurons tid*NPT ...
    const int i0 = tid * NPT;
    double V[NPT], lk[NPT];
    int ref[LEAN ? 1 : NPT], in_word[LEAN ? 2 : NPT];
    unsigned in_mask[LEAN ? 2 : NPT];
    unsigned refn[(NPT + 7) / 8];                      // LEAN: 4-bit refractory counters, 8 neurons per word
    if (LEAN) {
#pragma unroll
        for (int k = 0; k < NPT; ++k) V[k] = 0.0;
#pragma unroll
        for (int w = 0; w < (NPT + 7) / 8; ++w) {
            refn[w] = 0u;
            const int i = i0 + 8 * w;                  // the only slots that can carry an input: internal index = 8 * row
            const int row = i >> 3;
            const bool has = ((i & 7) == 0) && row < a.C;
            in_word[w] = has ? (row >> 5) : 0;
            in_mask[w] = has ? (1u << (row & 31)) : 0u;
        }
        ref[0] = 0; lk[0] = a.leak0;
    } else {
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            const int i = i0 + k;
            V[k] = 0.0; ref[k] = 0;
            const int r = __ldg(a.in_row + i);            // padded to n_pad with -1
            in_word[k] = r >= 0 ? (r >> 5) : (r == -2 ? -2 : 0);
            in_mask[k] = r >= 0 ? (1u << (r & 31)) : 0u;
            lk[k] = i >= N ? a.leak0 : __ldg(a.leak + i);
        }
    }
    const unsigned pitch = (unsigned)a.n_pad * 4u;
    res_sync<BAR>(nthr);

    // Dead time.  Until the first input spike every membrane potential is exactly +0, nobody fires and no statistic changes
    // ((0 - leak*0) + 0 = +0), so the simulation starts at that step.  And once the input has ended and a step passes without
    // a spike, no neuron can fire again (without current |V| only shrinks: 0 <= leak <= 1, theta > 0 - checked on the host), so
    // the rasters and statistics are final and the loop stops.  Leading and trailing silence of an utterance cost nothing; the
    // result is exactly that of simulating all T steps.  (With a raster dump all steps are simulated, which writes their zeros.)
    int t_first = 0, t_last_in = T;
    if (!a.raster && a.skip_dead_time) {
        for (int i = tid; i < T * CW; i += nthr)
            if (s_bits[i]) { atomicMin(&s_cnt[3], i / CW); atomicMax(&s_cnt[4], i / CW); }
        res_sync<BAR>(nthr);
        t_first = s_cnt[3];
        t_last_in = s_cnt[4];
    }

    int c_cur = 0, c_nxt = 1, c_zero = 2;
    for (int t = t_first; t < T; ++t) {
        const unsigned short *list = s_list + (t & 1) * NL;
        unsigned short *list_next = s_list + ((t + 1) & 1) * NL;
        const int n_prev = s_cnt[c_cur];
        if (t > t_last_in && n_prev == 0) break;      // input over and the reservoir has fallen silent (uniform across the CTA)
        if (tid == 0) s_cnt[c_zero] = 0;

        // ---- recurrent current: exact integer sum over the neurons that fired at t-1; one 16-byte
        //      load per (presynaptic row, 4 consecutive postsynaptic neurons)
        int acc[NPT];
        double accd[W64 ? NPT : 1];
#pragma unroll
        for (int k = 0; k < NPT; ++k) acc[k] = 0;
        if (W64) {
#pragma unroll
            for (int k = 0; k < NPT; ++k) accd[k] = 0.0;
#pragma unroll 1
            for (int q = 0; q < n_prev; ++q) {                       // ascending presynaptic index: one rounded addition per spike
                const double2 *row = reinterpret_cast<const double2 *>(a.wt64 + (size_t)list[q] * a.n_pad + i0);
#pragma unroll
                for (int u = 0; u < NPT / 2; ++u) {
                    const double2 w = __ldg(row + u);
                    accd[2 * u] = add64(accd[2 * u], w.x);
                    accd[2 * u + 1] = add64(accd[2 * u + 1], w.y);
                }
            }
        }
#pragma unroll 1   // keep the step loop small: the unrolled form (16 rows per trip) made instruction fetch a top stall
        for (int q = 0; !W64 && q < n_prev; q += 4) {
            const uint2 jj = *reinterpret_cast<const uint2 *>(list + q);
            // entries past the end of the list select the all-zero weight row (index N)
            const int j0 = jj.x & 0xffff;
            const int j1 = (q + 1 < n_prev) ? (int)(jj.x >> 16) : a.zero_row;
            const int j2 = (q + 2 < n_prev) ? (int)(jj.y & 0xffff) : a.zero_row;
            const int j3 = (q + 3 < n_prev) ? (int)(jj.y >> 16) : a.zero_row;
            // 32-bit int4 indices off the uniform base pointer: one IMAD per row instead of 64-bit address arithmetic
            const int4 *w4 = reinterpret_cast<const int4 *>(a.wt);
            const unsigned p4 = pitch >> 4, t4 = (unsigned)i0 >> 2;
            const int4 *r0 = w4 + ((unsigned)j0 * p4 + t4), *r1 = w4 + ((unsigned)j1 * p4 + t4);
            const int4 *r2 = w4 + ((unsigned)j2 * p4 + t4), *r3 = w4 + ((unsigned)j3 * p4 + t4);
#pragma unroll
            for (int u = 0; u < NPT / 4; ++u) {
                const int4 w0 = __ldg(r0 + u);
                const int4 w1 = __ldg(r1 + u);
                const int4 w2 = __ldg(r2 + u);
                const int4 w3 = __ldg(r3 + u);
                acc[4 * u + 0] += (w0.x + w1.x) + (w2.x + w3.x);
                acc[4 * u + 1] += (w0.y + w1.y) + (w2.y + w3.y);
                acc[4 * u + 2] += (w0.z + w1.z) + (w2.z + w3.z);
                acc[4 * u + 3] += (w0.w + w1.w) + (w2.w + w3.w);
            }
        }
        // ---- membrane update, threshold, reset, refractory (spec R6), branch-free
        const unsigned *bits_t = s_bits + t * CW;
        unsigned fired = 0;                            // bit k: neuron i0+k fired
        if (LEAN) {
#pragma unroll
            for (int w = 0; w < (NPT + 7) / 8; ++w) {
                // nz: bit 4j+3 set iff counter j of this word is non-zero (the neuron sits out this step); then count down
                const unsigned r = refn[w];
                const unsigned nz = (((r & 0x77777777u) + 0x77777777u) | r) & 0x88888888u;
                refn[w] = r - (nz >> 3);
                const bool in_on = (bits_t[in_word[w]] & in_mask[w]) != 0u;
                unsigned fire_n = 0u;                  // bit 4j: neuron 8w+j fires now
#pragma unroll
                for (int j = 0; j < 8 && 8 * w + j < NPT; ++j) {
                    const int k = 8 * w + j;
                    const double cur = lean_current(acc[k], a.hi_magic, (j == 0 && in_on) ? a.c_on : a.c_off);
                    const double v = add64(sub64(V[k], mul64(a.leak0, V[k])), cur);
                    const bool active = (nz & (8u << (4 * j))) == 0u;
                    const bool ge = v >= a.theta;              // one comparison; fire and keep are predicate logic on it
                    const bool fire = active && ge;
                    const bool keep = active && !ge;
                    V[k] = keep ? v : 0.0;
                    fire_n |= fire ? (1u << (4 * j)) : 0u;
                    fired |= fire ? (1u << k) : 0u;
                }
                refn[w] += fire_n * (unsigned)a.refractory;
            }
        } else {
#pragma unroll
            for (int k = 0; k < NPT; ++k) {
                const int i = i0 + k;
                double i_in = 0.0;
                if (in_word[k] >= 0) {
                    if (bits_t[in_word[k]] & in_mask[k]) i_in = add64(0.0, __ldg(a.in_val + __ldg(a.in_rowptr + i)));
                } else {
                    for (int p = __ldg(a.in_rowptr + i); p < __ldg(a.in_rowptr + i + 1); ++p) {
                        const int rr = __ldg(a.in_col + p);
                        const double on = ((bits_t[rr >> 5] >> (rr & 31)) & 1u) ? 1.0 : 0.0;
                        i_in = add64(i_in, mul64(__ldg(a.in_val + p), on));
                    }
                }
                const double cur = W64 ? add64(i_in, accd[k]) : add64(i_in, mul64((double)acc[k], a.scale));
                const double v = add64(sub64(V[k], mul64(lk[k], V[k])), cur);
                const bool active = ref[k] == 0;
                const bool fire = active && (v >= a.theta) && (i < N);
                V[k] = (active && !fire) ? v : 0.0;
                ref[k] = fire ? a.refractory : (active ? 0 : ref[k] - 1);
                fired |= fire ? (1u << k) : 0u;
            }
        }
        if (a.raster) {
#pragma unroll
            for (int k = 0; k < NPT; ++k) {
                const int e = LEAN ? __ldg(a.ext_id + i0 + k) : i0 + k;     // external (caller's) neuron index
                if (e < N) a.raster[((size_t)utt * T + t) * N + e] = (fired >> k) & 1u;
            }
        }
        // ---- spikes are rare: statistics (spec R9) and list compaction only for warps that have one
        if (__any_sync(0xffffffffu, fired != 0u)) {
#pragma unroll
            for (int k = 0; k < NPT; ++k) {
                const bool fire = (fired >> k) & 1u;
                if (fire) {
                    const int sl = k * nthr + tid;
                    const int c = s_stat[sl];
                    if (c > 0) {
                        const int isi = t - s_stat[3 * S + sl];
                        s_stat[4 * S + sl] += isi * isi;
                        if (isi <= a.refractory + 1) s_stat[5 * S + sl] += 1;
                    } else s_stat[2 * S + sl] = t;
                    s_stat[sl] = c + 1;
                    s_stat[S + sl] += t;
                    s_stat[3 * S + sl] = t;
                }
                const unsigned m = W64 ? 0u : __ballot_sync(0xffffffffu, fire);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_cnt[c_nxt], __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (fire) list_next[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(i0 + k);
                }
            }
        }
        if (W64) {
            // ordered compaction: position = spikes of lower threads + this thread's lower neurons (neuron index = tid * NPT + k)
            const int mine = __popc(fired);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_wsum[tid >> 5] = incl;
            res_sync<BAR>(nthr);
            int base = 0;
            for (int w = 0; w < (tid >> 5); ++w) base += s_wsum[w];
            int pos = base + incl - mine;
#pragma unroll
            for (int k = 0; k < NPT; ++k)
                if ((fired >> k) & 1u) list_next[pos++] = (unsigned short)(i0 + k);
            if (tid == nthr - 1) s_cnt[c_nxt] = base + incl;
        }
        res_sync<BAR>(nthr);
        const int tmp = c_cur; c_cur = c_nxt; c_nxt = c_zero; c_zero = tmp;
    }

    // ---- network diagnostics (extract_lsm_features.py:117-131) over ALL neurons: participation and activity
    if (a.diag) {
        int active = 0, total = 0;
#pragma unroll
        for (int k = 0; k < NPT; ++k) {
            { const int c = s_stat[k * nthr + tid]; active += c > 0 ? 1 : 0; total += c; }      // padding neurons never fire
        }
        active = __reduce_add_sync(0xffffffffu, active);
        total = __reduce_add_sync(0xffffffffu, total);
        if (tid < 2) s_cnt[tid] = 0;
        res_sync<BAR>(nthr);
        if (lane == 0) { atomicAdd(&s_cnt[0], active); atomicAdd(&s_cnt[1], total); }
        res_sync<BAR>(nthr);
        if (tid < 2) a.diag[(size_t)utt * 2 + tid] = s_cnt[tid];
    }

    // ---- K3: feature readout, key-major [nkeys][n_out] (extract_lsm_features.py:85-87).  Each key's values are
    //      gathered in shared memory (the input-bit plane is free now) and written as one contiguous, coalesced
    //      run: full 128-byte lines whether the destination is HBM or pinned host memory across PCIe.
    if (a.features) {
        double *f = a.features + (size_t)utt * a.nkeys * a.n_out;
        double *s_buf = reinterpret_cast<double *>(s_bits);
        const int cap = (T * CW * (int)sizeof(unsigned)) / (int)sizeof(double);   // doubles that fit in the bit plane
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        int o_k[NPT];
#pragma unroll
        for (int k = 0; k < NPT; ++k) o_k[k] = __ldg(a.out_slot + i0 + k);      // [n_pad], -1 for padding and non-output neurons
        int slot = 0;
        for (int key = 0; key < 8; ++key) {
            if (!(a.feature_mask & (1u << key))) continue;
            for (int tile = 0; tile < a.n_out; tile += cap) {
#pragma unroll
                for (int k = 0; k < NPT; ++k) {
                    const int o = o_k[k];
                    if (o < tile || o >= tile + cap) continue;
                    const int sl = k * nthr + tid;
                    const int cnt = s_stat[sl], sumt = s_stat[S + sl], first = s_stat[2 * S + sl], last = s_stat[3 * S + sl];
                    const int s2 = s_stat[4 * S + sl], burst = s_stat[5 * S + sl];
                    const double c = (double)cnt;
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
                    if (a.nan_to_num && v != v) v = 0.0;
                    s_buf[o - tile] = v;
                }
                res_sync<BAR>(nthr);
                const int n_here = min(cap, a.n_out - tile);
                for (int idx = tid; idx < n_here; idx += nthr) __stcs(f + (size_t)slot * a.n_out + tile + idx, s_buf[idx]);
                if (a.n_gather > 0) {
                    // the collective, fused: the same coalesced run into every rank's gather matrix (own one included)
                    const size_t off = ((size_t)(a.gather_row0 + utt) * a.nkeys + slot) * a.n_out + tile;
                    for (int p = 0; p < a.n_gather; ++p)
                        for (int idx = tid; idx < n_here; idx += nthr) __stcs(a.gather_out[p] + off + idx, s_buf[idx]);
                }
                res_sync<BAR>(nthr);
            }
            ++slot;
        }
    }
}
