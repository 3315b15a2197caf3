/*
 * TEST INFRASTRUCTURE — plain-C CPU restatement of the reference's audio -> spikes -> LSM ->
 * features path.  Never linked into or called by the product library; only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() load it.
 *
 * PARITY STATUS: "parity unpinned" for the three third-party packages the reference delegates
 * to (gammatone==1.0.3, librosa==0.11.0, snn_reservoir_py==2.0.0: not vendored, not installable,
 * the reference has no tests).  Pinned pieces: the IIR recurrence (bit-exact vs scipy.signal.lfilter),
 * the window mean (numpy, F-ordered segment => sequential sum), zoom (bit-exact vs
 * scipy.ndimage.zoom order=1), the hysteresis encoder and w_critico (bit-exact vs the reference's
 * own functions run verbatim) — see tests/test_oracle_*.py and tests/golden/make_golden.py.
 *
 * Every floating-point result here is a fixed sequence of IEEE-754 binary64/binary32
 * +,-,*,/,sqrt operations (compile with -ffp-contract=off, no -ffast-math), so it is the same on
 * any machine and the CUDA kernels can reproduce it bit for bit with __dadd_rn/__dmul_rn.
 * The single deviation from "what numpy would do" is log10: numpy calls the platform libm
 * (or SVML), whose last bit differs between machines; we use the fdlibm/FreeBSD algorithm
 * below (< 1 ulp) on both CPU and GPU.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -pthread -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* Utterances are independent (create_dataset.py:143, extract_lsm_features.py:78 are plain
 * per-sample loops), so the multi-core baseline is a dynamic parallel-for over utterances. */
typedef void (*utt_fn)(void *ctx, int b, void *scratch);
typedef struct {
    utt_fn fn; void *ctx; int B; atomic_int next;
    void *(*mk)(void *ctx); void (*rm)(void *scratch);
} pf_job;

static void *pf_worker(void *arg)
{
    pf_job *j = (pf_job *)arg;
    void *scratch = j->mk(j->ctx);
    for (;;) {
        int b = atomic_fetch_add(&j->next, 1);
        if (b >= j->B) break;
        j->fn(j->ctx, b, scratch);
    }
    j->rm(scratch);
    return NULL;
}

static void parallel_for(pf_job *j, int nthreads)
{
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads > j->B) nthreads = j->B;
    if (nthreads < 1) nthreads = 1;
    atomic_store(&j->next, 0);
    if (nthreads == 1) { pf_worker(j); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, pf_worker, j);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
}

/* ------------------------------------------------------------------------------------------
 * Deterministic log10 for finite x > 0 (normal numbers; the path only feeds it sqrt(mean)+1e-9).
 * fdlibm e_log.c kernel + FreeBSD e_log10.c hi/lo recombination, basic operations only.      */
static const double
    LG1 = 6.666666666666735130e-01, LG2 = 3.999999999940941908e-01,
    LG3 = 2.857142874366239149e-01, LG4 = 2.222219843214978396e-01,
    LG5 = 1.818357216161805012e-01, LG6 = 1.531383769920937332e-01,
    LG7 = 1.479819860511658591e-01,
    IVLN10HI = 4.34294481878168880939e-01,  /* 0x3fdbcb7b15200000 */
    IVLN10LO = 2.50829467116452752298e-11,  /* 0x3dbb9438ca9aadd5 */
    LOG10_2HI = 3.01029995663611771306e-01, /* 0x3FD34413509F6000 */
    LOG10_2LO = 3.69423907715893078616e-13; /* 0x3D59FEF311F12B36 */

double oracle_log10(double x)
{
    uint64_t ix;
    memcpy(&ix, &x, 8);
    int32_t hx = (int32_t)(ix >> 32);
    int32_t k = 0;
    if (hx < 0x00100000) {           /* subnormal: scale up (not reached on this path) */
        x *= 18014398509481984.0;    /* 2**54 */
        memcpy(&ix, &x, 8);
        hx = (int32_t)(ix >> 32);
        k -= 54;
    }
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    /* normalise x or x/2 so that sqrt(2)/2 < x < sqrt(2) */
    ix = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (ix & 0xffffffffu);
    memcpy(&x, &ix, 8);
    k += (i >> 20);
    double f = x - 1.0;
    double hfsq = 0.5 * f * f;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (LG2 + w * (LG4 + w * LG6));
    double t2 = z * (LG1 + w * (LG3 + w * (LG5 + w * LG7)));
    double R = t2 + t1;
    double hi = f - hfsq;
    uint64_t ih;
    memcpy(&ih, &hi, 8);
    ih &= 0xffffffff00000000ull;
    memcpy(&hi, &ih, 8);
    double lo = (f - hi) - hfsq + s * (hfsq + R);
    double val_hi = hi * IVLN10HI;
    double dk = (double)k;
    double y2 = dk * LOG10_2HI;
    double val_lo = dk * LOG10_2LO + (lo + hi) * IVLN10LO + lo * IVLN10HI;
    double ww = y2 + val_hi;
    val_lo += (y2 - ww) + val_hi;
    val_hi = ww;
    return val_lo + val_hi;
}

/* ------------------------------------------------------------------------------------------
 * Stage 1, gammatone branch: create_dataset.py:49-78 + gammatone.gtgram.
 * coefs: [C][10] rows {A0,A11,A12,A13,A14,A2,B0,B1,B2,gain}, row 0 = lowest centre frequency.
 * spec: out [C][ncols] = sqrt(mean(window of (y4/gain)^2))                                   */
static void gammatone_energy(const float *pcm, int L, const double *coefs, int C,
                             int nwin, int hop, int ncols, double *spec)
{
    for (int ch = 0; ch < C; ++ch) {
        const double *c = coefs + 10 * ch;
        /* scipy lfilter normalises by a[0] = B0 first (linear_filter: "ptr_b[n] /= a0") */
        const double a0 = c[6];
        const double b0 = c[0] / a0, b2 = c[5] / a0, a1 = c[7] / a0, a2 = c[8] / a0;
        const double b1[4] = {c[1] / a0, c[2] / a0, c[3] / a0, c[4] / a0};
        const double gain = c[9];
        double z0[4] = {0, 0, 0, 0}, z1[4] = {0, 0, 0, 0};
        /* up to ceil(nwin/hop) windows are open at any sample */
        double acc[8];
        int nopen = (nwin + hop - 1) / hop;
        if (nopen > 8) nopen = 8;
        for (int q = 0; q < 8; ++q) acc[q] = 0.0;
        const int last = (ncols - 1) * hop + nwin; /* samples beyond this are never read */
        for (int n = 0; n < last && n < L; ++n) {
            double x = (double)pcm[n];
            for (int st = 0; st < 4; ++st) {
                /* scipy.signal.lfilter, direct form II transposed, len(b)=len(a)=3 */
                double y = z0[st] + b0 * x;
                z0[st] = (z1[st] + x * b1[st]) - y * a1;
                z1[st] = x * b2 - y * a2;
                x = y;
            }
            double v = x / gain;
            double e = v * v;
            /* window c covers [c*hop, c*hop+nwin); window c lives in slot c % nopen */
            int cfirst = (n - nwin + hop) / hop; /* smallest c with c*hop + nwin > n */
            if (n - nwin + hop < 0) cfirst = 0;
            int clast = n / hop;
            if (clast > ncols - 1) clast = ncols - 1;
            for (int cw = cfirst; cw <= clast; ++cw) {
                int slot = cw % nopen;
                if (n == cw * hop) acc[slot] = 0.0 + e;   /* np.add.reduce starts from the first element */
                else acc[slot] = acc[slot] + e;
                if (n == cw * hop + nwin - 1)
                    spec[(size_t)ch * ncols + cw] = sqrt(acc[slot] / (double)nwin);
            }
        }
    }
}

/* The product's SPECULATIVE arrangement of the same cascade (csrc/gammatone_core.cuh gt_filter_fast), restated with the
 * same fused multiply-adds so that tests/test_error_bound.py can check, without a GPU, that its window amplitudes stay
 * within the derived bound (csrc/error_bound.cu) of the reference-order ones above.  Test infrastructure like the rest
 * of this file.  amp_exact / amp_fast: [C][ncols].                                                                   */
void oracle_gammatone_two_arrangements(const float *pcm, int L, const double *coefs, int C, int nwin, int hop,
                                       int ncols, double *amp_exact, double *amp_fast)
{
    gammatone_energy(pcm, L, coefs, C, nwin, hop, ncols, amp_exact);
    const int r_old = nwin - 2 * hop;
    const int n_used = (ncols - 1) * hop + nwin;
    const int n_blocks = (n_used + hop - 1) / hop;
    for (int ch = 0; ch < C; ++ch) {
        const double *c = coefs + 10 * ch;
        const double A0 = c[0], a0 = c[6];
        const double c1 = c[1] / A0, c2 = c[2] / A0, c3 = c[3] / A0, c4 = c[4] / A0;
        const double na1 = -(c[7] / a0), na2 = -(c[8] / a0);
        const double sc = c[0] / c[6];
        const double G = (sc * sc) * (sc * sc) / c[9];
        const double g2n = G * G / (double)nwin;
        double xp = 0, p1 = 0, q1 = 0, p2 = 0, q2 = 0, p3 = 0, q3 = 0, p4 = 0, q4 = 0;
        double acc = 0, full1 = 0, full2 = 0;
        for (int m = 0; m < n_blocks; ++m) {
            const int n_here = (n_used - m * hop) < hop ? (n_used - m * hop) : hop;
            acc = 0.0;
            for (int p = 0; p < n_here; ++p) {
                if (p == r_old && m >= 2) amp_fast[(size_t)ch * ncols + (m - 2)] = sqrt(((full2 + full1) + acc) * g2n);
                const int n = m * hop + p;
                const double x = n < L ? (double)pcm[n] : 0.0;
                /* unskewed evaluation: the same values as the skewed loop of the kernel, stage by stage.  Stage k consumes
                 * the stage k-1 outputs of this sample and of the previous one (o1..o3 = p1..p3 before their update). */
                const double n1 = fma(na1, p1, fma(na2, q1, fma(c1, xp, x)));
                const double o1 = p1;
                xp = x; q1 = p1; p1 = n1;
                const double n2 = fma(na1, p2, fma(na2, q2, fma(c2, o1, n1)));
                const double o2 = p2;
                q2 = p2; p2 = n2;
                const double n3 = fma(na1, p3, fma(na2, q3, fma(c3, o2, n2)));
                const double o3 = p3;
                q3 = p3; p3 = n3;
                const double n4 = fma(na1, p4, fma(na2, q4, fma(c4, o3, n3)));
                q4 = p4; p4 = n4;
                acc = fma(n4, n4, acc);
            }
            if (n_here == r_old && m >= 2) amp_fast[(size_t)ch * ncols + (m - 2)] = sqrt(((full2 + full1) + acc) * g2n);
            full2 = full1;
            full1 = acc;
        }
    }
}

/* create_dataset.py:59-78 on a [C][ncols] energy matrix -> spec_norm [C][nbins] (fp64).
 * zi0/zf: zoom table (source index, fraction) for each output bin.  Returns 0 if degenerate
 * (all-zero output, create_dataset.py:64-65).                                                 */
static int db_normalise_zoom(const double *spec, int C, int ncols, int nbins,
                             const int32_t *zi0, const double *zf, double *db, double *norm)
{
    double mx = -INFINITY;
    for (int i = 0; i < C * ncols; ++i) {
        db[i] = 20.0 * oracle_log10(spec[i] + 1e-9);
        if (db[i] > mx) mx = db[i];
    }
    const double floor_db = mx - 80.0;
    double mn = INFINITY;
    for (int i = 0; i < C * ncols; ++i) {
        if (db[i] < floor_db) db[i] = floor_db;
        if (db[i] < mn) mn = db[i];
    }
    if ((mx - mn) < 1e-8) {
        for (int i = 0; i < C * nbins; ++i) norm[i] = 0.0;
        return 0;
    }
    const double den = mx - mn + 1e-8;
    for (int i = 0; i < C * ncols; ++i) db[i] = (db[i] - mn) / den;
    for (int ch = 0; ch < C; ++ch) {
        const double *row = db + (size_t)ch * ncols;
        for (int j = 0; j < nbins; ++j) {
            if (ncols == nbins) { norm[(size_t)ch * nbins + j] = row[j]; continue; }
            int i0 = zi0[j];
            double f = zf[j];
            double v = row[i0] * (1.0 - f);
            if (i0 + 1 < ncols) v = v + row[i0 + 1] * f;
            norm[(size_t)ch * nbins + j] = v;
        }
    }
    return 1;
}

/* create_dataset.py:81-98 (+ :101-104 redundancy).  thr[] already sorted DESCENDING, lower[k] =
 * thr[k] - gap computed by the caller in fp64.  spikes: [C*R][nbins*K]                        */
static void hysteresis_encode_f64(const double *norm, int C, int nbins, const double *thr,
                                  const double *lower, int K, int R, uint8_t *spikes)
{
    const int T = nbins * K;
    for (int ch = 0; ch < C; ++ch) {
        uint8_t *row0 = spikes + (size_t)ch * R * T;
        for (int k = 0; k < K; ++k) {
            int on = 0;
            for (int b = 0; b < nbins; ++b) {
                double v = norm[(size_t)ch * nbins + b];
                if (!on && v > thr[k]) on = 1;
                else if (on && v < lower[k]) on = 0;
                row0[b * K + k] = (uint8_t)on;
            }
        }
        for (int r = 1; r < R; ++r) memcpy(row0 + (size_t)r * T, row0, (size_t)T);
    }
}

/* Checker for the GPU's constant-divisor division (csrc/frontend_gammatone.cu div_by_const):
 * counts how many of the n quotients q = fma(fma(-g, x*r, x), r, x*r), r = 1/g, differ from x/g. */
int64_t oracle_check_const_division(const double *x, int64_t n, double g)
{
    const double r = 1.0 / g;
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double q0 = x[i] * r;
        const double e = fma(-g, q0, x[i]);
        const double q = fma(e, r, q0);
        const double want = x[i] / g;
        if (memcmp(&q, &want, 8) != 0) ++bad;
    }
    return bad;
}

/* Encoder alone (for the known-answer tests minted from the reference's own function). */
void oracle_hysteresis_encode(const double *norm, int C, int nbins, const double *thr_desc,
                              const double *lower, int K, int R, uint8_t *spikes)
{
    hysteresis_encode_f64(norm, C, nbins, thr_desc, lower, K, R, spikes);
}

/* Batched stage 1 (gammatone).  Returns 0.  spec_norm_out may be NULL.  nthreads <= 0: all cores. */
typedef struct {
    const float *pcm; int L; const double *coefs; int C, nwin, hop, ncols, nbins;
    const int32_t *zi0; const double *zf; const double *thr, *lower; int K, R;
    uint8_t *spikes; double *spec_norm_out;
} gt_ctx;

static void *gt_mk(void *vc)
{
    gt_ctx *c = (gt_ctx *)vc;
    return malloc(sizeof(double) * ((size_t)2 * c->C * c->ncols + (size_t)c->C * c->nbins));
}

static void gt_one(void *vc, int b, void *scratch)
{
    gt_ctx *c = (gt_ctx *)vc;
    double *spec = (double *)scratch, *db = spec + (size_t)c->C * c->ncols, *norm = db + (size_t)c->C * c->ncols;
    gammatone_energy(c->pcm + (size_t)b * c->L, c->L, c->coefs, c->C, c->nwin, c->hop, c->ncols, spec);
    db_normalise_zoom(spec, c->C, c->ncols, c->nbins, c->zi0, c->zf, db, norm);
    hysteresis_encode_f64(norm, c->C, c->nbins, c->thr, c->lower, c->K, c->R,
                          c->spikes + (size_t)b * c->C * c->R * c->nbins * c->K);
    if (c->spec_norm_out)
        memcpy(c->spec_norm_out + (size_t)b * c->C * c->nbins, norm, sizeof(double) * c->C * c->nbins);
}

int oracle_gammatone_encode(const float *pcm, int B, int L, const double *coefs, int C,
                            int nwin, int hop, int nbins, const int32_t *zi0, const double *zf,
                            const double *thr_desc, const double *lower, int K, int R,
                            uint8_t *spikes, double *spec_norm_out, int nthreads)
{
    gt_ctx c = {pcm, L, coefs, C, nwin, hop, 1 + (L - nwin) / hop, nbins, zi0, zf, thr_desc, lower, K, R,
                spikes, spec_norm_out};
    pf_job j = {gt_one, &c, B, 0, gt_mk, free};
    parallel_for(&j, nthreads);
    return 0;
}


/* ------------------------------------------------------------------------------------------
 * Stage 1, mel branch: create_dataset.py:43-48 + :62-78 (librosa==0.11.0 melspectrogram + power_to_db,
 * SURVEY Appendix B.2), with numpy 1.26 (the reference's pin) scalar-promotion semantics where they matter.
 *
 * STFT: n_fft = 2048, hop 160, centre zero padding, periodic hann in fp64 times float32 samples => fp64
 * frame; fp64 FFT; result stored as complex64 (librosa's dtype for float32 input).  librosa's FFT is
 * pocketfft, whose internal operation order is not reproducible here; this oracle fixes ONE order
 * (radix-2 decimation in time on the 1024-point complex packing of the real frame, then the real-input
 * untangle, twiddles from a table) that the CUDA kernel repeats operation for operation.  The two differ
 * from pocketfft at the 1e-16 level before the rounding to float32 (tests compare against scipy.fft).
 * tw:  double[512][2]  exp(-2*pi*i*q/1024), q = 0..511
 * tw2: double[1025][2] exp(-2*pi*i*k/2048), k = 0..1024
 * win: double[2048]    periodic hann                                                              */
static void fft1024(double *re, double *im, const double *tw)
{
    /* in: bit-reversed order already applied by the caller; 10 radix-2 DIT stages */
    for (int s = 1; s <= 10; ++s) {
        const int m = 1 << s, half = m >> 1, stride = 1024 >> s;
        for (int base = 0; base < 1024; base += m)
            for (int j = 0; j < half; ++j) {
                const double wr = tw[2 * (j * stride)], wi = tw[2 * (j * stride) + 1];
                const int p = base + j, q = p + half;
                const double tr = wr * re[q] - wi * im[q];
                const double ti = wr * im[q] + wi * re[q];
                const double ur = re[p], ui = im[p];
                re[p] = ur + tr; im[p] = ui + ti;
                re[q] = ur - tr; im[q] = ui - ti;
            }
    }
}

static inline int bitrev10(int x)
{
    int r = 0;
    for (int b = 0; b < 10; ++b) r |= ((x >> b) & 1) << (9 - b);
    return r;
}

/* power spectrum of one frame -> S[1025] float32 */
static void frame_power(const float *pcm, int L, int start, const double *win, const double *tw, const double *tw2,
                        double *re, double *im, float *S)
{
    for (int j = 0; j < 1024; ++j) {
        const int n0 = 2 * j, n1 = 2 * j + 1;
        const int i0 = start + n0, i1 = start + n1;
        const double x0 = (i0 >= 0 && i0 < L) ? win[n0] * (double)pcm[i0] : 0.0;
        const double x1 = (i1 >= 0 && i1 < L) ? win[n1] * (double)pcm[i1] : 0.0;
        const int r = bitrev10(j);
        re[r] = x0; im[r] = x1;
    }
    fft1024(re, im, tw);
    for (int k = 0; k <= 1024; ++k) {
        const int k1 = k & 1023, k2 = (1024 - k) & 1023;
        const double zr = re[k1], zi = im[k1], cr = re[k2], ci = -im[k2];   /* Z[k], conj(Z[N-k]) */
        const double ar = 0.5 * (zr + cr), ai = 0.5 * (zi + ci);
        const double br = 0.5 * (zr - cr), bi = 0.5 * (zi - ci);
        const double wr = tw2[2 * k], wi = tw2[2 * k + 1];
        const double c_re = wi, c_im = -wr;                                 /* -i * W */
        const double xr = ar + (c_re * br - c_im * bi);
        const double xi = ai + (c_re * bi + c_im * br);
        const float r32 = (float)xr, i32 = (float)xi;                       /* complex64 STFT */
        const float mag = (float)sqrt((double)r32 * (double)r32 + (double)i32 * (double)i32);  /* np.abs = hypotf */
        S[k] = mag * mag;                                                   /* ** 2.0 in float32 */
    }
}

static float oracle_log10f(float x) { return (float)oracle_log10((double)x); }

/* One utterance.  mel_w/lo/n/off: packed non-zero mel weights (float32), ascending bins.
 * norm: out float32[C][nbins].  Returns 0 if degenerate.                                          */
static int mel_one(const float *pcm, int L, int n_fft, int hop, int C, int ncols, int nbins,
                   const double *win, const double *tw, const double *tw2,
                   const float *mel_w, const int32_t *mel_lo, const int32_t *mel_n, const int32_t *mel_off,
                   const int32_t *zi0, const double *zf, double *re, double *im, float *S, float *M, float *norm)
{
    (void)n_fft;
    for (int t = 0; t < ncols; ++t) {
        frame_power(pcm, L, t * hop - 1024, win, tw, tw2, re, im, S);
        for (int m = 0; m < C; ++m) {
            float acc = 0.0f;
            const float *w = mel_w + mel_off[m];
            for (int q = 0; q < mel_n[m]; ++q) {
                const float prod = w[q] * S[mel_lo[m] + q];
                acc = acc + prod;
            }
            M[(size_t)m * ncols + t] = acc;
        }
    }
    /* power_to_db(ref=np.max, amin=1e-10, top_db=80) */
    float ref = M[0];
    for (int i = 1; i < C * ncols; ++i) if (M[i] > ref) ref = M[i];
    const float amin = 1e-10f;
    const double refd = ((double)ref > 1e-10) ? (double)ref : 1e-10;          /* scalar path: float64 in numpy 1.26 */
    const float ref_db = (float)(10.0 * oracle_log10(refd));
    float mx = -INFINITY;
    for (int i = 0; i < C * ncols; ++i) {
        const float v = M[i] > amin ? M[i] : amin;
        float d = 10.0f * oracle_log10f(v);
        d = d - ref_db;
        M[i] = d;
        if (d > mx) mx = d;
    }
    const float floor_db = (float)((double)mx - 80.0);
    float mn = INFINITY;
    for (int i = 0; i < C * ncols; ++i) {
        if (M[i] < floor_db) M[i] = floor_db;
        if (M[i] < mn) mn = M[i];
    }
    /* create_dataset.py:62-67 */
    const float diff = mx - mn;
    if ((double)diff < 1e-8) {
        for (int i = 0; i < C * nbins; ++i) norm[i] = 0.0f;
        return 0;
    }
    const float den = (float)((double)diff + 1e-8);
    for (int i = 0; i < C * ncols; ++i) M[i] = (M[i] - mn) / den;
    /* zoom (create_dataset.py:69-78): float64 arithmetic, float32 result */
    for (int ch = 0; ch < C; ++ch) {
        const float *row = M + (size_t)ch * ncols;
        for (int j = 0; j < nbins; ++j) {
            if (ncols == nbins) { norm[(size_t)ch * nbins + j] = row[j]; continue; }
            const int i0 = zi0[j];
            const double f = zf[j];
            double v = (double)row[i0] * (1.0 - f);
            if (i0 + 1 < ncols) v = v + (double)row[i0 + 1] * f;
            norm[(size_t)ch * nbins + j] = (float)v;
        }
    }
    return 1;
}

/* create_dataset.py:81-98 on a float32 spectrogram: comparisons in float32 against float32-rounded bounds */
static void hysteresis_encode_f32(const float *norm, int C, int nbins, const double *thr,
                                  const double *lower, int K, int R, uint8_t *spikes)
{
    const int T = nbins * K;
    for (int ch = 0; ch < C; ++ch) {
        uint8_t *row0 = spikes + (size_t)ch * R * T;
        for (int k = 0; k < K; ++k) {
            const float th = (float)thr[k], lo = (float)lower[k];
            int on = 0;
            for (int b = 0; b < nbins; ++b) {
                const float v = norm[(size_t)ch * nbins + b];
                if (!on && v > th) on = 1;
                else if (on && v < lo) on = 0;
                row0[b * K + k] = (uint8_t)on;
            }
        }
        for (int r = 1; r < R; ++r) memcpy(row0 + (size_t)r * T, row0, (size_t)T);
    }
}

void oracle_hysteresis_encode_f32(const float *norm, int C, int nbins, const double *thr_desc,
                                  const double *lower, int K, int R, uint8_t *spikes)
{
    hysteresis_encode_f32(norm, C, nbins, thr_desc, lower, K, R, spikes);
}

typedef struct {
    const float *pcm; int L, n_fft, hop, C, ncols, nbins;
    const double *win, *tw, *tw2; const float *mel_w; const int32_t *mel_lo, *mel_n, *mel_off;
    const int32_t *zi0; const double *zf; const double *thr, *lower; int K, R;
    uint8_t *spikes; float *spec_norm_out;
} mel_ctx;

static void *mel_mk(void *vc)
{
    mel_ctx *c = (mel_ctx *)vc;
    return malloc(sizeof(double) * 2048 + sizeof(float) * (1025 + 8 + (size_t)c->C * c->ncols + (size_t)c->C * c->nbins));
}

static void mel_run_one(void *vc, int b, void *scratch)
{
    mel_ctx *c = (mel_ctx *)vc;
    double *re = (double *)scratch, *im = re + 1024;
    float *S = (float *)(im + 1024), *M = S + 1032, *norm = M + (size_t)c->C * c->ncols;
    mel_one(c->pcm + (size_t)b * c->L, c->L, c->n_fft, c->hop, c->C, c->ncols, c->nbins, c->win, c->tw, c->tw2,
            c->mel_w, c->mel_lo, c->mel_n, c->mel_off, c->zi0, c->zf, re, im, S, M, norm);
    hysteresis_encode_f32(norm, c->C, c->nbins, c->thr, c->lower, c->K, c->R,
                          c->spikes + (size_t)b * c->C * c->R * c->nbins * c->K);
    if (c->spec_norm_out)
        memcpy(c->spec_norm_out + (size_t)b * c->C * c->nbins, norm, sizeof(float) * c->C * c->nbins);
}

int oracle_mel_encode(const float *pcm, int B, int L, int n_fft, int hop, int C, int nbins,
                      const double *win, const double *tw, const double *tw2,
                      const float *mel_w, const int32_t *mel_lo, const int32_t *mel_n, const int32_t *mel_off,
                      const int32_t *zi0, const double *zf, const double *thr_desc, const double *lower, int K, int R,
                      uint8_t *spikes, float *spec_norm_out, int nthreads)
{
    if (n_fft != 2048) return -1;
    mel_ctx c = {pcm, L, n_fft, hop, C, 1 + L / hop, nbins, win, tw, tw2, mel_w, mel_lo, mel_n, mel_off, zi0, zf,
                 thr_desc, lower, K, R, spikes, spec_norm_out};
    pf_job j = {mel_run_one, &c, B, 0, mel_mk, free};
    parallel_for(&j, nthreads);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Stages 2+3: frozen reservoir spec; call sites
 * extract_lsm_features.py:79-83.  Event-driven over the previous step's spike list, integer
 * recurrent sums (exact), fp64 membrane.
 * wt_rowptr/wt_col/wt_q: CSR over the PRESYNAPTIC neuron (outgoing edges).                     */
typedef struct {
    int N, C, T, refractory, n_out;
    double theta;
    int w_shift;
} oracle_res_dims;

/* wt_val != NULL: strict reservoir: fp64 weights, and the recurrent current of a neuron is their sum taken
 * one by one in ascending presynaptic index - the spike list is ascending, so scattering it row by row adds in that order. */
static void simulate_one(const oracle_res_dims *d, const uint8_t *x,
                         const int32_t *wt_rowptr, const int32_t *wt_col, const int32_t *wt_q, const double *wt_val,
                         const int32_t *in_rowptr, const int32_t *in_col, const double *in_val,
                         const double *leak, const int32_t *out_idx, uint32_t feature_mask,
                         int nan_to_num, double *features, uint8_t *raster,
                         double *V, int32_t *ref, int32_t *spk, int64_t *acc, int32_t *st)
{
    const int N = d->N, T = d->T;
    double *accd = (double *)acc;                        /* strict reservoirs reuse the accumulator array as doubles */
    const double scale = ldexp(1.0, -d->w_shift);
    int nspk = 0;
    /* per-neuron streaming statistics: count, sum t, first, last, sum isi^2, bursts */
    int32_t *cnt = st, *sumt = st + N, *first = st + 2 * N, *last = st + 3 * N, *burst = st + 4 * N;
    int64_t *s2 = (int64_t *)(st + 5 * N + (N & 1));
    for (int i = 0; i < N; ++i) {
        V[i] = 0.0; ref[i] = 0; cnt[i] = 0; sumt[i] = 0; first[i] = -1; last[i] = -1; burst[i] = 0; s2[i] = 0;
    }
    for (int t = 0; t < T; ++t) {
        if (wt_val) {
            for (int i = 0; i < N; ++i) accd[i] = 0.0;
            for (int q = 0; q < nspk; ++q) {
                int j = spk[q];
                for (int p = wt_rowptr[j]; p < wt_rowptr[j + 1]; ++p) accd[wt_col[p]] = accd[wt_col[p]] + wt_val[p];
            }
        } else {
            for (int i = 0; i < N; ++i) acc[i] = 0;
            for (int q = 0; q < nspk; ++q) {
                int j = spk[q];
                for (int p = wt_rowptr[j]; p < wt_rowptr[j + 1]; ++p) acc[wt_col[p]] += wt_q[p];
            }
        }
        nspk = 0;
        for (int i = 0; i < N; ++i) {
            double i_in = 0.0;
            for (int p = in_rowptr[i]; p < in_rowptr[i + 1]; ++p)
                i_in = i_in + in_val[p] * (x[(size_t)in_col[p] * T + t] ? 1.0 : 0.0);  /* level signal: non-zero = on */
            double cur = wt_val ? i_in + accd[i] : i_in + (double)acc[i] * scale;
            int fire = 0;
            if (ref[i] == 0) {
                double v = (V[i] - leak[i] * V[i]) + cur;
                if (v >= d->theta) { fire = 1; v = 0.0; ref[i] = d->refractory; }
                V[i] = v;
            } else {
                V[i] = 0.0;
                ref[i] -= 1;
            }
            if (raster) raster[(size_t)t * N + i] = (uint8_t)fire;
            if (fire) {
                spk[nspk++] = i;
                if (cnt[i] > 0) {
                    int isi = t - last[i];
                    s2[i] += (int64_t)isi * isi;
                    if (isi <= d->refractory + 1) burst[i] += 1;
                } else first[i] = t;
                cnt[i] += 1; sumt[i] += t; last[i] = t;
            }
        }
    }
    if (!features) return;
    /* key-major layout: all output neurons of key 0, then key 1, ... (extract_lsm_features.py:85-87) */
    int slot = 0;
    for (int key = 0; key < 8; ++key) {
        if (!(feature_mask & (1u << key))) continue;
        for (int o = 0; o < d->n_out; ++o) {
            int i = out_idx[o];
            double c = (double)cnt[i], v = NAN;
            switch (key) {
            case 0: v = c; break;
            case 1: { double p = c / (double)T; v = p * (1.0 - p); } break;
            case 2: if (cnt[i] >= 1) v = (double)sumt[i] / c; break;
            case 3: if (cnt[i] >= 1) v = (double)first[i]; break;
            case 4: if (cnt[i] >= 1) v = (double)last[i]; break;
            case 5: if (cnt[i] >= 2) v = (double)(last[i] - first[i]) / (double)(cnt[i] - 1); break;
            case 6: if (cnt[i] >= 2) {
                        int64_t n = cnt[i] - 1, s1 = last[i] - first[i];
                        v = (double)(n * s2[i] - s1 * s1) / (double)(n * n);
                    } break;
            case 7: v = (double)burst[i]; break;
            }
            if (nan_to_num && v != v) v = 0.0;
            features[(size_t)slot * d->n_out + o] = v;
        }
        ++slot;
    }
}

typedef struct {
    oracle_res_dims d;
    const int32_t *wt_rowptr, *wt_col, *wt_q; const double *wt_val; const int32_t *in_rowptr, *in_col; const double *in_val, *leak;
    const int32_t *out_idx; const uint8_t *spikes; uint32_t feature_mask; int nan_to_num, nkeys;
    double *features; uint8_t *raster;
} rs_ctx;

static void *rs_mk(void *vc)
{
    rs_ctx *c = (rs_ctx *)vc;
    size_t N = (size_t)c->d.N;
    /* V[N] f64 | acc[N] i64 | ref[N] i32 | spk[N] i32 | st[8N+4] i32 */
    return malloc(8 * N + 8 * N + 4 * N + 4 * N + 4 * (8 * N + 4));
}

static void rs_one(void *vc, int b, void *scratch)
{
    rs_ctx *c = (rs_ctx *)vc;
    size_t N = (size_t)c->d.N;
    double *V = (double *)scratch;
    int64_t *acc = (int64_t *)(V + N);
    int32_t *ref = (int32_t *)(acc + N), *spk = ref + N, *st = spk + N + (N & 1);
    simulate_one(&c->d, c->spikes + (size_t)b * c->d.C * c->d.T, c->wt_rowptr, c->wt_col, c->wt_q, c->wt_val,
                 c->in_rowptr, c->in_col, c->in_val, c->leak, c->out_idx, c->feature_mask, c->nan_to_num,
                 c->features ? c->features + (size_t)b * c->nkeys * c->d.n_out : NULL,
                 c->raster ? c->raster + (size_t)b * c->d.T * N : NULL, V, ref, spk, acc, st);
}

int oracle_reservoir_run(int N, int C, int T, double theta, int refractory, int w_shift,
                         const int32_t *wt_rowptr, const int32_t *wt_col, const int32_t *wt_q, const double *wt_val_or_null,
                         const int32_t *in_rowptr, const int32_t *in_col, const double *in_val,
                         const double *leak, const int32_t *out_idx, int n_out,
                         const uint8_t *spikes, int B, uint32_t feature_mask, int nan_to_num,
                         double *features, uint8_t *raster, int nthreads)
{
    rs_ctx c = {{N, C, T, refractory, n_out, theta, w_shift}, wt_rowptr, wt_col, wt_q, wt_val_or_null, in_rowptr, in_col,
                in_val, leak, out_idx, spikes, feature_mask, nan_to_num, 0, features, raster};
    for (int k = 0; k < 8; ++k) c.nkeys += (feature_mask >> k) & 1;
    pf_job j = {rs_one, &c, B, 0, rs_mk, free};
    parallel_for(&j, nthreads);
    return 0;
}

int oracle_num_threads(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }
