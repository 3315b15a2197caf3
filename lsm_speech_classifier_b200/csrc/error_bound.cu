// Worst-case distance between the two arrangements of the gammatone cascade (host code, run once per front end).
//
// The speculative filter (gammatone_core.cuh gt_filter_fast and the lane = utterance kernels) and the reference-order filter
// (gt_filter_exact = scipy.signal.lfilter x 4 as gammatone==1.0.3 calls it, /root/reference/create_dataset.py:51-58) evaluate
// the same linear filter in floating point.  Each differs from the infinitely precise result by its rounding errors, and a
// rounding error committed in section k at time n reaches the output through the rest of the cascade.  With
//   u        = 2^-53 (every operation is a correctly rounded fp64 +, *, or fma),
//   c_k      = A1k/A0, a1 = B1/B0, a2 = B2/B0, sections H_k(z) = (1 + c_k z^-1) / (1 + a1 z^-1 + a2 z^-2), P(z) = 1 / (1 + a1 z^-1 + a2 z^-2),
//   M_k      = || h_1 * ... * h_k ||_1  (so that max |output of section k| <= M_k X for max |x| = X; M_0 = 1),
//   L_k      = || p * h_(k+1) * ... * h_4 ||_1  (an error injected at the output node of section k is multiplied by at most L_k),
// the error of the normalised cascade output is, to first order in u, at most  u X sum_k L_k C_k  where C_k adds up the
// magnitudes of the values that are rounded in section k:
//   speculative (direct form, three FMAs: t1 = c_k v' + v, t2 = -a2 y'' + t1, y = -a1 y' + t2; the stored c_k is itself a
//                rounded quotient, a relative perturbation u of the term c_k v'):
//       C_k = (2 (1 + |c_k|) + |c_k|) M_(k-1) + (1 + |a1| + 2 |a2|) M_k
//   exact       (direct form II transposed, seven roundings: b0 v; z0 + .; v b1; z1 + .; y a1; . - .; -(y a2)), in the
//                same normalised units:
//       C'_k = (1 + 3 |c_k|) M_(k-1) + (1 + 2 |a1| + 3 |a2|) M_k
// All norms are taken over the n_samples taps an utterance can excite (zero initial state).  The window amplitude
// sqrt(mean(y^2)) is 1-Lipschitz in the sup norm of y, so
//   | amplitude_speculative - amplitude_exact |  <=  kappa X + (relative terms of the sums and square roots) amplitude,
//   kappa = 2 * G u sum_k L_k (C_k + C'_k),   G = A0^4 / gain,
// the factor 2 covering the second-order terms and the fact that M_k bounds the exact, not the computed, signals.
// tests/test_error_bound.py recomputes kappa with scipy; tests/test_gpu_parity.py checks on the GPU that no cell of ten
// thousand adversarial clips comes anywhere near it (lsm_frontend_audit).
#include <math.h>

#include <vector>

#include "lsm_common.cuh"

namespace {

// y = x filtered by (1 + c z^-1) / (1 + a1 z^-1 + a2 z^-2), zero state; c = 0 and unit numerator give P
void section(const std::vector<double> &x, double c, double a1, double a2, std::vector<double> &y)
{
    const size_t n = x.size();
    y.resize(n);
    double xp = 0.0, y1 = 0.0, y2 = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double v = x[i] + c * xp - a1 * y1 - a2 * y2;
        xp = x[i];
        y2 = y1;
        y1 = v;
        y[i] = v;
    }
}

double l1(const std::vector<double> &x)
{
    double s = 0.0;
    for (double v : x) s += fabs(v);
    return s;
}

}  // namespace

extern "C" int lsm_gammatone_error_bound(const double *h_table, int32_t channels, int32_t n_samples, double *h_kappa)
{
    if (!h_table || !h_kappa || channels <= 0 || n_samples <= 0) return LSM_ERR_INVALID;
    const double u = ldexp(1.0, -53);
    std::vector<double> delta((size_t)n_samples, 0.0), tmp, tail[5], pre[5];
    delta[0] = 1.0;
    for (int ch = 0; ch < channels; ++ch) {
        const double *r = h_table + 10 * (size_t)ch;
        const double A0 = r[0], B0 = r[6], gain = r[9];
        if (!(A0 != 0.0) || !(B0 != 0.0) || !(gain != 0.0)) return LSM_ERR_INVALID;
        const double c[4] = {r[1] / A0, r[2] / A0, r[3] / A0, r[4] / A0};
        const double a1 = r[7] / B0, a2 = r[8] / B0;
        // pre[k] = h_1 * ... * h_k (pre[0] = delta); tail[k] = h_(k+1) * ... * h_4 (tail[4] = delta)
        pre[0] = delta;
        for (int k = 0; k < 4; ++k) section(pre[k], c[k], a1, a2, pre[k + 1]);
        tail[4] = delta;
        for (int k = 3; k >= 0; --k) section(tail[k + 1], c[k], a1, a2, tail[k]);
        double sum = 0.0;
        for (int k = 1; k <= 4; ++k) {
            section(tail[k], 0.0, a1, a2, tmp);                 // p * h_(k+1..4)
            const double Lk = l1(tmp), Mp = l1(pre[k - 1]), Mk = l1(pre[k]), ck = fabs(c[k - 1]);
            const double Cs = (2.0 * (1.0 + ck) + ck) * Mp + (1.0 + fabs(a1) + 2.0 * fabs(a2)) * Mk;
            const double Ce = (1.0 + 3.0 * ck) * Mp + (1.0 + 2.0 * fabs(a1) + 3.0 * fabs(a2)) * Mk;
            sum += Lk * (Cs + Ce);
        }
        const double G = fabs((A0 * A0) * (A0 * A0) / gain);
        const double kappa = 2.0 * G * u * sum;
        if (!(kappa == kappa) || kappa > 1e300) return LSM_ERR_INVALID;
        h_kappa[ch] = kappa;
    }
    return LSM_OK;
}
