"""CPU: the derived bound behind the speculative gammatone filter (csrc/error_bound.cu).

1. lsm_gammatone_error_bound (host code of the C library) against an independent scipy computation of the same formula.
2. The bound itself, against both arrangements run on the CPU with identical IEEE arithmetic (the oracle restates the
   product's 13-FMA arrangement next to the reference-order cascade): over adversarial clips every window amplitude of one
   lies within kappa * max|x| (+ the relative terms) of the other."""
import ctypes as C

import numpy as np
import pytest

import adversarial


def _table():
    from lsm_speech_classifier_b200 import filterbank as fb
    return fb.gammatone_coefs(16000, 128, 50)


def _kappa_lib(table, n_samples=16000):
    from lsm_speech_classifier_b200 import _lib
    lib = _lib.load()
    out = np.zeros(len(table), np.float64)
    assert lib.lsm_gammatone_error_bound(_lib._np_ptr(table), len(table), n_samples, _lib._np_ptr(out)) == 0
    return out


def _kappa_scipy(table, n_samples=16000):
    from scipy.signal import lfilter
    u = 2.0 ** -53
    imp = np.zeros(n_samples); imp[0] = 1.0
    out = []
    for r in table:
        A0, B0, gain = r[0], r[6], r[9]
        c = [r[1 + k] / A0 for k in range(4)]
        den = [1.0, r[7] / B0, r[8] / B0]
        a1, a2 = abs(den[1]), abs(den[2])
        pre = [imp]
        for k in range(4):
            pre.append(lfilter([1.0, c[k]], den, pre[-1]))
        tail = [None] * 5
        tail[4] = imp
        for k in range(3, -1, -1):
            tail[k] = lfilter([1.0, c[k]], den, tail[k + 1])
        total = 0.0
        for k in range(1, 5):
            Lk = np.abs(lfilter([1.0], den, tail[k])).sum()
            Mp, Mk, ck = np.abs(pre[k - 1]).sum(), np.abs(pre[k]).sum(), abs(c[k - 1])
            total += Lk * (((2 * (1 + ck) + ck) * Mp + (1 + a1 + 2 * a2) * Mk) + ((1 + 3 * ck) * Mp + (1 + 2 * a1 + 3 * a2) * Mk))
        out.append(2.0 * abs(A0 ** 4 / gain) * u * total)
    return np.array(out)


def test_bound_matches_independent_computation():
    table = _table()
    got = _kappa_lib(table)
    want = _kappa_scipy(table)
    np.testing.assert_allclose(got, want, rtol=1e-9)
    # magnitudes: worst for the narrow low channels (row 0 = 50 Hz), tiny for the wide high ones
    assert 1e-11 < got[0] < 1e-10 and got[0] == got.max() and got.min() < 1e-12
    # rejects a degenerate table
    from lsm_speech_classifier_b200 import _lib
    bad = table.copy(); bad[3, 9] = 0.0
    out = np.zeros(128)
    assert _lib.load().lsm_gammatone_error_bound(_lib._np_ptr(bad), 128, 16000, _lib._np_ptr(out)) != 0


@pytest.mark.parametrize("start", [0, 12, 24, 36])
def test_two_arrangements_stay_within_the_bound_on_adversarial_clips(start):
    from oracle import coracle
    table = _table()
    kappa = _kappa_lib(table)
    worst = 0.0
    for i in range(start, start + 12):
        x = adversarial.clip(i, seed=7, coefs=table)
        exact, fast = coracle.two_arrangements(x, table, 400, 160, 98)
        X = float(np.max(np.abs(x)))
        bound = kappa[:, None] * X + 6e-14 * np.maximum(exact, fast)
        d = np.abs(exact - fast)
        assert np.all(d <= bound), (i, float((d / bound).max()))
        worst = max(worst, float((d / np.maximum(bound, 1e-300)).max()))
    # the worst-case bound is pessimistic by orders of magnitude on any actual signal - but it is a bound
    assert worst < 0.5
