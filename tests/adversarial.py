"""Adversarial utterances for the speculative filter's error bound (test infrastructure).

Each family targets one way the two arrangements of the gammatone cascade could drift apart or the bound could be
too small: strong out-of-band energy next to an in-band component near the -80 dB floor, full-scale clipping, PCM16
extremes, 80 dB dynamic range, single-sample impulses, resonance at a channel's centre frequency, and the
sign patterns that attain the l1 norms the bound is built from (x[n] = sign(h[N - n]) drives |y[N]| to ||h||_1)."""
from __future__ import annotations

import numpy as np

L = 16000
FS = 16000.0
N_FAMILIES = 12


def _tone(f, amp=1.0, phase=0.0):
    return amp * np.sin(2 * np.pi * f * np.arange(L) / FS + phase)


def clip(i: int, seed: int = 0, coefs: np.ndarray | None = None) -> np.ndarray:
    """Clip number i -> float32[16000].  `coefs` (the gammatone table) is needed for the l1-attaining family."""
    rng = np.random.default_rng(1_000_003 * seed + i)
    fam = i % N_FAMILIES
    t = np.arange(L) / FS
    if fam == 0:      # out-of-band tone + in-band component 60..79 dB below it
        x = _tone(rng.uniform(5000, 7800), 0.9) + _tone(rng.uniform(50, 120), 0.9 * 10 ** (-rng.uniform(60, 79) / 20))
    elif fam == 1:    # full-scale square wave (clipped sine)
        x = np.sign(_tone(rng.uniform(40, 7900), 1.0, rng.uniform(0, 6.28)))
    elif fam == 2:    # PCM16 extremes, alternating in blocks of random length
        k = int(rng.integers(1, 200))
        x = np.where((np.arange(L) // k) % 2 == 0, 32767.0, -32768.0) / 32768.0
    elif fam == 3:    # chirp over the whole band with an 80 dB exponential amplitude ramp
        f0, f1 = (50.0, 7800.0) if rng.random() < 0.5 else (7800.0, 50.0)
        ph = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t)
        x = np.sin(ph) * 10 ** (-4 * (t if rng.random() < 0.5 else 1 - t))
    elif fam == 4:    # single-sample impulses
        x = np.zeros(L)
        for _ in range(int(rng.integers(1, 4))):
            x[int(rng.integers(0, L))] = rng.choice([-1.0, 1.0])
    elif fam == 5:    # white noise, full scale or tiny
        x = rng.uniform(-1, 1, L) * (1.0 if rng.random() < 0.5 else 10 ** (-rng.uniform(3, 6)))
    elif fam == 6:    # DC plus a tiny tone
        x = rng.choice([-1.0, 1.0]) * np.ones(L) + _tone(rng.uniform(100, 4000), 10 ** (-rng.uniform(2, 5)))
    elif fam == 7:    # resonance: a tone exactly at a channel's centre frequency (ERB scale, 128 channels from 50 Hz)
        c = 9.26449 * 24.7
        ch = int(rng.integers(1, 129))
        cf = -c + np.exp((ch / 128) * (np.log(50 + c) - np.log(8000 + c))) * (8000 + c)
        x = _tone(cf, 1.0, rng.uniform(0, 6.28))
    elif fam == 8:    # loud burst between silences: long decaying tails (denormal range in the high channels)
        x = np.zeros(L)
        a, b = sorted(rng.integers(0, L, 2))
        x[a:b + 1] = rng.uniform(-1, 1, b + 1 - a)
    elif fam == 9:    # two-level: loud first half, then the same signal 70 dB down
        s = _tone(rng.uniform(100, 3000), 1.0) + 0.3 * rng.uniform(-1, 1, L)
        x = s * np.where(t < 0.5, 1.0, 10 ** (-70 / 20))
        x /= np.max(np.abs(x))
    elif fam == 10:   # l1-attaining sign pattern of a low channel's impulse response (worst case of the signal bound)
        from scipy.signal import lfilter
        if coefs is None:
            x = np.sign(_tone(50.0))
        else:
            ch = int(rng.integers(0, 16))
            r = coefs[ch]
            h = np.zeros(L); h[0] = 1.0
            for k in range(4):
                h = lfilter([1.0, r[1 + k] / r[0]], [1.0, r[7] / r[6], r[8] / r[6]], h)
            n_end = int(rng.integers(4000, L))
            x = np.zeros(L)
            x[:n_end + 1] = np.sign(h[n_end::-1])
            x[x == 0] = 1.0
    else:             # speech-like clip of the bench generator, rescaled to full scale
        from lsm_speech_classifier_b200 import synth
        x = synth.synth_utterance(int(rng.integers(0, 12)), int(rng.integers(0, 100000))).astype(np.float64) * 2.0
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def clips(start: int, n: int, seed: int = 0, coefs: np.ndarray | None = None) -> np.ndarray:
    return np.stack([clip(start + i, seed, coefs) for i in range(n)])
