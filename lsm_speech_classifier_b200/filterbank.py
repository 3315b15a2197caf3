"""Host-side filterbank design tables (tiny, once per run; the per-sample work is CUDA).

The reference gets these numbers from two un-vendored packages:

* ``gammatone==1.0.3`` (requirements.txt:85), called at
  /root/reference/create_dataset.py:51-58 as
  ``gtgram.gtgram(wave, fs=16000, window_time=0.025, hop_time=0.01, channels=n, f_min=50)``.
  The design below restates that package's published algorithm (Slaney's Auditory
  Toolbox ``ERBSpace`` / ``MakeERBFilters``): SURVEY.md Appendix B.1.
* ``librosa==0.11.0`` (requirements.txt:31), called at create_dataset.py:45-48 as
  ``melspectrogram(y, sr=16000, n_mels=n, hop_length=160)`` (n_fft 2048, periodic hann,
  centre zero padding, power 2, Slaney mel scale with slaney norm): Appendix B.2.

Only the *design* (O(channels) transcendental work) is done here with numpy, exactly as
the reference does it on the host.  The tables are handed to the CUDA library as data,
so the CPU oracle and the GPU always filter with bit-identical coefficients.
"""
from __future__ import annotations

import numpy as np

EAR_Q = 9.26449  # Glasberg and Moore
MIN_BW = 24.7


def erb_space(low_freq: float, high_freq: float, num: int) -> np.ndarray:
    """ERB-spaced centre frequencies, DESCENDING (index 0 is the highest)."""
    frac = np.arange(1, num + 1) / num
    c = EAR_Q * MIN_BW
    return -c + np.exp(frac * (-np.log(high_freq + c) + np.log(low_freq + c))) * (high_freq + c)


def centre_freqs(fs: float, num_freqs: int, cutoff: float) -> np.ndarray:
    return erb_space(cutoff, fs / 2, num_freqs)


def make_erb_filters(fs: float, cfs: np.ndarray, width: float = 1.0) -> np.ndarray:
    """Gammatone coefficients, one row per centre frequency:
    [A0, A11, A12, A13, A14, A2, B0, B1, B2, gain] (float64[len(cfs), 10])."""
    cfs = np.asarray(cfs, dtype=np.float64)
    T = 1 / fs
    erb = width * ((cfs / EAR_Q) + MIN_BW)
    B = 1.019 * 2 * np.pi * erb
    arg = 2 * cfs * np.pi * T
    vec = np.exp(2j * arg)

    A0 = T
    A2 = 0
    B0 = 1
    B1 = -2 * np.cos(arg) / np.exp(B * T)
    B2 = np.exp(-2 * B * T)

    rt_pos = np.sqrt(3 + 2 ** 1.5)
    rt_neg = np.sqrt(3 - 2 ** 1.5)
    common = -T * np.exp(-(B * T))

    k11 = np.cos(arg) + rt_pos * np.sin(arg)
    k12 = np.cos(arg) - rt_pos * np.sin(arg)
    k13 = np.cos(arg) + rt_neg * np.sin(arg)
    k14 = np.cos(arg) - rt_neg * np.sin(arg)

    A11 = common * k11
    A12 = common * k12
    A13 = common * k13
    A14 = common * k14

    gain_arg = np.exp(1j * arg - B * T)
    gain = np.abs(
        (vec - gain_arg * k11)
        * (vec - gain_arg * k12)
        * (vec - gain_arg * k13)
        * (vec - gain_arg * k14)
        * (T * np.exp(B * T) / (-1 / np.exp(B * T) + 1 + vec * (1 - np.exp(B * T)))) ** 4
    )
    ones = np.ones_like(cfs)
    return np.column_stack([A0 * ones, A11, A12, A13, A14, A2 * ones, B0 * ones, B1, B2, gain])


def gammatone_coefs(fs: int, channels: int, f_min: float) -> np.ndarray:
    """The table ``gtgram_xe`` filters with: rows flipped so row 0 is the LOWEST
    centre frequency (f_min).  float64[channels, 10], C-contiguous."""
    return np.ascontiguousarray(np.flipud(make_erb_filters(fs, centre_freqs(fs, channels, f_min))))


def _round_half_away_from_zero(x: float) -> float:
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def gtgram_strides(fs: int, window_time: float, hop_time: float, n_samples: int):
    """(nwin, hop, ncols): (16000, 0.025, 0.01, 16000) -> (400, 160, 98)."""
    nwin = int(_round_half_away_from_zero(window_time * fs))
    hop = int(_round_half_away_from_zero(hop_time * fs))
    ncols = 1 + int(np.floor((n_samples - nwin) / hop))
    return nwin, hop, ncols


def zoom_table(n_in: int, n_out: int):
    """scipy.ndimage.zoom(order=1) along one axis, as data: for output bin j the
    source coordinate is j*(n_in-1)/(n_out-1); out = v[i0]*(1-f) + v[i0+1]*f, the second
    term dropped when i0+1 == n_in.  Pinned bit-exact against scipy here
    (tests/test_oracle_frontend.py).  Returns (i0 int32[n_out], f float64[n_out])."""
    zz = (n_in - 1) / (n_out - 1)
    cc = np.arange(n_out, dtype=np.float64) * zz
    i0 = np.floor(cc)
    f = cc - i0
    return i0.astype(np.int32), f


# ---------------------------------------------------------------- mel (librosa restated)

def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: float | None = None) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney') -> float32[n_mels, 1 + n_fft//2]."""
    if fmax is None:
        fmax = sr / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True) as float64."""
    return 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)


def fft_tables(n_fft: int = 2048):
    """Twiddle tables for the 1024-point complex FFT of a packed real 2048-point frame:
    tw[q] = exp(-2*pi*i*q/(n_fft/2)), q < n_fft/4;  tw2[k] = exp(-2*pi*i*k/n_fft), k <= n_fft/2.
    float64[..., 2] (re, im), built once on the host and handed to both the CUDA kernel and the oracle."""
    half = n_fft // 2
    q = np.arange(half // 2, dtype=np.float64)
    k = np.arange(half + 1, dtype=np.float64)
    tw = np.stack([np.cos(2 * np.pi * q / half), -np.sin(2 * np.pi * q / half)], axis=1)
    tw2 = np.stack([np.cos(2 * np.pi * k / n_fft), -np.sin(2 * np.pi * k / n_fft)], axis=1)
    return np.ascontiguousarray(tw), np.ascontiguousarray(tw2)


def pack_mel_basis(basis: np.ndarray):
    """Non-zero run of every mel triangle: (weights float32[total], lo int32[C], n int32[C], off int32[C]).
    Zeros add exactly nothing in the projection, so skipping them is exact."""
    basis = np.asarray(basis, dtype=np.float32)
    lo, n, off, w = [], [], [], []
    total = 0
    for row in basis:
        nz = np.nonzero(row)[0]
        if len(nz) == 0:
            lo.append(0); n.append(0); off.append(total)
            continue
        a, b = int(nz[0]), int(nz[-1]) + 1
        lo.append(a); n.append(b - a); off.append(total)
        w.append(row[a:b])
        total += b - a
    wcat = np.concatenate(w) if w else np.zeros(1, np.float32)
    return (np.ascontiguousarray(wcat, dtype=np.float32), np.array(lo, np.int32), np.array(n, np.int32), np.array(off, np.int32))
