"""TEST INFRASTRUCTURE — CPU restatement of the reference's audio -> spikes -> LSM -> features
path in numpy/scipy.  Never imported by the product package; only tests/, bench.py's
cpu_baseline leg and __graft_entry__.smoke() may use anything under oracle/.

PARITY STATUS (also in DESIGN.md):
* Encoder, redundancy, w_critico: pinned.  The reference's own functions
  (/root/reference/create_dataset.py:81-104, extract_lsm_features.py:33-60) run verbatim
  in the build container (third-party imports stubbed) and were used to mint
  tests/golden/*.npz (tests/golden/make_golden.py).
* lfilter / zoom / window mean: pinned against the installed scipy/numpy, which are the
  reference's dependencies (newer versions than its requirements.txt).
* gammatone (gammatone==1.0.3), mel (librosa==0.11.0), reservoir (snn_reservoir_py==2.0.0,
  module snnpy): PARITY UNPINNED.  Those packages are not vendored in /root/reference and are
  not installable here; the reference has no tests or golden vectors.  The code below
  restates their published algorithms and, for snnpy, a frozen
  spec (DESIGN.md "Reservoir spec") consistent with every call site in the reference.

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import zoom
from scipy.signal import lfilter

SAMPLE_RATE = 16000          # create_dataset.py:10
TIME_BINS = 100              # create_dataset.py:12
SPIKE_THRESHOLDS = [0.70, 0.80, 0.90, 0.95]  # create_dataset.py:13
HYSTERESIS_GAP = 0.1         # create_dataset.py:14


# --------------------------------------------------------------------------- gammatone
# gammatone==1.0.3: filters.erb_space / make_erb_filters / erb_filterbank, gtgram.gtgram
# (call site create_dataset.py:51-58).  Written channel-by-channel with python scalars
# on purpose: an independent derivation from the product's vectorised design table.

def gammatone_design(fs, channels, f_min):
    """-> float64[channels,10] rows [A0,A11,A12,A13,A14,A2,B0,B1,B2,gain], row 0 = f_min."""
    ear_q, min_bw = 9.26449, 24.7
    c = ear_q * min_bw
    hi = fs / 2
    rows = []
    T = 1 / fs
    rt_pos = np.sqrt(3 + 2 ** 1.5)
    rt_neg = np.sqrt(3 - 2 ** 1.5)
    for i in range(1, channels + 1):
        cf = -c + np.exp((i / channels) * (-np.log(hi + c) + np.log(f_min + c))) * (hi + c)
        erb = 1.0 * ((cf / ear_q) + min_bw)
        B = 1.019 * 2 * np.pi * erb
        arg = 2 * cf * np.pi * T
        vec = np.exp(2j * arg)
        B1 = -2 * np.cos(arg) / np.exp(B * T)
        B2 = np.exp(-2 * B * T)
        common = -T * np.exp(-(B * T))
        ks = [np.cos(arg) + rt_pos * np.sin(arg), np.cos(arg) - rt_pos * np.sin(arg),
              np.cos(arg) + rt_neg * np.sin(arg), np.cos(arg) - rt_neg * np.sin(arg)]
        g = np.exp(1j * arg - B * T)
        gain = np.abs((vec - g * ks[0]) * (vec - g * ks[1]) * (vec - g * ks[2]) * (vec - g * ks[3])
                      * (T * np.exp(B * T) / (-1 / np.exp(B * T) + 1 + vec * (1 - np.exp(B * T)))) ** 4)
        rows.append([T, common * ks[0], common * ks[1], common * ks[2], common * ks[3], 0.0,
                     1.0, B1, B2, gain])
    return np.array(rows[::-1], dtype=np.float64)  # flipud: gtgram_xe


def gtgram(wave, fs, window_time, hop_time, channels, f_min, coefs=None):
    """gammatone.gtgram.gtgram restated -> float64[channels, ncols]."""
    if coefs is None:
        coefs = gammatone_design(fs, channels, f_min)
    wave = np.asarray(wave)
    xf = np.zeros((coefs.shape[0], wave.shape[0]))
    for ch in range(coefs.shape[0]):
        A0, A11, A12, A13, A14, A2, B0, B1, B2, gain = coefs[ch]
        den = [B0, B1, B2]
        y1 = lfilter([A0, A11, A2], den, wave)
        y2 = lfilter([A0, A12, A2], den, y1)
        y3 = lfilter([A0, A13, A2], den, y2)
        y4 = lfilter([A0, A14, A2], den, y3)
        xf[ch, :] = y4 / gain
    xe = np.power(xf, 2)
    nwin = int(np.sign(window_time * fs) * np.floor(np.abs(window_time * fs) + 0.5))
    hop = int(np.sign(hop_time * fs) * np.floor(np.abs(hop_time * fs) + 0.5))
    ncols = 1 + int(np.floor((xe.shape[1] - nwin) / hop))
    y = np.zeros((channels, ncols))
    for cnum in range(ncols):
        # fancy indexing gives an F-ordered segment => mean(1) is a plain left-to-right sum
        segment = xe[:, cnum * hop + np.arange(nwin)]
        y[:, cnum] = np.sqrt(segment.mean(1))
    return y


# --------------------------------------------------------------------------- mel
# librosa==0.11.0 melspectrogram + power_to_db (call site create_dataset.py:43-48).

def _slaney_mel_edges(n_mels, fmin, fmax):
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def h2m(f):
        return min_log_mel + np.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    def m2h(m):
        return min_log_hz * np.exp(logstep * (m - min_log_mel)) if m >= min_log_mel else f_sp * m

    return np.array([m2h(m) for m in np.linspace(h2m(fmin), h2m(fmax), n_mels + 2)])


def mel_filters(sr, n_fft, n_mels):
    edges = _slaney_mel_edges(n_mels, 0.0, sr / 2)
    freqs = np.arange(1 + n_fft // 2) * (sr / n_fft)
    w = np.zeros((n_mels, len(freqs)), dtype=np.float32)
    for i in range(n_mels):
        lower = (freqs - edges[i]) / (edges[i + 1] - edges[i])
        upper = (edges[i + 2] - freqs) / (edges[i + 2] - edges[i + 1])
        w[i] = np.maximum(0, np.minimum(lower, upper))
        w[i] *= 2.0 / (edges[i + 2] - edges[i])
    return w


def mel_power_db(audio, n_mels, sr=SAMPLE_RATE, n_fft=2048, hop=160, basis=None):
    """melspectrogram(power=2) then power_to_db(ref=max, amin=1e-10, top_db=80) -> float32[n_mels,101].
    STFT: centre zero padding, periodic hann (float64) times float32 frames => float64 rFFT,
    stored as complex64; |.|^2 in float32; mel projection in float32, ascending-bin order
    (librosa's BLAS order is unknowable, ours is the stated one)."""
    from scipy.fft import rfft
    y = np.asarray(audio, dtype=np.float32)
    ypad = np.concatenate([np.zeros(n_fft // 2, np.float32), y, np.zeros(n_fft // 2, np.float32)])
    n_frames = 1 + (len(ypad) - n_fft) // hop
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)
    frames = np.stack([ypad[t * hop: t * hop + n_fft] for t in range(n_frames)], axis=1)  # [n_fft, T]
    stft = rfft(win[:, None] * frames, axis=0).astype(np.complex64)
    re = stft.real.astype(np.float64)
    im = stft.imag.astype(np.float64)
    mag = np.sqrt(re * re + im * im).astype(np.float32)      # hypotf computed in double
    S = (mag * mag).astype(np.float32)                       # ** 2.0 in float32
    if basis is None:
        basis = mel_filters(sr, n_fft, n_mels)
    M = np.zeros((n_mels, n_frames), dtype=np.float32)
    for m in range(n_mels):
        nz = np.nonzero(basis[m])[0]
        acc = np.zeros(n_frames, dtype=np.float32)
        for f in nz:  # ascending bin order, float32 multiply then float32 add
            acc = (acc + (basis[m, f] * S[f]).astype(np.float32)).astype(np.float32)
        M[m] = acc
    # power_to_db under numpy 1.26 (the reference's pin): array ops stay float32, scalar-only expressions
    # (the reference level) are float64 and are rounded to float32 when they meet the array
    amin = np.float32(1e-10)
    ref = np.max(M)
    log_spec = (np.float32(10.0) * np.log10(np.maximum(amin, M))).astype(np.float32)
    ref_db = np.float32(10.0 * np.log10(max(1e-10, float(ref))))
    log_spec = (log_spec - ref_db).astype(np.float32)
    return np.maximum(log_spec, np.float32(float(log_spec.max()) - 80.0))


# --------------------------------------------------------------------------- stage 1 glue

def audio_to_spectrogram(audio, n_filters, filterbank, coefs=None):
    """create_dataset.py:39-78."""
    if filterbank == "mel":
        spec_db = mel_power_db(audio, n_filters, basis=coefs)        # :44-48
    else:
        hop_time = len(audio) / (SAMPLE_RATE * TIME_BINS)            # :50
        spec = gtgram(audio, SAMPLE_RATE, 0.025, hop_time, n_filters, 50, coefs=coefs)  # :51-58
        spec_db = 20 * np.log10(spec + 1e-9)                         # :59
        spec_db = np.maximum(spec_db, spec_db.max() - 80.0)          # :60
    lo = spec_db.min()                                               # :62
    hi = spec_db.max()                                               # :63
    if float(hi - lo) < 1e-8:                                        # :64-65
        return np.zeros((n_filters, TIME_BINS), dtype=np.float32)
    # :67 - the denominator is a scalar-only expression (float64 under numpy 1.26), then meets the array's dtype
    spec_norm = (spec_db - lo) / spec_db.dtype.type(float(hi - lo) + 1e-8)
    if spec_norm.shape[1] != TIME_BINS:                              # :69-72
        spec_norm = zoom(spec_norm, (1, TIME_BINS / spec_norm.shape[1]), order=1)
    return spec_norm[:, :TIME_BINS]                                  # :78


def hysteresis_encode(spectrogram, thresholds=SPIKE_THRESHOLDS, gap=HYSTERESIS_GAP):
    """create_dataset.py:81-98 restated: one Schmitt trigger per (channel, threshold); the
    level (not the edge) of trigger t_idx at time bin b lands in column b*K + t_idx,
    thresholds visited in DESCENDING order.  Comparisons are strict and done in the
    spectrogram's own dtype (float32 for mel, float64 for gammatone)."""
    spec = np.asarray(spectrogram)
    C, NB = spec.shape
    order = sorted(thresholds, reverse=True)
    K = len(order)
    out = np.zeros((C, NB * K), dtype=np.uint8)
    for k, thr in enumerate(order):
        lower = thr - gap
        on = np.zeros(C, dtype=bool)
        for b in range(NB):
            col = spec[:, b]
            on = np.where(on, ~(col < lower), col > thr)
            out[:, b * K + k] = on
    return out


def redundancy(spike_train, factor):
    """create_dataset.py:101-104."""
    return np.repeat(spike_train, factor, axis=0)


def utterance_to_spikes(audio, n_filters=128, filterbank="gammatone", coefs=None, redundancy_factor=1):
    """create_dataset.py:148-158 for one utterance -> uint8[C*R, 400]."""
    spec = audio_to_spectrogram(audio, n_filters, filterbank, coefs=coefs)
    return redundancy(hysteresis_encode(spec), redundancy_factor), spec


def w_critico(k, theta, refractory, spike_data):
    """extract_lsm_features.py:33-60."""
    n = min(500, len(spike_data))
    total = sum(int(np.sum(s)) for s in spike_data[:n])
    elems = sum(s.shape[0] * s.shape[1] for s in spike_data[:n])
    if elems == 0:
        return 0.007
    avg_i = total / elems
    beta = k / 2
    if beta == 0:
        return 0.007
    return (theta - 2 * avg_i * refractory) / beta


# --------------------------------------------------------------------------- reservoir (frozen spec)
# snnpy.snn.SNN.simulate / extract_features_from_spikes (call sites
# extract_lsm_features.py:79-83).  Semantics = DESIGN.md "Reservoir spec" R1-R10.

def simulate(x, w_rows, w_cols, w_q, w_shift, in_rowptr, in_col, in_val, leak, theta, refractory, w_val=None):
    """One utterance.  x: uint8[C,T].  Recurrent weights as COO-by-row CSR pieces
    (postsynaptic row pointer `w_rows` int[N+1], presynaptic `w_cols`, integer weights
    `w_q` meaning w = w_q * 2**-w_shift).  Input map CSR: neuron i sums in_val[p]*x[in_col[p],t]
    for p in in_rowptr[i]:in_rowptr[i+1] (ascending).  Returns raster uint8[T,N].
    w_val (float64 per edge): strict reservoirs, SURVEY.md 8c S3/S6 - the recurrent current of neuron i is the fp64 sum of the
    weights of its spiking presynaptic neurons added one by one in ascending presynaptic index."""
    C, T = x.shape
    N = len(w_rows) - 1
    V = np.zeros(N)
    ref = np.zeros(N, dtype=np.int64)
    s_prev = np.zeros(N, dtype=bool)
    raster = np.zeros((T, N), dtype=np.uint8)
    scale = 2.0 ** (-w_shift)
    # dense integer matrix: exact integer sums, order-free by construction
    Wq = np.zeros((N, N), dtype=np.int64)
    for i in range(N):
        Wq[i, w_cols[w_rows[i]:w_rows[i + 1]]] = w_q[w_rows[i]:w_rows[i + 1]]
    for t in range(T):
        if w_val is None:
            i_rec = (Wq[:, s_prev].sum(axis=1)).astype(np.float64) * scale
        else:
            i_rec = np.zeros(N)
            if s_prev.any():
                for i in range(N):
                    acc = 0.0
                    for p in range(w_rows[i], w_rows[i + 1]):                   # columns ascend inside a row
                        if s_prev[w_cols[p]]:
                            acc = acc + float(w_val[p])
                    i_rec[i] = acc
        i_in = np.zeros(N)
        for i in range(N):
            acc = 0.0
            for p in range(in_rowptr[i], in_rowptr[i + 1]):
                acc = acc + in_val[p] * (1.0 if x[in_col[p], t] else 0.0)   # level signal: non-zero = on
            i_in[i] = acc
        cur = i_in + i_rec
        active = ref == 0
        Vn = (V - leak * V) + cur
        V = np.where(active, Vn, 0.0)
        ref = np.where(active, ref, ref - 1)
        fire = active & (V >= theta)
        V = np.where(fire, 0.0, V)
        ref = np.where(fire, refractory, ref)
        raster[t] = fire
        s_prev = fire
    return raster


FEATURE_KEYS = ['spike_counts', 'spike_variances', 'mean_spike_times', 'first_spike_times',
                'last_spike_times', 'mean_isi', 'isi_variances', 'burst_counts']


def features_from_raster(raster, out_idx, refractory):
    """Eight per-output-neuron statistics (keys: extract_lsm_features.py:20-22), defined
    through integer sufficient statistics and ONE rounding each."""
    T = raster.shape[0]
    nan = float("nan")
    f = {k: np.full(len(out_idx), nan) for k in FEATURE_KEYS}
    for o, n in enumerate(out_idx):
        times = np.nonzero(raster[:, n])[0].astype(np.int64)
        c = len(times)
        p = c / T
        f['spike_counts'][o] = float(c)
        f['spike_variances'][o] = p * (1.0 - p)
        f['burst_counts'][o] = 0.0
        if c >= 1:
            f['mean_spike_times'][o] = float(times.sum()) / float(c)
            f['first_spike_times'][o] = float(times[0])
            f['last_spike_times'][o] = float(times[-1])
        if c >= 2:
            isi = np.diff(times)
            n_isi = c - 1
            s1 = int(isi.sum())
            s2 = int((isi * isi).sum())
            f['mean_isi'][o] = float(s1) / float(n_isi)
            f['isi_variances'][o] = float(n_isi * s2 - s1 * s1) / float(n_isi * n_isi)
            f['burst_counts'][o] = float(int((isi <= refractory + 1).sum()))
    return f


# ---------------------------------------------------------------------------------------------
# Ingest: sample-rate conversion.  The reference resamples inside librosa.load
# (/root/reference/create_dataset.py:26); librosa's res_type="polyphase" is scipy.signal.resample_poly, whose arithmetic
# (scipy/signal/_upfirdn_apply.pyx _apply_impl) is restated here sample by sample: for every kept output one float32 multiply and
# one float32 add per tap, in ascending input order, zero padding outside the signal.  Pinned: equals scipy's output bit for bit
# (tests/test_oracle_frontend.py).  Small signals only - this is a Python loop.
def resample_poly_ref(x, up, down, taps, hpp, n_pre_remove, n_out):
    """x float32[n_in]; (up, down, taps[up][hpp], hpp, n_pre_remove, n_out) as ingest.polyphase_design returns them."""
    x = np.asarray(x, dtype=np.float32)
    n_in = len(x)
    out = np.zeros(n_out, dtype=np.float32)
    for o in range(n_out):
        pos = (o + n_pre_remove) * down
        t, xi = pos % up, pos // up
        acc = np.float32(0.0)
        for j in range(hpp):
            k = xi - hpp + 1 + j
            if 0 <= k < n_in:
                acc = np.float32(acc + np.float32(x[k] * taps[t, j]))
        out[o] = acc
    return out


# ---------------------------------------------------------------------------------------------
# Reservoir construction, second implementation (the product's builder, reservoir.py, had no independent
# counterpart).  The reference builds its liquid with SNN(simulation_params=...) (/root/reference/extract_lsm_features.py:164-188,
# snnpy un-vendored); the frozen spec R1-R7 is restated here with plain Python loops over neurons and edges - no
# adjacency matrix, no vectorised draws beyond the ones the spec names - and tests/test_oracle_reservoir.py requires every
# array of reservoir.build_reservoir to equal this one's.
def build_reservoir_ref(num_neurons, k, p, mean_weight, weight_variance, num_inputs, num_outputs, theta, leak_coefficient,
                        leak_variance_divisor=None, input_gain=None, seed=42, w_shift=24):
    n = int(num_neurons)
    rs = np.random.RandomState(seed)                                   # R1: one generator, fixed order of use
    half = k // 2
    nb = [set() for _ in range(n)]                                     # R2: ring lattice, k/2 neighbours each side
    for u in range(n):
        for j in range(1, half + 1):
            v = (u + j) % n
            nb[u].add(v)
            nb[v].add(u)
    if p > 0:
        for j in range(1, half + 1):                                   # right-hand edges (u, u+j), one distance at a time
            draw = rs.random_sample(n)
            for u in range(n):
                if not draw[u] < p:
                    continue
                v = (u + j) % n
                if len(nb[u]) >= n - 1 or v not in nb[u]:
                    continue
                while True:                                            # uniformly random non-neighbour
                    w = int(rs.randint(n))
                    if w != u and w not in nb[u]:
                        break
                nb[u].discard(v); nb[v].discard(u)
                nb[u].add(w); nb[w].add(u)
    edges = [(i, j) for i in range(n) for j in sorted(nb[i])]          # directed, (post, pre) row-major
    sd = abs(mean_weight) / weight_variance if weight_variance else 0.0
    if sd > 0:                                                         # R3: one normal draw per directed edge, rounded to 2^-w_shift
        w = rs.normal(mean_weight, sd, size=len(edges))
    else:
        w = np.full(len(edges), float(mean_weight))
    w_q = [int(np.rint(x * float(1 << w_shift))) for x in w]
    w_rowptr = [0]
    for i in range(n):
        w_rowptr.append(w_rowptr[-1] + len(nb[i]))
    if num_inputs <= n:                                                # R4: input row r -> one neuron, distinct while rows <= n
        in_neuron = [int(v) for v in rs.permutation(n)[:num_inputs]]
    else:
        in_neuron = [int(v) for v in rs.randint(n, size=num_inputs)]
    gain = float(theta if input_gain is None else input_gain)
    in_rowptr, in_col = [0], []
    for i in range(n):
        rows = [r for r in range(num_inputs) if in_neuron[r] == i]     # ascending input row inside a neuron's list
        in_col += rows
        in_rowptr.append(len(in_col))
    n_out = min(int(num_outputs), n)
    out_idx = sorted(int(v) for v in rs.permutation(n)[:n_out])        # R5
    if leak_variance_divisor:                                          # R7
        leak = np.clip(rs.normal(leak_coefficient, leak_coefficient / leak_variance_divisor, size=n), 0.0, 1.0)
    else:
        leak = np.full(n, float(leak_coefficient))
    return dict(w_rowptr=np.array(w_rowptr, np.int32), w_col=np.array([j for _, j in edges], np.int32), w_q=np.array(w_q, np.int32),
                in_rowptr=np.array(in_rowptr, np.int32), in_col=np.array(in_col, np.int32), in_val=np.full(len(in_col), gain),
                out_idx=np.array(out_idx, np.int32), leak=np.asarray(leak, np.float64))
