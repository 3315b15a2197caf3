"""CPU: the oracle against the golden vectors minted from the reference's own functions
(tests/golden/make_golden.py) and against the installed scipy/numpy (the reference's deps)."""
import numpy as np
import pytest

from lsm_speech_classifier_b200 import filterbank
from oracle import coracle, pyref

THR = [0.70, 0.80, 0.90, 0.95]
GAP = 0.1


def test_encoder_kats_from_reference(golden):
    g = golden("encoder_kats.npz")
    assert list(g["thresholds"]) == THR and float(g["gap"]) == GAP
    assert np.array_equal(pyref.hysteresis_encode(g["spec64"], THR, GAP), g["spikes64"])
    assert np.array_equal(pyref.hysteresis_encode(g["spec32"], THR, GAP), g["spikes32"])
    assert np.array_equal(coracle.hysteresis_encode(g["spec64"], THR, GAP), g["spikes64"])
    assert np.array_equal(coracle.hysteresis_encode(g["spec64"][:5], THR, GAP, redundancy=3), g["redundancy3"])
    assert np.array_equal(pyref.redundancy(g["spikes64"][:5], 3), g["redundancy3"])


def test_encoder_hand_derived():
    # ramp 0 -> 1 over 100 bins: trigger k turns on at the first bin strictly above its threshold
    ramp = np.linspace(0.0, 1.0, 100)[None, :]
    s = coracle.hysteresis_encode(ramp, THR, GAP)[0].reshape(100, 4)
    for k, thr in enumerate(sorted(THR, reverse=True)):
        first = int(np.argmax(ramp[0] > thr))
        assert s[:first, k].sum() == 0 and s[first:, k].all()
    assert coracle.hysteresis_encode(np.zeros((3, 100)), THR, GAP).sum() == 0
    assert coracle.hysteresis_encode(np.ones((3, 100)), THR, GAP).all()
    # level exactly on a threshold never switches on (strict >), exactly on the lower bound never off (strict <)
    assert coracle.hysteresis_encode(np.full((1, 100), 0.95), THR, GAP)[0].reshape(100, 4)[:, 0].sum() == 0


def test_strides_and_zoom_kats():
    assert filterbank.gtgram_strides(16000, 0.025, 0.01, 16000) == (400, 160, 98)
    from scipy.ndimage import zoom
    rng = np.random.default_rng(1)
    for ncol, dt in ((98, np.float64), (101, np.float32)):
        v = rng.random((16, ncol)).astype(dt)
        i0, f = filterbank.zoom_table(ncol, 100)
        want = zoom(v, (1, 100 / ncol), order=1)
        vd = v.astype(np.float64)
        nxt = np.where(i0 + 1 < ncol, i0 + 1, i0)
        got = vd[:, i0] * (1.0 - f) + np.where(i0 + 1 < ncol, vd[:, nxt] * f, 0.0)
        got = np.where(i0 + 1 < ncol, got, vd[:, i0] * (1.0 - f)).astype(dt)
        assert want.shape == (16, 100)
        assert np.array_equal(got, want)
        assert np.array_equal(want[:, 0], v[:, 0]) and np.array_equal(want[:, -1], v[:, -1])


def test_design_table_close_to_independent_derivation():
    a = filterbank.gammatone_coefs(16000, 128, 50)
    b = pyref.gammatone_design(16000, 128, 50)
    assert a.shape == (128, 10)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=0)
    cf = filterbank.centre_freqs(16000, 128, 50)
    assert abs(cf[-1] - 50.0) < 1e-9 and np.all(np.diff(cf) < 0)
    # unit gain at the centre frequency after /gain (4 cascaded biquads)
    for ch in (0, 31, 64, 127):
        A0, A11, A12, A13, A14, A2, B0, B1, B2, gain = a[ch]
        z = np.exp(-2j * np.pi * cf[::-1][ch] / 16000)
        h = 1.0
        for A1 in (A11, A12, A13, A14):
            h *= (A0 + A1 * z + A2 * z * z) / (B0 + B1 * z + B2 * z * z)
        assert abs(abs(h) / gain - 1.0) < 1e-9


def test_c_oracle_gammatone_vs_golden(golden):
    g = golden("frontend_gammatone.npz")
    pcm, coefs = g["pcm"], g["coefs"]
    want = np.unpackbits(g["spikes_packed"], axis=-1)[:, :, :400]
    i0, f = filterbank.zoom_table(98, 100)
    got, spec = coracle.gammatone_encode(pcm, coefs, 400, 160, 100, i0, f, THR, GAP, want_spec=True)
    # spikes: bit-exact against (scipy restatement + the reference's own encoder)
    assert np.array_equal(got, want)
    # spectrogram: identical up to log10's last bit (numpy uses the platform libm; the oracle's is fixed)
    np.testing.assert_allclose(spec[:2], g["spec_norm"], rtol=0, atol=1e-12)
    assert got[3].sum() == 0 and np.all(spec[3] == 0)          # silent clip (create_dataset.py:64-65)
    assert got[4].sum() > 0                                     # zero-padded short clip
    # threads must not change a bit
    assert np.array_equal(coracle.gammatone_encode(pcm, coefs, 400, 160, 100, i0, f, THR, GAP, nthreads=1), got)


def test_pyref_gammatone_pinned_pieces():
    """lfilter recurrence and window mean, restated scalar-by-scalar, against scipy/numpy."""
    from scipy.signal import lfilter
    rng = np.random.default_rng(3)
    x = rng.standard_normal(2000).astype(np.float32)
    c = filterbank.gammatone_coefs(16000, 8, 50)[5]
    y = lfilter([c[0], c[2], c[5]], [c[6], c[7], c[8]], x)
    z0 = z1 = 0.0
    out = np.empty(len(x))
    for n, xn in enumerate(x.astype(np.float64)):
        yn = z0 + c[0] * xn
        z0 = (z1 + xn * c[2]) - yn * c[7]
        z1 = xn * c[5] - yn * c[8]
        out[n] = yn
    assert np.array_equal(y, out)
    xe = rng.random((4, 1000))
    seg = xe[:, 160 + np.arange(400)]
    m = seg.mean(1)
    for r in range(4):
        acc = 0.0
        for v in xe[r, 160:560]:
            acc = acc + v
        assert acc / 400 == m[r]


def test_log10_is_a_faithful_libm(golden):
    xs = np.concatenate([np.random.default_rng(0).uniform(1e-9, 1.0, 3000),
                         10 ** np.random.default_rng(1).uniform(-9, 3, 3000), [1.0, 1e-9, 10.0, 0.5]])
    mine = coracle.log10(xs)
    ref = np.log10(xs)
    ulp = np.abs(mine - ref) / np.spacing(np.maximum(np.abs(ref), 1e-300))
    assert ulp.max() <= 2.0
    assert coracle.log10(np.array([1.0]))[0] == 0.0


def test_w_critico_kats(golden):
    g = golden("w_critico_kats.npz")
    for i in range(4):
        assert pyref.w_critico(200, 2.0, 2, list(g[f"d{i}"])) == g["answers"][i]
    assert g["answers"][3] == (2.0 - 2 * 0.1 * 2) / 100
    assert pyref.w_critico(0, 2.0, 2, list(g["d0"])) == 0.007
    assert pyref.w_critico(200, 2.0, 2, []) == 0.007


def test_constant_divisor_division_is_ieee_exact():
    """K1 divides by the per-channel gain with q0 = x*r, e = fma(-g,q0,x), q = fma(e,r,q0) (r = 1/g).
    Markstein: correctly rounded when r = RN(1/g) and the residual cannot underflow (|x| >= 2^-900).
    Checked here against IEEE division for every design gain over random mantissas and exponents."""
    gains = filterbank.gammatone_coefs(16000, 128, 50)[:, 9]
    rng = np.random.default_rng(7)
    for g in list(gains[::9]) + [gains[0], gains[-1], 3.0, 1.0 / 3.0, 0.1]:
        mant = rng.uniform(1.0, 2.0, 200000)
        expo = rng.integers(-890, 300, 200000)
        x = np.ldexp(mant, expo) * rng.choice([-1.0, 1.0], 200000)
        x[:1000] = np.ldexp(1.0 + np.arange(1000) * 2.0 ** -52, -3)         # neighbouring mantissas
        x[1000:1003] = [0.0, -0.0, 1.0]
        assert coracle.check_const_division(x, g) <= 1                        # only -0.0 may differ (sign of zero)
    # below the window the quotient need not be exact, but its square is exactly 0 either way
    tiny = np.ldexp(rng.uniform(1.0, 2.0, 1000), -901 - rng.integers(0, 120, 1000))
    assert np.all((tiny / gains.min()) ** 2 == 0.0)


def test_c_oracle_mel_vs_scipy_restatement(golden):
    """Mel branch: the C oracle (fixed radix-2 FFT order, deterministic log10, numpy-1.26 scalar semantics) against
    the goldens minted by the scipy.fft restatement + the reference's own encoder."""
    g = golden("frontend_mel64.npz")
    pcm = golden("frontend_gammatone.npz")["pcm"][:3]
    basis = filterbank.mel_basis(16000, 2048, 64)
    tw, tw2 = filterbank.fft_tables(2048)
    i0, f = filterbank.zoom_table(101, 100)
    got, spec = coracle.mel_encode(pcm, basis, filterbank.hann_periodic(2048), tw, tw2, filterbank.pack_mel_basis(basis),
                                   160, 100, i0, f, THR, GAP, want_spec=True)
    want = np.unpackbits(g["spikes_packed"], axis=-1)[:, :, :400]
    assert (got == want).mean() >= 0.999            # pocketfft vs our FFT order, np.log10 vs fixed log10: near-ties only
    assert np.array_equal(got, want)                # ... and on these clips there are none
    np.testing.assert_allclose(spec[:1], g["spec_norm"], rtol=0, atol=2e-6)
    # the mel basis restated scalar-by-scalar equals the vectorised product table
    assert np.array_equal(basis, pyref.mel_filters(16000, 2048, 64))
    # 256 bands: narrow low-frequency triangles may be empty -> silent rows, never NaN
    b256 = filterbank.mel_basis(16000, 2048, 256)
    s256 = coracle.mel_encode(pcm[:1], b256, filterbank.hann_periodic(2048), tw, tw2, filterbank.pack_mel_basis(b256),
                              160, 100, i0, f, THR, GAP)
    assert s256.shape == (1, 256, 400)


def test_polyphase_resampler_restatement_is_scipy_bit_for_bit():
    """ingest.polyphase_design + pyref.resample_poly_ref (what the GPU kernel repeats) against scipy.signal.resample_poly =
    librosa.resample(res_type="polyphase"): the resampling step of load_audio_file (create_dataset.py:26)."""
    from scipy.signal import resample_poly
    from lsm_speech_classifier_b200 import ingest
    from oracle import pyref
    rs = np.random.RandomState(1)
    for sr, n in ((44100, 1500), (48000, 1400), (8000, 700), (22050, 900), (11025, 400), (16001, 300), (44100, 1), (48000, 7)):
        x = (rs.standard_normal(n) * 0.3).astype(np.float32)
        want = resample_poly(x, 16000, sr)
        got = pyref.resample_poly_ref(x, *ingest.polyphase_design(n, 16000, sr))
        assert want.dtype == np.float32 and np.array_equal(got, want), sr


def _write_wav(path, tag, bits, rate, frames, extensible=False):
    """frames: float array [n, ch] in [-1, 1); returns the float32 array a libsndfile-style decoder yields."""
    import struct
    n, ch = frames.shape
    if tag == 1 and bits == 8:
        q = np.clip(np.round(frames * 128.0) + 128, 0, 255).astype(np.uint8); body = q.tobytes(); dec = (q.astype(np.float32) - 128.0) / 128.0
    elif tag == 1 and bits == 16:
        q = np.clip(np.round(frames * 32768.0), -32768, 32767).astype("<i2"); body = q.tobytes(); dec = q.astype(np.float32) / 32768.0
    elif tag == 1 and bits == 24:
        q = np.clip(np.round(frames * 8388608.0), -8388608, 8388607).astype(np.int32)
        b = np.stack([(q & 0xFF), (q >> 8) & 0xFF, (q >> 16) & 0xFF], axis=-1).astype(np.uint8); body = b.tobytes(); dec = q.astype(np.float32) / 8388608.0
    elif tag == 1 and bits == 32:
        q = np.clip(np.round(frames.astype(np.float64) * 2147483648.0), -2147483648, 2147483647).astype("<i4"); body = q.tobytes()
        dec = (q.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif tag == 3 and bits == 32:
        q = frames.astype("<f4"); body = q.tobytes(); dec = q.astype(np.float32)
    else:
        q = frames.astype("<f8"); body = q.tobytes(); dec = q.astype(np.float32)
    align = ch * bits // 8
    if extensible:
        fmt = struct.pack("<HHIIHHHHIH", 0xFFFE, ch, rate, rate * align, align, bits, 22, bits, 0, tag) + b"\x00" * 14
    else:
        fmt = struct.pack("<HHIIHH", tag, ch, rate, rate * align, align, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"data" + struct.pack("<I", len(body)) + body
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)
    return dec.reshape(n, ch)


def test_wav_reader_formats(tmp_path):
    """ingest.read_wav: the decode half of librosa.load (create_dataset.py:26) for RIFF/WAVE files."""
    from lsm_speech_classifier_b200 import ingest
    rs = np.random.RandomState(2)
    for tag, bits, ch, ext in ((1, 8, 1, False), (1, 16, 1, False), (1, 16, 2, False), (1, 24, 2, False), (1, 32, 1, False),
                               (3, 32, 2, False), (3, 64, 1, False), (1, 16, 2, True), (3, 32, 1, True)):
        frames = np.clip(rs.standard_normal((501, ch)) * 0.3, -0.99, 0.99)
        p = tmp_path / f"t{tag}_{bits}_{ch}_{int(ext)}.wav"
        want = _write_wav(p, tag, bits, 22050, frames, ext)
        got, rate = ingest.read_wav(p)
        assert rate == 22050 and got.dtype == np.float32 and np.array_equal(got, want), (tag, bits, ch, ext)
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFF\x00\x00\x00\x00WAVEnope")
    with pytest.raises(ValueError):
        ingest.read_wav(bad)
    with pytest.raises(ValueError):
        ingest.read_wav(__file__)
