"""CPU: host-side arithmetic the CUDA path relies on, checked with numpy against the oracle / plain IEEE arithmetic.
No GPU, no compute calls into the library."""
import numpy as np
import pytest


def test_lean_current_identity_is_bit_exact():
    """reservoir_core.cuh lean_current(): I = I_in + acc * 2^-w (spec R6, one rounding) is formed as D - c, where the low word
    of the double D holds acc ^ 0x80000000 and c = 2^(52-w) + 2^(31-w) - I_in.  Same bits for every accumulator value."""
    rng = np.random.default_rng(3)
    w = 24
    acc = np.concatenate([rng.integers(-2**31, 2**31, 200000), [0, 1, -1, 2**31 - 1, -2**31, 195000, 5 * 273000]]).astype(np.int64)
    hi_magic = np.uint64((1075 - w) << 20) << np.uint64(32)
    lo = (acc.astype(np.int64) & 0xFFFFFFFF).astype(np.uint64) ^ np.uint64(0x80000000)
    D = (hi_magic | lo).view(np.float64)
    c_off = 2.0 ** (52 - w) + 2.0 ** (31 - w)
    for i_in in (0.0, 2.0, 0.5, 127.0, -3.0):
        assert (i_in * 2.0 ** w) == np.floor(i_in * 2.0 ** w)            # the host's exactness condition on the gain
        c = c_off - i_in
        assert c + i_in == c_off                                          # exact: the folded constant loses nothing
        want = i_in + acc.astype(np.float64) * 2.0 ** -w                  # oracle: add64(i_in, mul64((double)acc, scale))
        got = D - c
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), i_in


def test_refractory_nibble_arithmetic():
    """4-bit refractory counters, 8 per word: decrement-if-nonzero and set-on-fire, against the per-neuron rule of spec R6."""
    rng = np.random.default_rng(4)
    R = 2
    ref = np.zeros(8, dtype=np.int64)
    word = 0
    for _ in range(2000):
        nz = (((word & 0x77777777) + 0x77777777) | word) & 0x88888888
        active = np.array([(nz >> (4 * j + 3)) & 1 == 0 for j in range(8)])
        assert np.array_equal(active, ref == 0)
        fire = active & (rng.random(8) < 0.3)
        word = (word - (nz >> 3)) + sum((1 << (4 * j)) for j in range(8) if fire[j]) * R
        ref = np.where(fire, R, np.where(ref == 0, 0, ref - 1))
        assert [(word >> (4 * j)) & 15 for j in range(8)] == ref.tolist()


def test_speculative_arrangement_is_the_same_filter():
    """gt_filter_fast's arrangement (normalised numerators, direct form, block energy sums, gain at the window level) against
    the oracle's reference-order spectrogram: the normalised planes differ by rounding noise only (here without FMA, in numpy)."""
    from lsm_speech_classifier_b200 import synth, filterbank as fb
    from oracle import coracle
    pcm, _ = synth.synth_dataset(3, 2)
    pcm = np.concatenate([pcm, np.full((1, 16000), 0.25, np.float32)])
    coefs = fb.gammatone_coefs(16000, 128, 50)
    nwin, hop, ncols = fb.gtgram_strides(16000, 0.025, 0.01, 16000)
    zi0, zf = fb.zoom_table(ncols, 100)
    _, spec = coracle.gammatone_encode(pcm, coefs, nwin, hop, 100, zi0, zf, [0.7, 0.8, 0.9, 0.95], 0.1, want_spec=True)
    U, C = pcm.shape[0], 128
    c = [coefs[:, 1 + i] / coefs[:, 0] for i in range(4)]
    a1, a2 = coefs[:, 7] / coefs[:, 6], coefs[:, 8] / coefs[:, 6]
    G = (coefs[:, 0] / coefs[:, 6]) ** 4 / coefs[:, 9]
    x = pcm.astype(np.float64)
    yp = [np.zeros((U, C)) for _ in range(4)]
    yq = [np.zeros((U, C)) for _ in range(4)]
    xp = np.zeros((U, 1))
    n_used = (ncols - 1) * hop + nwin
    sub = np.zeros((U, C, n_used // 80))
    acc = np.zeros((U, C))
    for n in range(n_used):
        inp, pin = np.broadcast_to(x[:, n:n + 1], (U, C)), np.broadcast_to(xp, (U, C))
        for k in range(4):
            y = (inp + c[k] * pin - a2 * yq[k]) - a1 * yp[k]
            pin, yq[k], yp[k], inp = yp[k], yp[k], y, y
        xp = x[:, n:n + 1]
        acc = acc + inp * inp
        if n % 80 == 79:
            sub[:, :, n // 80], acc = acc, np.zeros((U, C))
    db = np.stack([20 * np.log10(np.sqrt(sub[:, :, 2 * col:2 * col + 5].sum(-1) * (G * G / nwin)) + 1e-9) for col in range(ncols)], axis=-1)
    mx = db.max((1, 2), keepdims=True)
    d = np.maximum(db, mx - 80)
    mn = d.min((1, 2), keepdims=True)
    den = mx - mn + 1e-8
    norm = (d - mn) / den
    z = norm[:, :, zi0] * (1 - zf) + norm[:, :, np.minimum(zi0 + 1, ncols - 1)] * zf
    diff_db = (np.abs(z - spec) * den).max()
    assert diff_db < 1e-9, diff_db            # the GPU's near-tie margin is 1e-7 dB


def test_readout_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lsm_speech_classifier_b200 import _lib
    from lsm_speech_classifier_b200.readout import LogisticRegression, StandardScaler
    for cls in (StandardScaler, LogisticRegression):
        with pytest.raises(_lib.LsmError):
            cls()


def test_packed_spike_file_round_trip(tmp_path, monkeypatch):
    """create_dataset --packed: bits on disk, the reference's arrays back through load_spike_dataset."""
    import os
    from lsm_speech_classifier_b200 import create_dataset as cd, extract_lsm_features as ex
    rng = np.random.default_rng(7)
    X = (rng.random((9, 128, 400)) < 0.03).astype(np.uint8)
    X[3] = 0
    X[4] = 1
    y = np.arange(9, dtype=np.int32) % 4
    monkeypatch.chdir(tmp_path)
    cd.save_packed_spikes(cd.PACKED_FILE, X, y)
    ref = tmp_path / "ref.npz"
    np.savez_compressed(ref, X_spikes=X, y_labels=y)
    assert os.path.getsize(cd.PACKED_FILE) < os.path.getsize(ref)
    Xb, yb = ex.load_spike_dataset()                              # reference-schema file absent -> packed file
    assert Xb.dtype == np.uint8 and Xb.shape == X.shape and np.array_equal(Xb, X) and np.array_equal(yb, y) and yb.dtype == np.int32
    Xo = (rng.random((2, 3, 13)) < 0.5).astype(np.uint8)          # a step count that is not a multiple of 8
    cd.save_packed_spikes("odd.npz", Xo, np.zeros(2, np.int32))
    assert np.array_equal(cd.load_packed_spikes("odd.npz")[0], Xo)
    with pytest.raises(ValueError):
        cd.save_packed_spikes("bad.npz", X * 2, y)


def test_parallel_npz_writer_is_a_drop_in_for_savez_compressed(tmp_path):
    """npzio.savez_compressed: same container, keys, dtypes, shapes and array bytes as
    np.savez_compressed (create_dataset.py:176, extract_lsm_features.py:204), readable by np.load and zipfile."""
    import zipfile
    from lsm_speech_classifier_b200 import npzio
    rs = np.random.RandomState(0)
    X = (rs.random_sample((700, 128, 400)) < 0.04).astype(np.uint8)          # 36 MB: several deflate segments
    y = rs.randint(0, 12, 700).astype(np.int32)
    F = np.asfortranarray(rs.standard_normal((300, 2000)))
    ref, got = tmp_path / "ref.npz", tmp_path / "got"                          # extension added like numpy does
    kw = dict(X_spikes=X, y_labels=y, X_train_features=F, feature_set="original", leak_variance_divisor=None,
              empty=np.zeros((0, 5)), scalar=np.float64(3.5))
    np.savez_compressed(ref, **kw)
    npzio.savez_compressed(got, **kw)
    a, b = np.load(ref, allow_pickle=True), np.load(str(got) + ".npz", allow_pickle=True)
    assert a.files == b.files
    for k in a.files:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
        assert a[k].tobytes() == b[k].tobytes() if a[k].dtype != object else a[k].item() == b[k].item(), k
    assert b["X_train_features"].flags.f_contiguous
    with zipfile.ZipFile(str(got) + ".npz") as z:
        assert z.testzip() is None
        assert all(i.compress_type == zipfile.ZIP_DEFLATED for i in z.infolist())
    # CRC combination over many segments
    import zlib
    data = rs.bytes(5_000_000)
    crc = 0
    for lo in range(0, len(data), 700_001):
        part = data[lo:lo + 700_001]
        crc = npzio._crc32_combine(crc, zlib.crc32(part), len(part)) if lo else zlib.crc32(part)
    assert crc == zlib.crc32(data)
