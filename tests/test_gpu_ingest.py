"""GPU: the ingest step - resampling to 16 kHz on the device and load_audio_file (create_dataset.py:22-36).

The reference resamples inside librosa.load; librosa's res_type="polyphase" is scipy.signal.resample_poly, which this kernel
reproduces bit for bit (librosa's default soxr_hq recipe cannot be restated: non-16 kHz files are pinned to scipy, not librosa)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lsm_speech_classifier_b200 import _lib
    return _lib.context(0)


@pytest.mark.parametrize("sr", [44100, 48000, 8000, 22050, 11025, 32000, 16001])
def test_resampler_equals_scipy_resample_poly(env, sr):
    import torch
    from scipy.signal import resample_poly
    from lsm_speech_classifier_b200 import ingest
    rs = np.random.RandomState(sr)
    x = (rs.standard_normal((5, sr + 37)) * 0.3).astype(np.float32)
    x[1] = 0.0; x[2] = 1.0; x[3, :] = 0.0; x[3, 100] = 1.0                       # silence, DC, one impulse
    want = resample_poly(x, 16000, sr, axis=-1)
    got = ingest.resample_poly(x, sr, 16000)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want)
    d = ingest.resample_poly(torch.from_numpy(x).cuda(), sr, 16000)              # torch in, torch out
    assert d.is_cuda and np.array_equal(d.cpu().numpy(), want)
    one = ingest.resample_poly(x[0], sr, 16000)                                  # one signal
    assert np.array_equal(one, want[0])
    for n in (1, 2, 17):                                                         # shorter than the filter
        assert np.array_equal(ingest.resample_poly(x[0, :n], sr, 16000), resample_poly(x[0, :n], 16000, sr))
    assert ingest.resample_poly(x, 16000, 16000) is x


def test_load_audio_file_contract(env, tmp_path, capsys):
    """float32[16000] at 16 kHz whatever the file holds; None and the reference's message for unreadable files."""
    from scipy.signal import resample_poly
    from test_oracle_frontend import _write_wav
    from lsm_speech_classifier_b200.create_dataset import load_audio_file
    rs = np.random.RandomState(9)
    # 44.1 kHz stereo float, longer than a second: channel mean, resample, truncate
    frames = np.clip(rs.standard_normal((60000, 2)) * 0.2, -0.99, 0.99)
    dec = _write_wav(tmp_path / "a.wav", 3, 32, 44100, frames)
    got = load_audio_file(tmp_path / "a.wav")
    mono = dec.mean(axis=1, dtype=np.float32)
    want = resample_poly(mono[:int(np.ceil(1.05 * 44100))], 160, 441)[:16000]
    assert got.dtype == np.float32 and got.shape == (16000,) and np.array_equal(got, want)
    # the first second does not depend on how much of the tail was converted
    full = resample_poly(mono, 160, 441)[:16000]
    assert np.array_equal(got[:15000], full[:15000])
    # 16 kHz PCM16 mono, short: exactly sample / 32768, zero padded (the Speech Commands case; no resampler involved)
    dec = _write_wav(tmp_path / "b.wav", 1, 16, 16000, np.clip(rs.standard_normal((9000, 1)) * 0.2, -0.99, 0.99))
    got = load_audio_file(tmp_path / "b.wav")
    assert np.array_equal(got[:9000], dec[:, 0]) and not got[9000:].any() and got.shape == (16000,)
    # 8 kHz 24-bit PCM: upsampled
    dec = _write_wav(tmp_path / "c.wav", 1, 24, 8000, np.clip(rs.standard_normal((5000, 1)) * 0.2, -0.99, 0.99))
    got = load_audio_file(tmp_path / "c.wav")
    assert np.array_equal(got[:10000], resample_poly(dec[:, 0], 2, 1)) and not got[10000:].any()
    # not audio
    (tmp_path / "d.wav").write_bytes(b"not a wav file at all")
    assert load_audio_file(tmp_path / "d.wav") is None
    assert "Error loading" in capsys.readouterr().out
