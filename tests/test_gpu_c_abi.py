"""GPU: liblsmb200.so driven from plain C (gcc, no Python in the data path) gives the Python path's bytes."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_consumer_matches_python_path(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lsm_speech_classifier_b200 import _lib, synth
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm
    from lsm_speech_classifier_b200.frontend import Frontend
    from oracle import coracle

    exe = tmp_path / "c_abi_consumer"
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_consumer.c"),
                           "-o", str(exe), "-L", os.path.join(ROOT, "lsm_speech_classifier_b200"), "-llsmb200", "-lm",
                           "-Wl,-rpath," + os.path.join(ROOT, "lsm_speech_classifier_b200")])
    pcm, _ = synth.synth_dataset(3, 3)
    fe = Frontend(128, "gammatone")
    X = fe.encode(pcm)
    lsm = build_lsm(X, 0.6, verbose=False)
    r = lsm.reservoir
    with open(tmp_path / "in.bin", "wb") as f:
        np.array([len(pcm), r.num_neurons, len(r.w_q), len(r.in_col), len(r.out_idx)], np.int32).tofile(f)
        for a, dt in ((pcm, np.float32), (r.w_rowptr, np.int32), (r.w_col, np.int32), (r.w_q, np.int32), (r.in_rowptr, np.int32),
                      (r.in_col, np.int32), (r.in_val, np.float64), (r.leak, np.float64), (r.out_idx, np.int32),
                      (np.array([r.theta]), np.float64)):
            np.ascontiguousarray(a, dtype=dt).tofile(f)
    out = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "fused 1" in out.stdout
    raw = np.fromfile(tmp_path / "out.bin", dtype=np.uint8)
    nc = 128 * 10 * 8
    coefs = raw[:nc].view(np.float64).reshape(128, 10)
    spikes = raw[nc:nc + X.size].reshape(X.shape)
    feats = raw[nc + X.size:].view(np.float64).reshape(len(pcm), -1)
    # the C library's own design table (host libm) agrees with numpy's to the last bits ...
    np.testing.assert_allclose(coefs, fe.table, rtol=1e-12)
    # ... and what the C program computed is exactly what the oracle gives for THAT table
    want_spk = coracle.gammatone_encode(pcm, coefs, 400, 160, 100, fe.zoom_i0, fe.zoom_f, [0.70, 0.80, 0.90, 0.95], 0.1)
    assert np.array_equal(spikes, want_spk)
    want_f, _ = coracle.reservoir_run(r, want_spk, _lib.feature_mask(FEATURE_SETS["original"]), True, False)
    assert np.array_equal(feats, want_f)
