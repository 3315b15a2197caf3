"""GPU: BASELINE.json's configurations inside `pytest -m gpu` (they existed only as builder-run reports).

config 2   12 classes x 1000 utterances (9600 train / 2400 test), 128-channel gammatone, `original`, multiplier 0.6:
           spike trains identical, every neuron-utterance spike count exact, all features identical, same test accuracy
config 3   multiplier sweep 0.4 .. 1.0 (gammatone-128 and mel-128 inputs): participation and rasters, GPU vs oracle
config 4   N = 4000 reservoir: raster and features identical (N = 1000 / 2500 / 256 are in test_gpu_parity.py)
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

THR, GAP = [0.70, 0.80, 0.90, 0.95], 0.1


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lsm_speech_classifier_b200 import _lib
    return _lib.context(0)


@pytest.fixture(scope="module")
def config2_pcm():
    from lsm_speech_classifier_b200 import synth
    # worker processes fork before this module touches CUDA only if it runs first; keep the pool small and fork-safe
    return synth.synth_dataset(12, 1000, workers=1 if os.environ.get("LSM_TEST_SERIAL_SYNTH") else min(16, os.cpu_count() or 1))


def test_config2_twelve_thousand_utterances_identical_to_the_oracle(env, config2_pcm):
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import StandardScaler
    from lsm_speech_classifier_b200 import _lib
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    from oracle import coracle
    pcm, labels = config2_pcm
    keys = FEATURE_SETS["original"]
    mask = _lib.feature_mask(keys)
    fe = Frontend(128, "gammatone")
    X = fe.encode(pcm)                                                     # stage 1 as create_dataset does it
    Xo = coracle.gammatone_encode(pcm, fe.table, fe.params.nwin, fe.params.hop, fe.time_bins, fe.zoom_i0, fe.zoom_f, THR, GAP)
    assert np.array_equal(X, Xo), "spike trains differ from the oracle"
    idx = np.arange(len(pcm))
    i_tr, i_te, y_tr, y_te = train_test_split(idx, labels, test_size=0.2, random_state=42, stratify=labels)   # extract_lsm_features.py:160-162
    lsm = build_lsm(X[i_tr], 0.6, verbose=False)
    F_tr = lsm.simulate_batch(X[i_tr], keys)
    F_te = lsm.simulate_batch(X[i_te], keys)
    Fo_tr, _ = coracle.reservoir_run(lsm.reservoir, Xo[i_tr], mask, True, False)
    Fo_te, _ = coracle.reservoir_run(lsm.reservoir, Xo[i_te], mask, True, False)
    n_out = lsm.num_output_neurons
    exact = (np.sum(F_tr[:, :n_out] == Fo_tr[:, :n_out]) + np.sum(F_te[:, :n_out] == Fo_te[:, :n_out])) / (F_tr[:, :n_out].size + F_te[:, :n_out].size)
    assert exact == 1.0, f"{100 * exact:.4f} % of the neuron-utterance spike counts exact (bar: 99.9 %, here all of them)"
    assert np.array_equal(F_tr, Fo_tr) and np.array_equal(F_te, Fo_te)
    # the whole path in one call (the warp-specialised kernel) gives the same rows as the two staged calls
    whole = AudioToFeatures(fe, lsm).run_host(pcm[i_te], keys)
    assert np.array_equal(whole, F_te)
    # downstream: same features, so the same classifier; check the accuracy bar on a subsample to keep the test short
    sc = StandardScaler()
    clf = LogisticRegression(random_state=42, max_iter=300)
    sub = slice(0, 2400)
    clf.fit(sc.fit_transform(F_tr[sub]), y_tr[sub])
    acc_gpu = float((clf.predict(sc.transform(F_te)) == y_te).mean())
    clf_o = LogisticRegression(random_state=42, max_iter=300)
    sco = StandardScaler()
    clf_o.fit(sco.fit_transform(Fo_tr[sub]), y_tr[sub])
    acc_cpu = float((clf_o.predict(sco.transform(Fo_te)) == y_te).mean())
    assert abs(acc_gpu - acc_cpu) <= 0.005 and acc_gpu > 0.9


@pytest.mark.parametrize("front", ["gammatone", "mel"])
def test_config3_multiplier_sweep_matches_the_oracle(env, front):
    """Dynamics regimes from sub-critical to saturated (extract_lsm_features.py:121-151): the GPU raster is the oracle's at
    every multiplier, including the saturated one where hundreds of neurons fire per step."""
    from lsm_speech_classifier_b200 import filterbank as fb, synth
    from lsm_speech_classifier_b200.extract_lsm_features import build_lsm
    from lsm_speech_classifier_b200.frontend import Frontend
    from oracle import coracle
    pcm, _ = synth.synth_dataset(12, 4)
    fe = Frontend(128, front)
    spk = fe.encode(pcm)
    seen = []
    for mult in (0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0):
        lsm = build_lsm(spk, mult, verbose=False)
        fg, rg = lsm.simulate_batch(spk[:12], nan_to_num=False, return_raster=True)
        fo, ro = coracle.reservoir_run(lsm.reservoir, spk[:12], 0xFF, False, True)
        assert np.array_equal(rg, ro), (front, mult)
        assert np.array_equal(fg, fo, equal_nan=True), (front, mult)
        part, _, _ = lsm.diagnostics(spk[:12])
        np.testing.assert_allclose(part, [(r.sum(0) > 0).mean() * 100 for r in ro])
        seen.append((mult, float(part.mean()), float(rg.sum() / (12 * 400))))
        lsm.close()
    print(f"\n{front}: (multiplier, participation %, mean spikes per step) = " + ", ".join(f"({m:.1f}, {p:.1f}, {s:.1f})" for m, p, s in seen))
    assert seen[-1][1] >= seen[0][1]


def test_config4_n4000_raster_identical(env):
    from lsm_speech_classifier_b200 import synth
    from lsm_speech_classifier_b200.extract_lsm_features import build_lsm
    from lsm_speech_classifier_b200.frontend import Frontend
    from oracle import coracle
    pcm, _ = synth.synth_dataset(4, 3)
    X = Frontend(128, "gammatone").encode(pcm)
    lsm = build_lsm(X, 0.6, num_neurons=4000, verbose=False)
    fg, rg = lsm.simulate_batch(X[:8], nan_to_num=False, return_raster=True)
    fo, ro = coracle.reservoir_run(lsm.reservoir, X[:8], 0xFF, False, True)
    assert rg.shape == (8, 400, 4000) and ro.sum() > 0
    assert np.array_equal(rg, ro) and np.array_equal(fg, fo, equal_nan=True)
