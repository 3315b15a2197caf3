"""GPU: the device StandardScaler (lsm_standardize_fit / _transform) against scikit-learn's, bit for bit
(/root/reference/extract_lsm_features.py:199-201)."""
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


def _feature_like(n, F, seed):
    """Columns shaped like LSM features: small integer counts, means of spike times, constant and zero columns."""
    rng = np.random.default_rng(seed)
    X = np.empty((n, F))
    X[:, 0::4] = rng.integers(0, 9, (n, len(range(0, F, 4))))
    X[:, 1::4] = rng.random((n, len(range(1, F, 4)))) * 400.0
    X[:, 2::4] = rng.standard_normal((n, len(range(2, F, 4)))) * 1e3 + 1e6
    X[:, 3::4] = 0.0
    if F > 7:
        X[:, 7] = 3.25          # constant, non-zero
    return X


@pytest.mark.parametrize("n,F", [(320, 2000), (1, 5), (7, 33), (1920, 3200)])
def test_standard_scaler_matches_sklearn_bit_for_bit(env, n, F):
    import torch
    from sklearn.preprocessing import StandardScaler as SkScaler
    from lsm_speech_classifier_b200.readout import StandardScaler
    X = _feature_like(n, F, 11 + n)
    Xt = _feature_like(max(1, n // 4), F, 12 + n)
    sk = SkScaler()
    want_train = sk.fit_transform(X)
    want_test = sk.transform(Xt)
    sc = StandardScaler()
    got_train = sc.fit_transform(X)
    got_test = sc.transform(Xt)
    assert np.array_equal(sc.mean_, sk.mean_)
    assert np.array_equal(sc.var_, sk.var_)
    assert np.array_equal(sc.scale_, sk.scale_)
    assert sc.n_samples_seen_ == n
    assert np.array_equal(got_train, want_train)
    assert np.array_equal(got_test, want_test)
    # device tensors in -> device tensors out, same bits
    d = sc.transform(torch.from_numpy(Xt).cuda())
    assert d.is_cuda and np.array_equal(d.cpu().numpy(), want_test)


def test_standard_scaler_errors(env):
    from lsm_speech_classifier_b200 import _lib
    from lsm_speech_classifier_b200.readout import StandardScaler
    sc = StandardScaler()
    with pytest.raises(_lib.LsmError):
        sc.transform(np.zeros((2, 3)))
    sc.fit(np.arange(12.0).reshape(4, 3))
    with pytest.raises(ValueError):
        sc.transform(np.zeros((2, 4)))
    assert sc.transform(np.zeros((0, 3))).shape == (0, 3)


# ---------------------------------------------------------------- logistic regression readout (train_classifier.py:36-47)
def _blobs(n, F, K, seed, sep=1.2):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((K, F)) * sep / np.sqrt(F) * 6.0
    y = rng.integers(0, K, n).astype(np.int32)
    X = centers[y] + rng.standard_normal((n, F))
    X[:, ::7] = np.round(X[:, ::7] * 3)          # count-like columns
    return X, y


@pytest.mark.parametrize("n,F,K", [(1920, 400, 12), (320, 2000, 4), (600, 96, 3), (2100, 300, 35), (900, 64, 17)])
def test_logistic_regression_matches_sklearn(env, n, F, K):
    import time
    from sklearn.linear_model import LogisticRegression as SkLR
    from sklearn.preprocessing import StandardScaler as SkScaler
    from lsm_speech_classifier_b200.readout import LogisticRegression
    X, y = _blobs(n + n // 4, F, K, 5 + K)
    sc = SkScaler().fit(X[:n])
    Xtr, Xte, ytr, yte = sc.transform(X[:n]), sc.transform(X[n:]), y[:n], y[n:]
    t0 = time.perf_counter()
    sk = SkLR(random_state=42, max_iter=1000).fit(Xtr, ytr)
    t_sk = time.perf_counter() - t0
    t0 = time.perf_counter()
    clf = LogisticRegression(random_state=42, max_iter=1000).fit(Xtr, ytr)
    t_gpu = time.perf_counter() - t0
    print(f"\nlogreg n={n} F={F} K={K}: sklearn {t_sk * 1e3:.0f} ms ({sk.n_iter_[0]} it), device {t_gpu * 1e3:.0f} ms ({clf.n_iter_[0]} it)")
    assert np.array_equal(clf.classes_, sk.classes_)
    p_sk, p = sk.predict(Xte), clf.predict(Xte)
    acc_sk, acc = np.mean(p_sk == yte), np.mean(p == yte)
    assert abs(acc - acc_sk) <= 0.005 + 1.0 / len(yte), (acc, acc_sk)          # within 0.5 points (one sample of slack)
    assert np.mean(p == p_sk) >= 0.99
    assert np.mean(clf.predict(Xtr) == sk.predict(Xtr)) >= 0.99
    # same objective, same optimum: the device solution is as good as scikit-learn's to solver tolerance (both stop at
    # max|grad| <= 1e-4; with n < F the valley is flat, so compare objective values rather than coefficients)
    def objective(coef, icpt):
        z = Xtr @ coef.T + icpt
        z -= z.max(1, keepdims=True)
        lse = np.log(np.exp(z).sum(1))
        yi = np.searchsorted(sk.classes_, ytr)
        return np.mean(lse - z[np.arange(len(ytr)), yi]) + 0.5 / len(ytr) * np.sum(coef * coef)
    f_sk, f_dev = objective(sk.coef_, sk.intercept_), objective(clf.coef_, clf.intercept_)
    assert f_dev <= f_sk + 1e-5 * max(1.0, abs(f_sk)), (f_dev, f_sk)
    assert np.abs(clf.coef_ - sk.coef_).max() <= 0.25 * np.abs(sk.coef_).max()
    assert abs(clf.score(Xte, yte) - acc) < 1e-12


def test_logistic_regression_on_lsm_features(env):
    """Config-1-shaped end of the pipeline: features of 4 x 100 synthetic utterances -> split -> scaler -> classifier, device
    readout vs scikit-learn on the same standardised matrices (the reference's train_classifier.py, shimmed for sklearn 1.9)."""
    from sklearn.linear_model import LogisticRegression as SkLR
    from sklearn.model_selection import train_test_split
    from lsm_speech_classifier_b200 import synth
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm
    from lsm_speech_classifier_b200.readout import LogisticRegression, StandardScaler
    pcm, labels = synth.synth_dataset(4, 100)
    fe = Frontend(128, "gammatone")
    spikes = fe.encode(pcm)
    Xtr_s, Xte_s, ytr, yte = train_test_split(spikes, labels, test_size=0.2, random_state=42, stratify=labels)
    lsm = build_lsm(Xtr_s, 0.6, verbose=False)
    keys = FEATURE_SETS["original"]
    ftr, fte = lsm.simulate_batch(Xtr_s, keys), lsm.simulate_batch(Xte_s, keys)
    sc = StandardScaler()
    Xtr, Xte = sc.fit_transform(ftr), sc.transform(fte)
    sk = SkLR(random_state=42, max_iter=1000).fit(Xtr, ytr)
    clf = LogisticRegression(random_state=42, max_iter=1000).fit(Xtr, ytr)
    acc_sk, acc = sk.score(Xte, yte), clf.score(Xte, yte)
    print(f"\nconfig-1 readout: sklearn {acc_sk * 100:.2f} %, device {acc * 100:.2f} % ({clf.n_iter_[0]} iterations)")
    assert abs(acc - acc_sk) <= 0.005 + 1.0 / len(yte)
    assert np.mean(clf.predict(Xte) == sk.predict(Xte)) >= 0.97


def test_logistic_regression_errors(env):
    from lsm_speech_classifier_b200 import _lib
    from lsm_speech_classifier_b200.readout import LogisticRegression
    clf = LogisticRegression()
    with pytest.raises(_lib.LsmError):
        clf.predict(np.zeros((2, 3)))
    with pytest.raises(ValueError):
        clf.fit(np.zeros((4, 3)), np.zeros(4))                       # one class
    with pytest.raises(ValueError):
        clf.fit(np.zeros((4, 3)), np.arange(4) % 2)                  # two classes: scikit-learn's binary objective is a different one
    with pytest.raises(_lib.LsmError):
        clf.fit(np.random.default_rng(0).random((140, 3)), np.arange(140) % 65)    # > 64 classes
