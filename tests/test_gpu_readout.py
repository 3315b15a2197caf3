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
