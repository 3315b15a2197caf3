"""GPU: the dense (tensor-core) arm of the reservoir against the CPU oracle and against the event-driven arm.

BASELINE.json's north star names two ways to form the recurrent current of lsm.simulate()
(/root/reference/extract_lsm_features.py:81): the event-driven gather and a dense spikes[B,N] . W[N,N] tensor-core tile.
Both must give the oracle's rasters and features bit for bit (integer weights: the row sums are exact in any order).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lsm_speech_classifier_b200 import _lib, synth
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = synth.synth_dataset(12, 12)
    X = Frontend(128, "gammatone").encode(pcm)
    return _lib.context(0), X


def _dense_w(r):
    W = np.zeros((r.num_neurons, r.num_neurons), dtype=np.int64)            # [post][pre]
    post = np.repeat(np.arange(r.num_neurons), np.diff(r.w_rowptr))
    np.add.at(W, (post, r.w_col), r.w_q)
    return W


def test_contraction_equals_the_integer_matrix_product(env):
    """tcgen05.mma kind::i8 over three unsigned digit planes, recombined: exactly W . s."""
    from lsm_speech_classifier_b200.extract_lsm_features import build_lsm
    _, X = env
    lsm = build_lsm(X, 0.6, verbose=False)
    r = lsm.reservoir
    rs = np.random.RandomState(3)
    S = (rs.random_sample((200, r.num_neurons)) < 0.1).astype(np.uint8)
    S[0] = 0; S[1] = 1; S[2] = 0; S[2, 7] = 1; S[3] = 0; S[3, r.num_neurons - 1] = 1
    got = lsm.dense_probe(S).cpu().numpy().astype(np.int64)
    assert np.array_equal(got, S.astype(np.int64) @ _dense_w(r).T)
    lsm.close()


def test_contraction_with_negative_and_wide_weights_uses_a_signed_top_plane(env):
    """Weights outside [0, 2^24): four planes, the top one signed (u8 x s8)."""
    from lsm_speech_classifier_b200.reservoir import SimulationParams, build_reservoir
    from lsm_speech_classifier_b200.snn import SNN
    _, X = env
    p = SimulationParams(num_neurons=300, mean_weight=0.02, small_world_graph_k=20, input_spike_times=X[0])
    r = build_reservoir(p)
    rs = np.random.RandomState(5)
    r.w_q = rs.randint(-(1 << 26), 1 << 26, size=len(r.w_q)).astype(np.int32)      # |row sum| <= 22 * 2^26 < 2^31
    lsm = SNN(reservoir=r)
    S = (rs.random_sample((130, 300)) < 0.3).astype(np.uint8)
    got = lsm.dense_probe(S).cpu().numpy().astype(np.int64)
    assert np.array_equal(got, S.astype(np.int64) @ _dense_w(r).T)
    # and the whole simulation with such weights: both arms and the oracle agree
    from oracle import coracle
    fo, ro = coracle.reservoir_run(r, X[:6], 0xFF, False, True)
    lsm.set_mode("dense")
    fd, rd = lsm.simulate_batch(X[:6], nan_to_num=False, return_raster=True)
    assert np.array_equal(rd, ro) and np.array_equal(fd, fo, equal_nan=True)
    lsm.close()


@pytest.mark.parametrize("mult", [0.6, 1.0])
def test_dense_arm_rasters_and_features_equal_the_oracle(env, mult):
    from lsm_speech_classifier_b200.extract_lsm_features import build_lsm
    from oracle import coracle
    _, X = env
    lsm = build_lsm(X, mult, verbose=False)
    fo, ro = coracle.reservoir_run(lsm.reservoir, X[:24], 0xFF, False, True)
    lsm.set_mode("dense")
    fd, rd = lsm.simulate_batch(X[:24], nan_to_num=False, return_raster=True)
    assert ro.sum() > 0
    assert np.array_equal(rd, ro), "dense-arm raster differs from the oracle"
    assert np.array_equal(fd, fo, equal_nan=True)
    lsm.set_mode("event")
    fe_, re_ = lsm.simulate_batch(X[:24], nan_to_num=False, return_raster=True)
    assert np.array_equal(re_, rd) and np.array_equal(fe_, fd, equal_nan=True)
    lsm.close()


def test_dense_arm_ragged_batches_device_tensors_and_heterogeneous_leak(env):
    """Batch sizes that are not a multiple of the 128-utterance tile, torch in / torch out, per-neuron leak (non-lean layout)."""
    import torch
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm
    from oracle import coracle
    _, X = env
    keys = FEATURE_SETS["original"]
    from lsm_speech_classifier_b200 import _lib
    lsm = build_lsm(X, 0.8, leak_variance_divisor=5.0, verbose=False)
    want, _ = coracle.reservoir_run(lsm.reservoir, X, _lib.feature_mask(keys), True, False)
    lsm.set_mode("dense")
    for n in (1, 127, 129, len(X)):
        got = lsm.simulate_batch(torch.from_numpy(X[:n]).cuda(), keys)
        assert np.array_equal(got.cpu().numpy(), want[:n]), n
    lsm.close()


def test_dense_arm_refuses_what_it_cannot_serve(env):
    from lsm_speech_classifier_b200 import _lib
    from lsm_speech_classifier_b200.reservoir import SimulationParams, build_reservoir
    from lsm_speech_classifier_b200.snn import SNN
    _, X = env
    # more input rows than neurons: some neuron is driven by several rows
    p = SimulationParams(num_neurons=100, mean_weight=0.02, small_world_graph_k=10, input_spike_times=X[0])
    lsm = SNN(reservoir=build_reservoir(p))
    with pytest.raises(_lib.LsmError):
        lsm.set_mode("dense")
    lsm.close()
