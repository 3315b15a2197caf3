"""GPU: the CUDA path, through the C ABI, against the CPU oracle on the same inputs.
Bar: spike trains, rasters and features bit-exact (integer/byte outputs and fp64 in the oracle's
rounding order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

THR = [0.70, 0.80, 0.90, 0.95]
GAP = 0.1


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lsm_speech_classifier_b200 import _lib
    return _lib.context(0)


@pytest.fixture(scope="module")
def small_set():
    from lsm_speech_classifier_b200 import synth
    pcm, labels = synth.synth_dataset(4, 6)
    extra = np.stack([np.zeros(16000, np.float32),                                   # silent -> all-zero train
                      np.concatenate([pcm[5][3000:9000], np.zeros(10000, np.float32)]),  # zero padded: denormal tails
                      np.full(16000, 0.25, np.float32),                              # DC
                      (np.random.default_rng(5).standard_normal(16000) * 1e-3).astype(np.float32)])
    return np.concatenate([pcm, extra]), labels


def oracle_spikes(pcm, fe, want_spec=False):
    from oracle import coracle
    return coracle.gammatone_encode(pcm, fe.table, fe.params.nwin, fe.params.hop, fe.time_bins, fe.zoom_i0, fe.zoom_f,
                                    THR, GAP, redundancy=fe.redundancy, want_spec=want_spec)


def test_gammatone_spikes_and_spectrogram_bit_exact(env, small_set):
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = small_set
    fe = Frontend(128, "gammatone")
    want, want_spec = oracle_spikes(pcm, fe, want_spec=True)
    got, spec = fe.encode(torch.from_numpy(pcm).cuda(), return_spectrogram=True)
    torch.cuda.synchronize()
    assert np.array_equal(spec.cpu().numpy(), want_spec)          # fp64, every bit
    assert np.array_equal(got.cpu().numpy(), want)
    assert got[24].sum() == 0                                     # silent clip
    # host-buffer entry point gives the same bytes
    assert np.array_equal(fe.encode(pcm), want)
    # ragged / tiny / empty batches
    assert np.array_equal(fe.encode(pcm[:1]), want[:1])
    assert fe.encode(pcm[:0]).shape == (0, 128, 400)


def test_gammatone_golden_vectors(env, golden):
    from lsm_speech_classifier_b200.frontend import Frontend
    g = golden("frontend_gammatone.npz")
    fe = Frontend(128, "gammatone")
    # filter with the golden design table so the comparison does not depend on this host's libm
    import ctypes as C
    from lsm_speech_classifier_b200 import _lib
    fe.close()
    fe.table = np.ascontiguousarray(g["coefs"])
    h = C.c_void_p()
    fe.ctx.check(fe.ctx.lib.lsm_frontend_create(fe.ctx.h, C.byref(fe.params), _lib._np_ptr(fe.table),
                                                _lib._np_ptr(fe.zoom_i0), _lib._np_ptr(fe.zoom_f), C.byref(h)))
    fe.h = h
    want = np.unpackbits(g["spikes_packed"], axis=-1)[:, :, :400]
    got, spec = fe.encode(g["pcm"], return_spectrogram=True)
    assert np.array_equal(got, want)
    np.testing.assert_allclose(spec[:2], g["spec_norm"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("n_filters,redundancy", [(64, 1), (128, 2), (256, 1), (40, 3)])
def test_gammatone_other_shapes(env, small_set, n_filters, redundancy):
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = small_set
    fe = Frontend(n_filters, "gammatone", redundancy=redundancy)
    want = oracle_spikes(pcm[:9], fe)
    assert np.array_equal(fe.encode(pcm[:9]), want)


def build_snn(X, mult=0.6, **kw):
    from lsm_speech_classifier_b200.snn import SNN, SimulationParams
    from oracle import pyref
    k = kw.get("small_world_graph_k", 200)
    wc = pyref.w_critico(k, 2.0, 2, list(X))
    return SNN(SimulationParams(mean_weight=wc * mult, input_spike_times=X[0], **kw))


def test_reservoir_raster_and_features_bit_exact(env, small_set):
    from lsm_speech_classifier_b200.frontend import Frontend
    from oracle import coracle
    pcm, _ = small_set
    X = Frontend(128, "gammatone").encode(pcm)
    for kw in (dict(), dict(leak_variance_divisor=4.0),
               dict(num_neurons=256, small_world_graph_k=50, num_output_neurons=100),
               dict(num_neurons=2500, small_world_graph_k=500, num_output_neurons=400)):
        lsm = build_snn(X, **kw)
        want_f, want_r = coracle.reservoir_run(lsm.reservoir, X, 0xFF, False, True)
        got_f, got_r = lsm.simulate_batch(X, nan_to_num=False, return_raster=True)
        assert np.array_equal(got_r, want_r), kw
        assert np.array_equal(got_f, want_f, equal_nan=True), kw
        assert want_r.sum() > 0


def test_reservoir_golden_raster(env, golden):
    from lsm_speech_classifier_b200.snn import SNN, SimulationParams
    g = golden("reservoir.npz")
    X = np.unpackbits(golden("frontend_gammatone.npz")["spikes_packed"], axis=-1)[:, :, :400]
    lsm = SNN(SimulationParams(mean_weight=float(g["n1000_mean_weight"]), input_spike_times=X[0]))
    feats, raster = lsm.simulate_batch(X[list(g["utt_index"])], nan_to_num=False, return_raster=True)
    assert np.array_equal(raster, np.unpackbits(g["n1000_raster_packed"], axis=-1)[:, :, :1000])
    assert np.array_equal(feats, g["n1000_features"], equal_nan=True)


def test_feature_sets_and_nan_to_num(env, small_set):
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS
    from lsm_speech_classifier_b200.frontend import Frontend
    from oracle import coracle
    from lsm_speech_classifier_b200 import _lib
    pcm, _ = small_set
    X = Frontend(128, "gammatone").encode(pcm[:10])
    lsm = build_snn(X)
    for name, keys in FEATURE_SETS.items():
        want, _ = coracle.reservoir_run(lsm.reservoir, X, _lib.feature_mask(keys), True, False)
        got = lsm.simulate_batch(X, keys, nan_to_num=True)
        assert got.shape == (10, len(keys) * 400)
        assert np.array_equal(got, want), name


def test_snnpy_protocol_runs_like_the_reference_loop(env, small_set):
    """extract_lsm_features.py:76-89 verbatim shape: reset / set_input_spike_times / simulate /
    extract_features_from_spikes, one sample at a time, against the batched call."""
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = small_set
    X = Frontend(128, "gammatone").encode(pcm[:4])
    lsm = build_snn(X)
    keys = ['spike_counts', 'spike_variances', 'mean_spike_times', 'mean_isi', 'isi_variances']
    rows = []
    for sample in X:
        lsm.reset()
        lsm.set_input_spike_times(sample)
        lsm.simulate()
        fd = lsm.extract_features_from_spikes()
        rows.append(np.concatenate([np.nan_to_num(fd[k].copy()) for k in keys if k in fd]))
        assert lsm.spike_matrix.shape == (400, lsm.num_neurons)
    assert np.array_equal(np.array(rows), lsm.simulate_batch(X, keys, nan_to_num=True))


def test_pipeline_host_equals_staged_calls(env, small_set):
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    from oracle import coracle
    pcm, _ = small_set
    fe = Frontend(128, "gammatone")
    X = fe.encode(pcm)
    lsm = build_snn(X)
    keys = ['spike_counts', 'spike_variances', 'mean_spike_times', 'mean_isi', 'isi_variances']
    want, _ = coracle.reservoir_run(lsm.reservoir, X, 0b01100111, True, False)
    path = AudioToFeatures(fe, lsm)
    spikes_out = np.empty_like(X)
    got = path.run_host(pcm, keys, spikes_out=spikes_out)
    assert np.array_equal(got, want) and np.array_equal(spikes_out, X)
    dev, dspk = path.run(torch.from_numpy(pcm).cuda(), keys)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), want) and np.array_equal(dspk.cpu().numpy(), X)
    # a batch larger than one pipeline chunk, ragged tail: encode->simulate is per-utterance, so tiling must not matter
    big = np.concatenate([pcm] * 90)[:2477]
    got_big = path.run_host(big, keys)
    assert np.array_equal(got_big, np.concatenate([want] * 90)[:2477])           # pageable buffers: the pinned ring, three pieces
    # the same rows whichever way the host buffers travel: staged three-stream path, pinned zero-copy
    import os
    os.environ["LSM_NO_PAGEABLE_RING"] = "1"
    try:
        assert np.array_equal(path.run_host(big, keys), got_big)
    finally:
        del os.environ["LSM_NO_PAGEABLE_RING"]
    h_big = torch.from_numpy(big).pin_memory()
    assert np.array_equal(path.run_host(h_big.numpy(), keys, out=torch.empty(got_big.shape, dtype=torch.float64).pin_memory().numpy()), got_big)


@pytest.mark.parametrize("n_filters,kw", [(128, {}), (128, dict(leak_variance_divisor=4.0)), (64, {}), (256, {}),
                                          (128, dict(num_neurons=700, small_world_graph_k=140, num_output_neurons=300))])
def test_fused_kernel_equals_two_kernel_path_and_oracle(env, small_set, n_filters, kw, monkeypatch):
    """audio -> features as ONE kernel (spikes handed over in shared memory) vs K1 then K2 vs the oracle."""
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    from oracle import coracle
    pcm, _ = small_set
    fe = Frontend(n_filters, "gammatone")
    X = oracle_spikes(pcm, fe)
    lsm = build_snn(X, **kw)
    want, _ = coracle.reservoir_run(lsm.reservoir, X, 0xFF, True, False)
    keys = list(lsm_keys())
    path = AudioToFeatures(fe, lsm)
    d_pcm = torch.from_numpy(pcm).cuda()
    fused, spk = path.run(d_pcm, keys)
    fused_nospk, none = path.run(d_pcm, keys, want_spikes=False)
    torch.cuda.synchronize()
    assert path.fused == (kw.get("num_neurons", 1000) == 1000)
    assert (none is None) == path.fused
    monkeypatch.setenv("LSM_NO_FUSE", "1")
    unfused, spk2 = path.run(d_pcm, keys)
    torch.cuda.synchronize()
    monkeypatch.delenv("LSM_NO_FUSE")
    assert np.array_equal(spk.cpu().numpy(), X) and np.array_equal(spk2.cpu().numpy(), X)
    for got in (fused, fused_nospk, unfused):
        assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(path.run_host(pcm, keys), want)
    # pinned host buffers: the fused kernel reads PCM and writes features straight over PCIe (zero-copy path)
    h_pcm = torch.from_numpy(pcm).pin_memory()
    h_out = torch.empty((len(pcm), want.shape[1]), dtype=torch.float64).pin_memory()
    h_spk = np.empty_like(X)
    path.run_host(h_pcm.numpy(), keys, out=h_out.numpy(), spikes_out=h_spk)
    assert np.array_equal(h_out.numpy(), want) and np.array_equal(h_spk, X)


def lsm_keys():
    from lsm_speech_classifier_b200 import _lib
    return _lib.FEATURE_KEYS


def test_odd_number_of_steps_and_one_word_of_channels(env, small_set):
    """T * CW odd (297 steps, 32 channels): the reservoir's spike list sits behind the bit plane and is read with 8-byte loads, so
    the plane is padded to whole 8-byte words (round-1 advisor finding).  Stand-alone reservoir, fused kernel (one warp per CTA)
    and host path against the oracle."""
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import SNN, AudioToFeatures, SimulationParams
    from oracle import coracle, pyref
    pcm, _ = small_set
    thr = [0.5, 0.7, 0.9]
    fe = Frontend(32, "gammatone", thresholds=thr, hysteresis_gap=0.15, time_bins=99)      # hop 162, 97 columns zoomed to 99
    assert fe.steps == 297 and fe.rows == 32
    X = coracle.gammatone_encode(pcm, fe.table, fe.params.nwin, fe.params.hop, fe.time_bins, fe.zoom_i0, fe.zoom_f, thr, 0.15)
    assert np.array_equal(fe.encode(pcm), X) and X.shape == (len(pcm), 32, 297)
    wc = pyref.w_critico(50, 2.0, 2, list(X))
    lsm = SNN(SimulationParams(mean_weight=wc * 0.8, input_spike_times=X[0], num_neurons=250, small_world_graph_k=50,
                               num_output_neurons=100))
    fo, ro = coracle.reservoir_run(lsm.reservoir, X, 0xFF, False, True)
    fg, rg = lsm.simulate_batch(X, nan_to_num=False, return_raster=True)
    assert ro.sum() > 0 and np.array_equal(rg, ro) and np.array_equal(fg, fo, equal_nan=True)
    path = AudioToFeatures(fe, lsm)
    assert path.fused
    keys = list(lsm_keys())
    want = np.nan_to_num(fo)
    got, spk = path.run(torch.from_numpy(pcm).cuda(), keys)
    assert np.array_equal(got.cpu().numpy(), want) and np.array_equal(spk.cpu().numpy(), X)
    assert np.array_equal(path.run_host(pcm, keys), want)


@pytest.mark.parametrize("kw", [{}, dict(leak_variance_divisor=4.0)])
def test_warp_specialised_lane_channel_kernel_equals_the_oracle(env, small_set, kw, monkeypatch):
    """LSM_WS=1: filter + encoder group and reservoir group side by side in one CTA, bit planes handed over through named barriers
    (gammatone_ws_kernel); more utterances than resident groups, device and pinned-host PCM, PCM16, and a widened near-tie test
    so that the exact pass over the work list runs too."""
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    from oracle import coracle
    pcm, _ = small_set
    fe = Frontend(128, "gammatone")
    X = oracle_spikes(pcm, fe)
    lsm = build_snn(X, **kw)
    want, _ = coracle.reservoir_run(lsm.reservoir, X, 0xFF, True, False)
    keys = list(lsm_keys())
    path = AudioToFeatures(fe, lsm)
    big = np.concatenate([pcm] * 40)[:1403]
    want_big = np.concatenate([want] * 40)[:1403]
    monkeypatch.setenv("LSM_WS", "1")
    launches = fe.ctx.launches
    got, spk = path.run(torch.from_numpy(big).cuda(), keys)
    torch.cuda.synchronize()
    assert fe.ctx.launches - launches == 2                       # the warp-specialised kernel + the exact pass over its work list
    assert np.array_equal(got.cpu().numpy(), want_big) and np.array_equal(spk.cpu().numpy()[:len(pcm)], X)
    h_in = torch.from_numpy(big).pin_memory()
    h_out = torch.zeros((len(big), want.shape[1]), dtype=torch.float64).pin_memory()
    path.run_host(h_in.numpy(), keys, out=h_out.numpy())
    assert np.array_equal(h_out.numpy(), want_big)
    i16 = np.clip(np.round(pcm * 32768.0), -32768, 32767).astype(np.int16)
    want16 = path.run_host((i16.astype(np.float32) / 32768.0), keys)
    got16, _ = path.run(torch.from_numpy(i16).cuda(), keys)
    assert np.array_equal(got16.cpu().numpy(), want16)
    fe.set_mode("speculative", delta_db=1e-3)                    # many near-ties: the exact pass has work to do
    fe.reruns(reset=True)
    got2, _ = path.run(torch.from_numpy(big).cuda(), keys)
    assert np.array_equal(got2.cpu().numpy(), want_big)
    assert 0 < fe.reruns() < len(big)


def test_spike_density_matches_w_critico_inputs(env, small_set):
    import ctypes as C
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = small_set
    fe = Frontend(128, "gammatone")
    X = fe.encode(torch.from_numpy(pcm).cuda())
    out = np.zeros(2, np.int64)
    fe.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    fe.ctx.check(fe.ctx.lib.lsm_spike_density(fe.ctx.h, C.c_void_p(X.data_ptr()), X.numel(), C.c_void_p(out.ctypes.data)))
    assert out[0] == int(X.sum().item()) and out[1] == X.numel()


def test_errors_are_loud(env):
    from lsm_speech_classifier_b200 import _lib
    from lsm_speech_classifier_b200.frontend import Frontend
    fe = Frontend(128, "gammatone")
    with pytest.raises(ValueError):
        fe.encode(np.zeros((2, 8000), np.float32))
    with pytest.raises(_lib.LsmError):
        Frontend(512, "gammatone")


def oracle_mel(pcm, fe, want_spec=False):
    from lsm_speech_classifier_b200 import filterbank as fb
    from oracle import coracle
    return coracle.mel_encode(pcm, fe.table, fe.window, fe.tw, fe.tw2, fb.pack_mel_basis(fe.table), fe.params.mel_hop,
                              fe.time_bins, fe.zoom_i0, fe.zoom_f, THR, GAP, redundancy=fe.redundancy, want_spec=want_spec)


@pytest.mark.parametrize("n_filters,redundancy", [(128, 1), (64, 2), (256, 1), (40, 1)])
def test_mel_spikes_and_spectrogram_bit_exact(env, small_set, n_filters, redundancy):
    """Mel front end (config 3): GPU vs the C oracle, which fixes the FFT operation order (bit-exact), and vs the
    scipy.fft restatement on the goldens (float32 tolerance)."""
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = small_set
    fe = Frontend(n_filters, "mel", redundancy=redundancy)
    want, want_spec = oracle_mel(pcm, fe, want_spec=True)
    got, spec = fe.encode(torch.from_numpy(pcm).cuda(), return_spectrogram=True)
    torch.cuda.synchronize()
    assert np.array_equal(spec.cpu().numpy().astype(np.float32), want_spec)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(fe.encode(pcm), want)
    assert got[24].sum() == 0      # silent clip


@pytest.mark.parametrize("n_filters,kw", [(128, {}), (64, {}), (256, {}), (128, dict(leak_variance_divisor=4.0)),
                                          (128, dict(num_neurons=2500, small_world_graph_k=500))])
def test_mel_fused_kernel_equals_two_kernel_path_and_oracle(env, small_set, n_filters, kw, monkeypatch):
    """mel audio -> features as ONE kernel (spike train handed over as bits in shared memory, reservoir + readout in the same CTA)
    vs K1m then K2 vs the oracle; 256 channels and a heterogeneous leak take the generic reservoir layout; 2500 neurons do not fuse."""
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    from oracle import coracle
    pcm, _ = small_set
    fe = Frontend(n_filters, "mel")
    X = oracle_mel(pcm, fe)
    lsm = build_snn(X, **kw)
    want, _ = coracle.reservoir_run(lsm.reservoir, X, 0xFF, True, False)
    keys = list(lsm_keys())
    path = AudioToFeatures(fe, lsm)
    assert path.fused == (kw.get("num_neurons", 1000) == 1000)
    d_pcm = torch.from_numpy(pcm).cuda()
    fused, spk = path.run(d_pcm, keys)
    fused_nospk, none = path.run(d_pcm, keys, want_spikes=False)
    torch.cuda.synchronize()
    assert (none is None) == path.fused
    monkeypatch.setenv("LSM_NO_FUSE", "1")
    unfused, spk2 = path.run(d_pcm, keys)
    torch.cuda.synchronize()
    monkeypatch.delenv("LSM_NO_FUSE")
    assert np.array_equal(spk.cpu().numpy(), X) and np.array_equal(spk2.cpu().numpy(), X)
    for got in (fused, fused_nospk, unfused):
        assert np.array_equal(got.cpu().numpy(), want)
    # host buffers (chunked copy / kernel / copy pipeline), with and without the spike trains, ragged multi-chunk batch
    h_spk = np.empty_like(X)
    assert np.array_equal(path.run_host(pcm, keys, spikes_out=h_spk), want) and np.array_equal(h_spk, X)
    big = np.concatenate([pcm] * 70)[:1931]
    assert np.array_equal(path.run_host(big, keys), np.concatenate([want] * 70)[:1931])


def test_mel_config3_scale_every_frame_position_of_the_power_kernel(env):
    """BASELINE config 3 at a size that exercises the power kernel's whole grid: 1500 utterances x 101 frames are dealt to 148 CTAs
    in contiguous runs, so run boundaries, warp pairs whose partner has no frame left and utterance boundaries inside a run all occur.
    Spike trains and all eight features against the C oracle; device path, pinned ring (pageable numpy arrays) and a second batch
    size that moves every boundary."""
    import torch
    from lsm_speech_classifier_b200 import synth
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    from oracle import coracle
    pcm, _ = synth.synth_dataset(30, 50, workers=8)                       # 1500 utterances
    pcm[7] = 0.0                                                           # a silent one in the middle of a run
    pcm[1499, 4000:] = 0.0
    fe = Frontend(128, "mel")
    X = oracle_mel(pcm, fe)
    assert X.sum() > 100000
    got = fe.encode(torch.from_numpy(pcm).cuda()).cpu().numpy()
    assert np.array_equal(got, X)
    assert np.array_equal(fe.encode(torch.from_numpy(pcm[3:1180]).cuda()).cpu().numpy(), X[3:1180])
    lsm = build_snn(X[:300])
    want, _ = coracle.reservoir_run(lsm.reservoir, X, 0xFF, True, False)
    keys = list(lsm_keys())
    path = AudioToFeatures(fe, lsm)
    assert path.fused
    feats, _ = path.run(torch.from_numpy(pcm).cuda(), keys, want_spikes=False)
    assert np.array_equal(feats.cpu().numpy(), want)
    assert np.array_equal(path.run_host(pcm, keys), want)                 # pageable: two pieces of 768 through the pinned ring
    lsm.close()


def test_block_per_frame_kernels_stay_selectable(env):
    """LSM_MEL_BLOCK=1 selects round 1's block-per-frame kernels (read once per process, hence the child process): the same
    parity tests pass with them."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-x", "-q", "-m", "gpu", "-k",
                        "mel_spikes_and_spectrogram_bit_exact or mel_fused_kernel_equals"],
                       env=dict(os.environ, LSM_MEL_BLOCK="1"), cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_diagnostic_switches_give_the_same_bytes(env):
    """The library's diagnostic switches (read once per process, hence the child process): exact filter from the
    start, generic reservoir layout, no dead-time skipping, staged copies instead of zero-copy, two host copy threads.  The parity
    tests of the front end, the reservoir, the host pipeline and the fused kernel pass unchanged."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env2 = dict(os.environ, LSM_EXACT_FILTER="1", LSM_NO_LEAN="1", LSM_NO_DEAD_TIME_SKIP="1", LSM_NO_ZEROCOPY="1", LSM_COPY_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-x", "-q", "-m", "gpu", "-k",
                        "gammatone_spikes_and_spectrogram_bit_exact or reservoir_raster_and_features_bit_exact or "
                        "pipeline_host_equals_staged_calls or fused_kernel_equals_two_kernel_path or async_host_calls_on_two_lanes"],
                       env=env2, cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_mel_golden_vectors(env, golden):
    from lsm_speech_classifier_b200.frontend import Frontend
    g = golden("frontend_mel64.npz")
    pcm = golden("frontend_gammatone.npz")["pcm"][:3]
    fe = Frontend(64, "mel")
    got, spec = fe.encode(pcm, return_spectrogram=True)
    want = np.unpackbits(g["spikes_packed"], axis=-1)[:, :, :400]
    # the goldens come from scipy.fft (pocketfft order) + np.log10; ours is a fixed order: >= 99.9 % of bytes must agree
    assert (got == want).mean() >= 0.999
    np.testing.assert_allclose(spec[:1].astype(np.float32), g["spec_norm"], rtol=0, atol=2e-6)


def test_encoder_entry_point_matches_reference_kats(env, golden):
    """lsm_hysteresis_encode (the reference's convert_spectrogram_to_spikes_hysteresis on the GPU) on the KATs minted from
    the reference's own function, float64 and float32 spectrograms, other threshold sets."""
    from lsm_speech_classifier_b200.create_dataset import convert_spectrogram_to_spikes_hysteresis, create_pure_redundancy
    from oracle import pyref
    g = golden("encoder_kats.npz")
    assert np.array_equal(convert_spectrogram_to_spikes_hysteresis(g["spec64"], THR, GAP), g["spikes64"])
    assert np.array_equal(convert_spectrogram_to_spikes_hysteresis(g["spec32"], THR, GAP), g["spikes32"])
    assert np.array_equal(create_pure_redundancy(g["spikes64"][:5], 3), g["redundancy3"])
    for thr, gap in (([0.5], 0.05), ([0.2, 0.4, 0.6, 0.8, 0.9, 0.95], 0.15), ([0.9, 0.3], 0.05)):
        assert np.array_equal(convert_spectrogram_to_spikes_hysteresis(g["spec64"], thr, gap), pyref.hysteresis_encode(g["spec64"], thr, gap))


def test_audio_to_spectrogram_wrapper(env, small_set):
    from lsm_speech_classifier_b200.create_dataset import audio_to_spectrogram
    pcm, _ = small_set
    s = audio_to_spectrogram(pcm[0], 128, "gammatone")
    assert s.shape == (128, 100) and s.dtype == np.float64 and 0.0 <= s.min() and s.max() < 1.0
    m = audio_to_spectrogram(pcm[0], 64, "mel")
    assert m.shape == (64, 100) and m.dtype == np.float32
    z = audio_to_spectrogram(np.zeros(16000, np.float32), 128, "gammatone")
    assert z.dtype == np.float32 and not z.any()                       # create_dataset.py:64-65


def test_other_threshold_counts_and_empty_batches(env, small_set):
    from lsm_speech_classifier_b200.frontend import Frontend
    from oracle import coracle
    pcm, _ = small_set
    fe = Frontend(128, "gammatone", thresholds=[0.6, 0.8, 0.9], hysteresis_gap=0.05)
    want = coracle.gammatone_encode(pcm[:6], fe.table, fe.params.nwin, fe.params.hop, fe.time_bins, fe.zoom_i0, fe.zoom_f,
                                    [0.6, 0.8, 0.9], 0.05)
    got = fe.encode(pcm[:6])
    assert got.shape == (6, 128, 300) and np.array_equal(got, want)
    X = Frontend(128, "gammatone").encode(pcm[:2])
    lsm = build_snn(X)
    assert lsm.simulate_batch(X[:0]).shape == (0, 8 * 400)
    with pytest.raises(ValueError):
        lsm.simulate_batch(np.zeros((1, 64, 400), np.uint8))


def test_device_diagnostics_match_the_raster(env, small_set):
    """run_network_diagnostics' numbers (extract_lsm_features.py:117-131) reduced on the device vs the raster."""
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = small_set
    X = Frontend(128, "gammatone").encode(pcm[:10])
    for kw in (dict(), dict(num_neurons=2500, small_world_graph_k=500)):
        lsm = build_snn(X, **kw)
        part, dead, avg = lsm.diagnostics(X)
        _, raster = lsm.simulate_batch(X, ['spike_counts'], return_raster=True)
        per_neuron = raster.sum(axis=1, dtype=np.int64)                      # [B, N]
        active = (per_neuron > 0).sum(axis=1)
        assert np.array_equal(dead, lsm.num_neurons - active)
        np.testing.assert_allclose(part, active / lsm.num_neurons * 100)
        np.testing.assert_allclose(avg, per_neuron.mean(axis=1))


def test_back_to_back_calls_on_different_streams_do_not_race(env, small_set):
    """A device-pointer call (torch's stream, asynchronous) followed at once by a host-buffer call (the ctx's own stream)
    share the front end's scratch planes: the library must order them.  No synchronize() in between on purpose."""
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    from oracle import coracle
    pcm, _ = small_set
    big = np.concatenate([pcm] * 30)                 # long enough that the first kernel is still running
    fe = Frontend(128, "gammatone")
    X = oracle_spikes(pcm, fe)
    lsm = build_snn(X)
    keys = list(lsm_keys())
    want, _ = coracle.reservoir_run(lsm.reservoir, X, 0xFF, True, False)
    want_big = np.concatenate([want] * 30)
    path = AudioToFeatures(fe, lsm)
    d_big = torch.from_numpy(big).cuda()
    side = torch.cuda.Stream()
    for _ in range(3):
        out, _ = path.run(d_big, keys, want_spikes=False)            # async, torch's current stream
        host = path.run_host(pcm, keys)                              # synchronous, ctx's own stream
        with torch.cuda.stream(side):
            out2 = fe.encode(d_big)                                  # a third stream
        spikes_host = fe.encode(pcm)
        torch.cuda.synchronize()
        assert np.array_equal(host, want) and np.array_equal(spikes_host, X)
        assert np.array_equal(out.cpu().numpy(), want_big)
        assert np.array_equal(out2.cpu().numpy(), np.concatenate([X] * 30))


# ---------------------------------------------------------------- speculative filter (LSM_FILTER_SPECULATIVE)
def test_speculative_filter_gives_the_exact_spike_trains(env, small_set):
    """The default mode filters with 13 FMAs per sample and re-filters exactly every utterance in which the derived distance
    bound could change a comparison: same bytes as the oracle whatever the extra margin - none, a mix, everything re-executed -
    and, on these ordinary clips, even with the bound switched off (the plain speculative plane)."""
    from lsm_speech_classifier_b200 import synth
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = small_set
    more, _ = synth.synth_dataset(12, 8)
    pcm = np.concatenate([pcm, more])
    fe = Frontend(128, "gammatone")
    want = oracle_spikes(pcm, fe)
    fe.set_mode("exact")
    assert np.array_equal(fe.encode(pcm), want)
    assert fe.reruns() == 0
    fe.set_mode("speculative")                       # the derived bound alone
    assert np.array_equal(fe.encode(pcm), want)
    n_default = fe.reruns(reset=True)
    assert n_default <= 3, n_default                 # ~1e-3 expected per utterance
    fe.set_bound_scale(0.0)                          # no utterance can be flagged: the plain speculative plane
    assert np.array_equal(fe.encode(pcm), want)
    assert fe.reruns(reset=True) == 0
    fe.set_bound_scale(1.0)
    fe.set_mode("speculative", 1e9)                  # every non-degenerate utterance flagged -> exact re-execution
    assert np.array_equal(fe.encode(pcm), want)
    assert fe.reruns(reset=True) >= len(pcm) - 1
    fe.set_mode("speculative", 1e-3)                 # a mix of both inside one launch
    assert np.array_equal(fe.encode(pcm), want)
    n_mix = fe.reruns(reset=True)
    assert 0 < n_mix < len(pcm), n_mix
    with pytest.raises(Exception):
        fe.set_mode("speculative", -1.0)


def _audit_report(tag, a):
    print(f"\n{tag}: {len(a)} clips; largest |dB_spec - dB_exact| {a[:, 0].max():.3e} dB (median {np.median(a[:, 0]):.3e}); "
          f"largest distance/bound: cells {a[:, 1].max():.3e}, maximum {a[:, 2].max():.3e}, minimum {a[:, 3].max():.3e}; "
          f"bound on the minimum: median {np.median(a[:, 5]):.3e} dB, max {a[:, 5].max():.3e} dB")


def test_speculative_distance_is_inside_the_derived_bound(env, small_set):
    """The guarantee behind the speculative mode, measured: both arrangements run on the same clips (lsm_frontend_audit) and
    every window cell, the plane's maximum and its floored minimum of the speculative pass lie within the derived bound of
    the exact pass's.  Speech-like clips here; ten thousand adversarial ones below."""
    from lsm_speech_classifier_b200 import synth
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = small_set
    more, _ = synth.synth_dataset(12, 40, start_utt=300)
    tone = (np.sin(2 * np.pi * 3000 * np.arange(16000) / 16000) * 0.5).astype(np.float32)[None]
    pcm = np.concatenate([pcm, more, tone])
    fe = Frontend(128, "gammatone")
    a = fe.audit(pcm)
    _audit_report("speech-like", a)
    assert a[:, 1].max() <= 1.0 and a[:, 2].max() <= 1.0 and a[:, 3].max() <= 1.0
    assert a[:, 0].max() < 1e-6
    silent = a[24]                                    # all-zero clip: both planes are exactly -180 dB
    assert silent[0] == 0.0 and silent[6] == 0.0


def test_ten_thousand_adversarial_clips_stay_inside_the_bound(env):
    """>= 10 000 clips built to separate the two arrangements (out-of-band tone + in-band component at
    -60..-79 dB, full-scale clipping, PCM16 extremes, 80 dB chirps, impulses, resonances, l1-attaining sign patterns):
    (a) no cell, maximum or minimum leaves the derived bound (largest ratio reported), (b) the default mode's spike trains
    equal the exact mode's on all of them, (c) the re-execution rate the bound costs is reported."""
    import adversarial
    from lsm_speech_classifier_b200.frontend import Frontend
    fe = Frontend(128, "gammatone")
    worst = np.zeros(4)
    reruns = total = 0
    audits = []
    for start in range(0, 10080, 1008):
        pcm = adversarial.clips(start, 1008, seed=11, coefs=fe.table)
        a = fe.audit(pcm)
        audits.append(a)
        assert a[:, 1].max() <= 1.0 and a[:, 2].max() <= 1.0 and a[:, 3].max() <= 1.0, (start, a[:, 1:4].max(axis=0))
        fe.set_mode("exact")
        want = fe.encode(pcm)
        fe.set_mode("speculative")
        fe.reruns(reset=True)
        got = fe.encode(pcm)
        reruns += fe.reruns(reset=True)
        total += len(pcm)
        assert np.array_equal(got, want), start
    a = np.concatenate(audits)
    _audit_report("adversarial", a)
    print(f"exact re-executions under the derived bound: {reruns} of {total} adversarial clips ({100.0 * reruns / total:.2f} %)")
    assert total >= 10000
    # the fixed 1e-7 dB margin of round 1 would not have covered these clips: the planes differ by more than a third of it
    assert a[:, 0].max() > 3e-8


@pytest.mark.parametrize("pipeline", [False, True])
def test_fused_pipeline_same_features_in_both_filter_modes(env, small_set, monkeypatch, pipeline):
    if pipeline:
        monkeypatch.setenv("LSM_PIPELINE", "1")          # the warp-specialised kernel instead of the lane = channel fused one
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import SNN, AudioToFeatures, SimulationParams
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, calculate_theoretical_w_critico
    pcm, _ = small_set
    fe = Frontend(128, "gammatone")
    spikes = fe.encode(pcm)
    params = SimulationParams(input_spike_times=spikes[0])
    params.mean_weight = calculate_theoretical_w_critico(params, spikes, verbose=False) * 0.6
    lsm = SNN(simulation_params=params)
    pipe = AudioToFeatures(fe, lsm)
    assert pipe.fused
    keys = FEATURE_SETS["original"]
    fe.set_mode("exact")
    a = pipe.run_host(pcm, keys)
    fe.set_mode("speculative")
    b = pipe.run_host(pcm, keys)
    fe.set_mode("speculative", 1e9)
    c = pipe.run_host(pcm, keys)
    fe.set_mode("speculative", 1e-3)
    d = pipe.run_host(pcm, keys)
    assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d)


def test_speculative_filter_arrangements_agree(env, small_set, monkeypatch):
    """Two arrangements of the speculative filter in the stand-alone front end - lane = utterance (energy kernel + encoder
    kernel, the default there) and lane = channel (inside the encoder kernel, LSM_NO_LANES=1 or shapes the energy kernel does
    not take) - give the oracle's spike trains for ragged batch sizes (partial 32-utterance groups included)."""
    from lsm_speech_classifier_b200.frontend import Frontend
    pcm, _ = small_set
    fe = Frontend(128, "gammatone")
    want = oracle_spikes(pcm, fe)
    launches = fe.ctx.launches
    for n in (1, 2, 27, len(pcm)):
        assert np.array_equal(fe.encode(pcm[:n]), want[:n]), n
    assert fe.ctx.launches - launches == 8                    # two kernels per call: the lanes arrangement ran
    fe40 = Frontend(40, "gammatone")
    want40 = oracle_spikes(pcm[:9], fe40)
    assert np.array_equal(fe40.encode(pcm[:9]), want40)
    monkeypatch.setenv("LSM_NO_LANES", "1")
    launches = fe.ctx.launches
    for n in (1, 27, len(pcm)):
        assert np.array_equal(fe.encode(pcm[:n]), want[:n]), n
    assert fe.ctx.launches - launches == 3                    # one kernel per call
    assert np.array_equal(fe40.encode(pcm[:9]), want40)


@pytest.mark.parametrize("variant,n_launches", [("1", 6), ("2", 4)])
def test_warp_specialised_kernel_equals_the_lane_channel_kernel(env, small_set, monkeypatch, variant, n_launches):
    """Default shape, LSM_PIPELINE=1 (TMA-fed filter warps) / 2 (energy-unit filter warps, no peak pre-pass): the warp-specialised kernel (lane = utterance filter warps + encoder/reservoir units in
    one CTA, flagged utterances finished by the exact kernel) against the lane = channel fused kernel in exact mode:
    same features and spike trains for ragged batch sizes, whatever fraction of the batch takes the exact pass, float32 and
    PCM16, device and host buffers."""
    import torch
    from lsm_speech_classifier_b200 import synth
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import SNN, AudioToFeatures, SimulationParams
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, calculate_theoretical_w_critico
    pcm, _ = small_set
    more, _ = synth.synth_dataset(12, 30, start_utt=40)
    pcm = np.concatenate([pcm, more])                      # 388 utterances
    fe = Frontend(128, "gammatone")
    fe.set_mode("exact")
    spikes = fe.encode(pcm)
    params = SimulationParams(input_spike_times=spikes[0])
    params.mean_weight = calculate_theoretical_w_critico(params, spikes, verbose=False) * 0.6
    lsm = SNN(simulation_params=params)
    pipe = AudioToFeatures(fe, lsm)
    keys = FEATURE_SETS["original"]
    want = pipe.run_host(pcm, keys)
    monkeypatch.setenv("LSM_PIPELINE", variant)
    d_pcm = torch.from_numpy(pcm).cuda()
    for delta, lo, hi in ((0.0, 0, 3), (1e-3, 1, len(pcm) - 1), (1e9, len(pcm) - 1, len(pcm))):
        fe.set_mode("speculative", delta)
        fe.reruns(reset=True)
        launches = fe.ctx.launches
        spk = np.zeros_like(spikes)
        got = pipe.run_host(pcm, keys, spikes_out=spk)
        assert fe.ctx.launches - launches == n_launches, "two pieces on the two lanes: (peak pre-pass +) pipeline kernel + exact pass each"
        assert np.array_equal(got, want), delta
        assert np.array_equal(spk, spikes), delta
        assert lo <= fe.reruns() <= hi, (delta, fe.reruns())
    fe.set_mode("speculative")
    for n in (1, 31, 33, 64, 100, len(pcm)):               # partial groups, one piece / two pieces
        out, dspk = pipe.run(d_pcm[:n], keys)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want[:n]), n
        assert np.array_equal(dspk.cpu().numpy(), spikes[:n]), n
    out, none = pipe.run(d_pcm, keys, want_spikes=False)
    assert none is None and np.array_equal(out.cpu().numpy(), want)
    # PCM16 through the same kernel
    i16 = np.clip(np.round(pcm * 32768.0), -32768, 32767).astype(np.int16)
    as_f32 = i16.astype(np.float32) / np.float32(32768.0)
    monkeypatch.delenv("LSM_PIPELINE")
    want16 = pipe.run_host(as_f32, keys)
    monkeypatch.setenv("LSM_PIPELINE", variant)
    out16, _ = pipe.run(torch.from_numpy(i16).cuda(), keys)
    assert np.array_equal(out16.cpu().numpy(), want16)
    h_in = torch.from_numpy(i16).pin_memory()
    h_out = torch.zeros((len(i16), want.shape[1]), dtype=torch.float64).pin_memory()
    pipe.run_host_async(h_in, keys, out=h_out, lane=1)
    fe.ctx.sync_all()
    assert np.array_equal(h_out.numpy(), want16)


def test_async_host_calls_on_two_lanes(env, small_set):
    """lsm_pipeline_run_host_async: consecutive pinned batches alternate the two launch lanes (two scratch slots, two launches
    in flight) and give the synchronous call's feature rows; three launches on the same lane pair reuse the slots correctly."""
    import torch
    from lsm_speech_classifier_b200 import synth, _lib
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import SNN, AudioToFeatures, SimulationParams
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, calculate_theoretical_w_critico
    pcm, _ = small_set
    more, _ = synth.synth_dataset(12, 70, start_utt=900)           # 840 utterances: more than one resident wave
    batches = [np.concatenate([pcm, more[:400]]), more[400:], more[:333][::-1].copy(), pcm[:5]]
    fe = Frontend(128, "gammatone")
    spikes = fe.encode(pcm)
    params = SimulationParams(input_spike_times=spikes[0])
    params.mean_weight = calculate_theoretical_w_critico(params, spikes, verbose=False) * 0.6
    lsm = SNN(simulation_params=params)
    pipe = AudioToFeatures(fe, lsm)
    keys = FEATURE_SETS["original"]
    want = [pipe.run_host(b, keys) for b in batches]
    pinned_in = [torch.from_numpy(b).pin_memory() for b in batches]
    pinned_out = [torch.empty((len(b), want[0].shape[1]), dtype=torch.float64).pin_memory() for b in batches]
    for rep in range(2):
        for o in pinned_out:
            o.zero_()
        for i, (x, o) in enumerate(zip(pinned_in, pinned_out)):
            pipe.run_host_async(x, keys, out=o, lane=i & 1)
        fe.ctx.sync_all()
        for i, (o, w) in enumerate(zip(pinned_out, want)):
            assert np.array_equal(o.numpy(), w), (rep, i)
    with pytest.raises(_lib.LsmError):
        pipe.run_host_async(batches[3], keys, out=np.empty((5, want[0].shape[1])), lane=0)      # pageable buffers


def test_pcm16_input_gives_the_float32_results(env, small_set):
    """PCM16 ingest (what a 16 kHz WAV holds; create_dataset.py:22-36 via librosa.load -> sample / 32768): the kernel's
    int16 path gives the features and spike trains of the float32 path on the same samples, and both equal the oracle."""
    import torch
    from lsm_speech_classifier_b200 import _lib
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import SNN, AudioToFeatures, SimulationParams
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, calculate_theoretical_w_critico
    pcm, _ = small_set
    i16 = np.clip(np.round(pcm * 32768.0), -32768, 32767).astype(np.int16)
    as_f32 = (i16.astype(np.float32) / np.float32(32768.0))               # what soundfile/librosa hand to the reference
    fe = Frontend(128, "gammatone")
    want_spikes = oracle_spikes(as_f32, fe)
    params = SimulationParams(input_spike_times=want_spikes[0])
    params.mean_weight = calculate_theoretical_w_critico(params, want_spikes, verbose=False) * 0.6
    lsm = SNN(simulation_params=params)
    pipe = AudioToFeatures(fe, lsm)
    keys = FEATURE_SETS["original"]
    want = pipe.run_host(as_f32, keys)
    for mode in ("speculative", "exact"):
        fe.set_mode(mode)
        out, spk = pipe.run(torch.from_numpy(i16).cuda(), keys)
        torch.cuda.synchronize()
        assert np.array_equal(spk.cpu().numpy(), want_spikes), mode
        assert np.array_equal(out.cpu().numpy(), want), mode
    h_in = torch.from_numpy(i16).pin_memory()
    h_out = torch.zeros((len(i16), want.shape[1]), dtype=torch.float64).pin_memory()
    pipe.run_host_async(h_in, keys, out=h_out, lane=1)
    fe.ctx.sync_all()
    assert np.array_equal(h_out.numpy(), want)
    femel = Frontend(128, "mel")
    with pytest.raises(_lib.LsmError):
        AudioToFeatures(femel, lsm).run(torch.from_numpy(i16).cuda(), keys)       # not a fused pair


def test_large_pinned_batch_is_split_across_the_two_lanes(env, monkeypatch):
    """lsm_pipeline_run_host with pinned buffers and at least two resident waves of utterances launches the two halves on
    the two lanes; the feature rows are those of the unsplit call (LSM_NO_SPLIT=1), of the warp-specialised kernel and of
    the oracle."""
    import torch
    from oracle import coracle
    from lsm_speech_classifier_b200 import synth, _lib
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import SNN, AudioToFeatures, SimulationParams
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, calculate_theoretical_w_critico
    base, _ = synth.synth_dataset(12, 20)
    pcm = np.concatenate([base] * 8)[:1900]                      # > 2 x 888 resident CTAs
    pcm[7] = 0.0                                                 # a silent clip in the first half
    fe = Frontend(128, "gammatone")
    spikes = fe.encode(base)
    params = SimulationParams(input_spike_times=spikes[0])
    params.mean_weight = calculate_theoretical_w_critico(params, spikes, verbose=False) * 0.6
    lsm = SNN(simulation_params=params)
    pipe = AudioToFeatures(fe, lsm)
    keys = FEATURE_SETS["original"]
    h_in = torch.from_numpy(pcm).pin_memory()
    h_out = torch.zeros((len(pcm), 2000), dtype=torch.float64).pin_memory()
    launches = fe.ctx.launches
    pipe.run_host(h_in.numpy(), keys, out=h_out.numpy())
    assert fe.ctx.launches - launches == 2
    split = h_out.numpy().copy()
    monkeypatch.setenv("LSM_NO_SPLIT", "1")
    h_out.zero_()
    launches = fe.ctx.launches
    pipe.run_host(h_in.numpy(), keys, out=h_out.numpy())
    assert fe.ctx.launches - launches == 1
    assert np.array_equal(split, h_out.numpy())
    monkeypatch.delenv("LSM_NO_SPLIT")
    monkeypatch.setenv("LSM_PIPELINE", "1")
    h_out.zero_()
    pipe.run_host(h_in.numpy(), keys, out=h_out.numpy())        # warp-specialised kernel, copy-engine staging
    monkeypatch.delenv("LSM_PIPELINE")
    assert np.array_equal(split, h_out.numpy())
    want_spk = oracle_spikes(base, fe)
    want, _ = coracle.reservoir_run(lsm.reservoir, want_spk, _lib.feature_mask(keys), True, False)
    assert np.array_equal(split[240:480], want) and np.array_equal(split[1680:1900], want[:220])
    assert not split[7].any()


def test_front_end_batches_beyond_one_energy_pass(env):
    """The stand-alone front end works in passes of 8192 utterances (energy planes); a batch of 8200 crosses that boundary."""
    import torch
    from lsm_speech_classifier_b200 import synth
    from lsm_speech_classifier_b200.frontend import Frontend
    base, _ = synth.synth_dataset(4, 10)
    reps = 8200 // len(base) + 1
    pcm = np.concatenate([base] * reps)[:8200]
    fe = Frontend(128, "gammatone")
    want = oracle_spikes(base, fe)
    got = fe.encode(torch.from_numpy(pcm).cuda()).cpu().numpy()
    assert got.shape == (8200, 128, 400)
    tiled = np.concatenate([want] * reps)[:8200]
    assert np.array_equal(got[:80], tiled[:80]) and np.array_equal(got[8150:], tiled[8150:])
    assert np.array_equal(got, tiled)


def test_fused_gather_writes_rows_into_every_destination(env, small_set):
    """lsm_reservoir_set_gather: the readout epilogue also stores each feature row at row0 + u of every destination matrix
    (on one GPU here: two local matrices stand in for the ranks' IPC-mapped gather buffers); off again with an empty list."""
    import torch
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import SNN, AudioToFeatures, SimulationParams
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, calculate_theoretical_w_critico
    pcm, _ = small_set
    fe = Frontend(128, "gammatone")
    spikes = fe.encode(pcm)
    params = SimulationParams(input_spike_times=spikes[0])
    params.mean_weight = calculate_theoretical_w_critico(params, spikes, verbose=False) * 0.6
    lsm = SNN(simulation_params=params)
    pipe = AudioToFeatures(fe, lsm)
    keys = FEATURE_SETS["original"]
    d_pcm = torch.from_numpy(pcm).cuda()
    want, _ = pipe.run(d_pcm, keys)
    B, F = want.shape
    dst = [torch.full((3 * B + 5, F), -1.0, dtype=torch.float64, device="cuda") for _ in range(2)]
    lsm.set_gather([d.data_ptr() for d in dst], B + 5)
    got, _ = pipe.run(d_pcm, keys)                         # fused kernel
    got2 = lsm.simulate_batch(torch.from_numpy(spikes).cuda(), keys)      # stand-alone reservoir kernel: same epilogue
    torch.cuda.synchronize()
    lsm.set_gather([], 0)
    assert torch.equal(got, want) and torch.equal(got2, want)
    for d in dst:
        assert torch.equal(d[B + 5:2 * B + 5], want)
        assert bool((d[:B + 5] == -1).all()) and bool((d[2 * B + 5:] == -1).all())
    dst[0].fill_(-1.0)
    pipe.run(d_pcm, keys)
    torch.cuda.synchronize()
    assert bool((dst[0] == -1).all())
