"""BASELINE.json config 1: the reference's default run at CPU-runnable size - 4 classes x 100 synthetic clips, 128-channel
gammatone, feature set original, multiplier 0.6, N = 1000 - GPU against the CPU oracle (modes A and B of BASELINE.md section 3).
Markdown on stdout.    python tools/run_config1.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

pcm, labels = synth.synth_dataset(4, 100, workers=os.cpu_count() or 1)

import torch  # noqa: E402
from lsm_speech_classifier_b200 import _lib  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402
from oracle import coracle  # noqa: E402

keys = FEATURE_SETS["original"]
mask = _lib.feature_mask(keys)
fe = Frontend(128, "gammatone")
spikes = fe.encode(pcm)
lsm = build_lsm(spikes, 0.6, verbose=False)
path = AudioToFeatures(fe, lsm)
h_in = torch.from_numpy(pcm).pin_memory()
h_out = torch.empty((len(pcm), 2000), dtype=torch.float64).pin_memory()
for _ in range(3):
    path.run_host(h_in.numpy(), keys, out=h_out.numpy())
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    path.run_host(h_in.numpy(), keys, out=h_out.numpy())
t_gpu = (time.perf_counter() - t0) / reps
feats, raster = lsm.simulate_batch(spikes, keys, return_raster=True)

tables = (fe.table, fe.params.nwin, fe.params.hop, fe.time_bins, fe.zoom_i0, fe.zoom_f)
cores = coracle.num_threads()
t0 = time.perf_counter()
o_spk = coracle.gammatone_encode(pcm, *tables, [0.70, 0.80, 0.90, 0.95], 0.1, nthreads=0)
o_feat, o_ras = coracle.reservoir_run(lsm.reservoir, o_spk, mask, True, True, nthreads=0)
t_b = time.perf_counter() - t0
t0 = time.perf_counter()
s1 = coracle.gammatone_encode(pcm[:32], *tables, [0.70, 0.80, 0.90, 0.95], 0.1, nthreads=1)
coracle.reservoir_run(lsm.reservoir, s1, mask, True, False, nthreads=1)
t_a = (time.perf_counter() - t0) / 32

S = len(pcm)
print(f"## Config 1 - {S} synthetic clips (4 classes x 100), 128-ch gammatone, original, multiplier 0.6, N = 1000\n")
print(f"* GPU, pinned host PCM in -> feature rows in pinned host memory (one synchronous call): **{t_gpu * 1e3:.2f} ms** = "
      f"{S / t_gpu:,.0f} utterances/s = {S / t_gpu * 400000 / 1e9:.1f} G neuron-steps/s  (a batch of 400 is less than half a resident wave of 888 CTAs)")
print(f"* CPU oracle, mode A (1 thread, how the reference runs): {1 / t_a:.1f} utterances/s -> GPU {S / t_gpu * t_a:,.0f}x")
print(f"* CPU oracle, mode B ({cores} threads): {S / t_b:.0f} utterances/s -> GPU {S / t_gpu / (S / t_b):,.0f}x")
print(f"* spike trains bit-identical: **{bool(np.array_equal(spikes, o_spk))}**; rasters uint8[{S},400,1000] bit-identical: "
      f"**{bool(np.array_equal(raster, o_ras))}**; features bit-identical: **{bool(np.array_equal(feats, o_feat) and np.array_equal(h_out.numpy(), o_feat))}**")
