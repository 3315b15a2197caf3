"""Tiny run of every kernel for compute-sanitizer (one tool per gpurun call)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

pcm, _ = synth.synth_dataset(3, 2)
pcm = np.concatenate([pcm, np.zeros((1, 16000), np.float32)])

import torch  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

keys = FEATURE_SETS["all"]
for C, fb, R in ((128, "gammatone", 1), (64, "gammatone", 1), (256, "gammatone", 1), (40, "gammatone", 2), (64, "mel", 1), (128, "mel", 1)):
    fe = Frontend(C, fb, redundancy=R)
    X = fe.encode(pcm)
    Xd, spec = fe.encode(torch.from_numpy(pcm).cuda(), return_spectrogram=True)
    for kw in (dict(), dict(leak_variance_divisor=4.0), dict(num_neurons=300, small_world_graph_k=60, num_output_neurons=100)):
        from lsm_speech_classifier_b200.snn import SNN, SimulationParams
        lsm = SNN(SimulationParams(mean_weight=0.011 * 200 / kw.get("small_world_graph_k", 200), input_spike_times=X[0], **kw))
        f, r = lsm.simulate_batch(X, keys, return_raster=True)
        lsm.diagnostics(X)
        path = AudioToFeatures(fe, lsm)
        out, spk = path.run(torch.from_numpy(pcm).cuda(), keys)
        h = path.run_host(pcm, keys)
        hp = torch.from_numpy(pcm).pin_memory()
        ho = torch.empty((len(pcm), h.shape[1]), dtype=torch.float64).pin_memory()
        path.run_host(hp.numpy(), keys, out=ho.numpy())
        torch.cuda.synchronize()
        for name, got in (("run_host pageable", h), ("run_host pinned", ho.numpy()), ("fused device", out.cpu().numpy())):
            bad = np.argwhere(got != f)
            assert len(bad) == 0, (C, fb, kw, name, len(bad), bad[:4].tolist(), [(float(got[tuple(b)]), float(f[tuple(b)])) for b in bad[:4]])
        lsm.close()
    fe.close()
print("sanitize run ok")
