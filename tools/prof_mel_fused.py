"""ncu driver: the fused mel audio -> features kernel on one 2400-utterance batch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

base, _ = synth.synth_dataset(12, 20, workers=os.cpu_count() or 1)
pcm = np.concatenate([base] * 11)[:2400]

import torch  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

d_pcm = torch.from_numpy(pcm).cuda()
fe = Frontend(128, "mel")
lsm = build_lsm(fe.encode(d_pcm[:500]).cpu().numpy(), 0.6, verbose=False)
path = AudioToFeatures(fe, lsm)
keys = FEATURE_SETS["original"]
for _ in range(3):
    out, _ = path.run(d_pcm, keys, want_spikes=False)
torch.cuda.synchronize()
print("ok fused" if path.fused else "ok two-kernel", float(out.sum()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    path.run(d_pcm, keys, want_spikes=False)
b.record(); torch.cuda.synchronize()
print("ms per batch", a.elapsed_time(b) / 5)
