"""ncu driver: the mel front-end kernel on one 2400-utterance batch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2400
base, _ = synth.synth_dataset(12, 20, workers=os.cpu_count() or 1)
pcm = np.concatenate([base] * (B // len(base) + 1))[:B]

import torch  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402

d_pcm = torch.from_numpy(pcm).cuda()
fe = Frontend(int(os.environ.get("LSM_MELS", "128")), "mel")
for _ in range(3):
    s = fe.encode(d_pcm)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    fe.encode(d_pcm)
b.record(); torch.cuda.synchronize()
print("ok mel ms", a.elapsed_time(b) / 5, float(s.float().mean()))
