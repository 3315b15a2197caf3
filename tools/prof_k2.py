"""One launch of each reservoir arm for ncu (B = 4096, N = 1000): python tools/prof_k2.py [multiplier]
    ncu --set full --clock-control none --import-source on -k regex:'reservoir_kernel|dense_step_kernel' ..."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

mult = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
pcm, _ = synth.synth_dataset(12, 43, workers=min(16, os.cpu_count() or 1))
import torch  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402

keys = FEATURE_SETS["original"]
X = Frontend(128, "gammatone").encode(pcm)
spk = torch.from_numpy(np.concatenate([X] * 8)[:4096]).cuda()
lsm = build_lsm(X[:500], mult, verbose=False)
for mode in ("event", "dense"):
    lsm.set_mode(mode)
    f = lsm.simulate_batch(spk, keys)
    torch.cuda.synchronize()
    print(mode, float(f.sum()))
