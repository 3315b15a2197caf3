"""Small driver for ncu: K1 and K2 on one 2400-utterance batch, three launches each.
    python tools/prof_kernels.py [B]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2400
base, _ = synth.synth_dataset(12, 20, workers=os.cpu_count() or 1)
pcm = np.concatenate([base] * (B // len(base) + 1))[:B]

import torch  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402

d_pcm = torch.from_numpy(pcm).cuda()
fe = Frontend(128, "gammatone")
spikes = fe.encode(d_pcm)
lsm = build_lsm(spikes[:500].cpu().numpy(), 0.6, verbose=False)
keys = FEATURE_SETS["original"]
for _ in range(3):
    spikes = fe.encode(d_pcm)
    feats = lsm.simulate_batch(spikes, keys)
torch.cuda.synchronize()
if os.environ.get("LSM_TIME"):
    def tm(fn, reps=5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    print("K1 ms", tm(lambda: fe.encode(d_pcm)), "K2 ms", tm(lambda: lsm.simulate_batch(spikes, keys)), "minb", os.environ.get("LSM_K1_MINB"))
print("ok", spikes.float().mean().item(), feats.shape)
if os.environ.get("LSM_TIME"):
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    path = AudioToFeatures(fe, lsm)
    out, spk = path.run(d_pcm, keys)
    torch.cuda.synchronize()
    print("fused" if path.fused else "two-kernel", "pipeline ms", tm(lambda: path.run(d_pcm, keys, spikes=spk, out=out)),
          "no-spikes ms", tm(lambda: path.run(d_pcm, keys, out=out, want_spikes=False)))
