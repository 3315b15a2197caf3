"""ncu driver: K1 stand-alone and the fused kernel in the speculative filter mode (default), one batch."""
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
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

d_pcm = torch.from_numpy(pcm).cuda()
fe = Frontend(128, "gammatone")
fe.set_mode(os.environ.get("LSM_MODE", "speculative"))
lsm = build_lsm(fe.encode(d_pcm[:500]).cpu().numpy(), 0.6, verbose=False)
path = AudioToFeatures(fe, lsm)
keys = FEATURE_SETS["original"]
out, spk = path.run(d_pcm, keys)
for _ in range(2):
    fe.encode(d_pcm)
    path.run(d_pcm, keys, spikes=spk, out=out)
torch.cuda.synchronize()
if os.environ.get("LSM_TIME"):
    def tm(fn, reps=5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    print("K1 ms", tm(lambda: fe.encode(d_pcm)), "fused ms", tm(lambda: path.run(d_pcm, keys, spikes=spk, out=out)),
          "fused no-spikes ms", tm(lambda: path.run(d_pcm, keys, out=out, want_spikes=False)), "env",
          {k: v for k, v in os.environ.items() if k.startswith("LSM_")})
print("ok fused" if path.fused else "ok two-kernel", float(out.sum()), "reruns", fe.reruns())
