"""Three whole steps of the default path (for ncu): python tools/prof_pipe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

pcm, _ = synth.synth_dataset(12, 200, workers=os.cpu_count() or 1)
import torch  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

keys = FEATURE_SETS["original"]
fe = Frontend(128, "gammatone")
d_pcm = torch.from_numpy(pcm).cuda()
lsm = build_lsm(fe.encode(d_pcm[:500]).cpu().numpy(), 0.6, verbose=False)
path = AudioToFeatures(fe, lsm)
out = torch.empty((len(pcm), 2000), dtype=torch.float64, device="cuda")
for _ in range(int(os.environ.get("LSM_STEPS", "3"))):
    path.run(d_pcm, keys, out=out, want_spikes=False)
torch.cuda.synchronize()
print("ok")
