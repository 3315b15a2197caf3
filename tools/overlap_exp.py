"""Experiment: two front ends on two streams, so the fp64-bound energy kernel of batch i+1 can run beside the
issue-bound encoder+reservoir kernel of batch i.  LSM_LANES=1 LSM_ER_PER_SM=k python tools/overlap_exp.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

B = 2400
base, _ = synth.synth_dataset(12, 20, workers=os.cpu_count() or 1)
pcm = np.concatenate([base] * (B // len(base) + 1))[:B]

import torch  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

d_pcm = torch.from_numpy(pcm).cuda()
fes = [Frontend(128, "gammatone") for _ in range(2)]
lsm = build_lsm(fes[0].encode(d_pcm[:500]).cpu().numpy(), 0.6, verbose=False)
paths = [AudioToFeatures(fe, lsm) for fe in fes]
keys = FEATURE_SETS["original"]
outs = [torch.empty((B, 2000), dtype=torch.float64, device="cuda") for _ in range(2)]
prio = os.environ.get("PRIO", "0") == "1"
streams = [torch.cuda.Stream(priority=(-1 if (prio and i == 1) else 0)) for i in range(2)]


def run(steps, two_streams):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        k = i & 1
        if two_streams:
            with torch.cuda.stream(streams[k]):
                if i == 0:
                    streams[k].wait_event(a)
                paths[k].run(d_pcm, keys, out=outs[k], want_spikes=False)
        else:
            paths[k].run(d_pcm, keys, out=outs[k], want_spikes=False)
    if two_streams:
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


run(4, False); run(4, True)
print("one stream  ms/step", run(12, False))
print("two streams ms/step", run(12, True), {k: v for k, v in os.environ.items() if k.startswith("LSM_") or k == "PRIO"})
assert torch.equal(outs[0], outs[1])
