"""Timing experiments on the warp-specialised lane = channel kernel (one B200): python tools/ws_exp.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

pcm, _ = synth.synth_dataset(12, 200, workers=min(16, os.cpu_count() or 1))
import torch  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

keys = FEATURE_SETS["original"]
fe = Frontend(128, "gammatone")
d_pcm = torch.from_numpy(pcm).cuda()
lsm = build_lsm(fe.encode(d_pcm[:500]).cpu().numpy(), 0.6, verbose=False)
path = AudioToFeatures(fe, lsm)
B = len(pcm)
outs = [torch.empty((B, 2000), dtype=torch.float64, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def timed(two_streams, reps=8):
    def step(i):
        if two_streams:
            with torch.cuda.stream(streams[i & 1]):
                path.run(d_pcm, keys, out=outs[i & 1], want_spikes=False)
        else:
            path.run(d_pcm, keys, out=outs[0], want_spikes=False)
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for st in streams:
        st.wait_event(a)
    for i in range(reps):
        step(i)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def cfg(**env):
    for k in ("LSM_WS", "LSM_WS_DEBUG"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})


rows = [("lane = channel kernel, phases in turn (6 CTAs/SM)", dict())]
for dbg, name in ((0, "whole"), (1, "filter + encoder epilogue, no reservoir"), (2, "filter alone")):
    rows.append((f"warp-specialised (filter + epilogue group, reservoir group), {name}", dict(LSM_WS=1, LSM_WS_DEBUG=dbg)))
print("| configuration | one stream ms / 2400 utt | two streams ms / 2400 utt |")
print("|---|---|---|")
for name, env in rows:
    cfg(**env)
    print(f"| {name} | {timed(False):.3f} | {timed(True):.3f} |", flush=True)

