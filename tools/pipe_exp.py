"""Timing experiments on the warp-specialised kernel (one B200): whole step, filter warps alone, encoder/reservoir units alone,
the lane = channel kernel, and the copy engine's host-to-device rate.  python tools/pipe_exp.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

pcm, _ = synth.synth_dataset(12, 200, workers=os.cpu_count() or 1)
import torch  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

VARIANT = sys.argv[1] if len(sys.argv) > 1 else "1"
os.environ["LSM_PIPELINE"] = VARIANT          # the warp-specialised kernel is opt-in (1: TMA-fed filter warps, 2: energy-unit filter warps)
keys = FEATURE_SETS["original"]
fe = Frontend(128, "gammatone")
d_pcm = torch.from_numpy(pcm).cuda()
lsm = build_lsm(fe.encode(d_pcm[:500]).cpu().numpy(), 0.6, verbose=False)
path = AudioToFeatures(fe, lsm)
B = len(pcm)
out = torch.empty((B, 2000), dtype=torch.float64, device="cuda")


def timed(fn, reps=8):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def step():
    path.run(d_pcm, keys, out=out, want_spikes=False)


print(f"whole step (two half launches on the two lanes): {timed(step):.3f} ms")
os.environ["LSM_PIPE_DEBUG"] = "1"
print(f"filter warps alone:                              {timed(step):.3f} ms")
os.environ["LSM_PIPE_ONE_PIECE"] = "1"
print(f"filter alone, one launch of 2400, 1 CTA/SM:      {timed(step):.3f} ms")
os.environ["LSM_PIPE_GRID_MULT"] = "2"
print(f"filter alone, one launch of 2400, 2 CTAs/SM:     {timed(step):.3f} ms")
del os.environ["LSM_PIPE_DEBUG"]
print(f"whole step, one launch of 2400, 2 CTAs/SM:       {timed(step):.3f} ms")
del os.environ["LSM_PIPE_GRID_MULT"]
print(f"whole step, one launch of 2400, 1 CTA/SM:        {timed(step):.3f} ms")
del os.environ["LSM_PIPE_ONE_PIECE"]
os.environ["LSM_PIPE_DEBUG"] = "2"
print(f"encoder/reservoir units alone:                   {timed(step):.3f} ms")
del os.environ["LSM_PIPE_DEBUG"]
del os.environ["LSM_PIPELINE"]
print(f"lane = channel fused kernel (one launch):        {timed(step):.3f} ms")
os.environ["LSM_PIPELINE"] = VARIANT
for n in (1200, 600, 4800):
    x = torch.cat([d_pcm, d_pcm])[:n].contiguous()
    o = torch.empty((n, 2000), dtype=torch.float64, device="cuda")
    print(f"B = {n}: {timed(lambda: path.run(x, keys, out=o, want_spikes=False)):.3f} ms")

# copy engine: pinned host -> device
h = torch.from_numpy(pcm).pin_memory()
d = torch.empty_like(d_pcm)
ms = timed(lambda: d.copy_(h, non_blocking=True))
print(f"torch pinned H2D {h.numel() * 4 / 1e6:.0f} MB: {ms:.3f} ms = {h.numel() * 4 / ms / 1e6:.1f} GB/s")
h_out = torch.empty((B, 2000), dtype=torch.float64).pin_memory()
t0 = time.perf_counter()
for i in range(8):
    path.run_host_async(h, keys, out=h_out, lane=i & 1)
fe.ctx.sync_all()
print(f"run_host_async float32, 8 steps on alternating lanes: {(time.perf_counter() - t0) / 8 * 1e3:.3f} ms per step")
