"""Host feeds of the asynchronous call (zero-copy vs copy engine), repeated: python tools/feed_exp.py"""
import os
import sys
import time

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
lsm = build_lsm(fe.encode(pcm[:500]), 0.6, verbose=False)
path = AudioToFeatures(fe, lsm)
ctx = fe.ctx
B = len(pcm)
h_pcm = torch.from_numpy(pcm).pin_memory()
h_i16 = torch.from_numpy(np.clip(np.round(pcm * 32768.0), -32768, 32767).astype(np.int16)).pin_memory()
h_out = [torch.empty((B, 2000), dtype=torch.float64).pin_memory() for _ in range(2)]


def loop(h_in, feed, steps=10):
    ctx.set_host_feed(feed)
    ctx.sync_all(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        path.run_host_async(h_in, keys, out=h_out[i & 1], lane=i & 1)
    t_enq = time.perf_counter() - t0
    ctx.sync_all()
    dt = time.perf_counter() - t0
    return B * steps / dt, t_enq / steps * 1e3


for rep in range(3):
    for feed in ("zero_copy", "copy_engine"):
        for name, h in (("f32", h_pcm), ("i16", h_i16)):
            v, enq = loop(h, feed)
            print(f"rep {rep} {feed:12s} {name}: {v:9.0f} utt/s, enqueue {enq:.3f} ms per call", flush=True)
out_np = np.empty((B, 2000))
for rep in range(3):
    t0 = time.perf_counter()
    path.run_host(pcm, keys, out=out_np)
    print(f"pageable run_host: {B / (time.perf_counter() - t0):9.0f} utt/s")
