"""pipeline2 (energy-unit filter role) against the default kernel, bench-style: whole 2400-utterance launches on one / two streams.
    python tools/p2_exp.py"""
import os
import sys

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


def timed(two_streams, reps=10):
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
    for k in ("LSM_PIPELINE", "LSM_PIPE_ONE_PIECE", "LSM_WS"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})


cfg()
path.run(d_pcm, keys, out=outs[0], want_spikes=False); torch.cuda.synchronize()
want = outs[0].clone()
print("| configuration | one stream ms / 2400 utt | two streams ms / 2400 utt | same features |")
print("|---|---|---|---|")
for name, env in (("lane = channel kernel (default)", dict()),
                  ("pipeline2 (energy-unit filter role), one launch per call, lanes alternate", dict(LSM_PIPELINE=2)),
                  ("pipeline (TMA-fed), one launch per call", dict(LSM_PIPELINE=1, LSM_PIPE_ONE_PIECE=1))):
    cfg(**env)
    t1, t2 = timed(False), timed(True)
    torch.cuda.synchronize()
    print(f"| {name} | {t1:.3f} | {t2:.3f} | {bool(torch.equal(outs[0], want))} |", flush=True)
