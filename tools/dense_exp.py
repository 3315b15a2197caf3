"""Dense (tensor-core) reservoir arm vs the event-driven arm on one B200: contraction check, raster / feature identity, timings.
    python tools/dense_exp.py [--quick]            (writes a markdown table to stdout; profiles/r2_config4.md is made from it)"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true", help="contraction + identity checks only")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--sizes", type=str, default="1000,4000")
ap.add_argument("--dense-max", type=int, default=4000, help="largest N the dense arm is timed at")
ap.add_argument("--mults", type=str, default="0.4,0.6,0.8,0.9,1.0")
args = ap.parse_args()

pcm, _ = synth.synth_dataset(12, 43, workers=min(16, os.cpu_count() or 1))      # 516 utterances, before CUDA init (forks)

import torch  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402

keys = FEATURE_SETS["original"]
fe = Frontend(128, "gammatone")
X = fe.encode(pcm)


def dense_w(r):
    W = np.zeros((r.num_neurons, r.num_neurons), dtype=np.int64)          # [post][pre]
    post = np.repeat(np.arange(r.num_neurons), np.diff(r.w_rowptr))
    np.add.at(W, (post, r.w_col), r.w_q)
    return W


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ---- 1. the contraction alone: tensor cores + digit-plane recombination vs the integer matrix product
lsm = build_lsm(X[:500], 0.6, verbose=False)
r = lsm.reservoir
rs = np.random.RandomState(7)
S = (rs.random_sample((300, r.num_neurons)) < 0.05).astype(np.uint8)
S[0] = 0; S[1] = 1; S[2] = 0; S[2, 5] = 1; S[3] = 0; S[3, 999] = 1
got = lsm.dense_probe(S).cpu().numpy().astype(np.int64)
want = S.astype(np.int64) @ dense_w(r).T
ok = np.array_equal(got, want)
print(f"contraction (300 x {r.num_neurons}, 5 % density + all-zero / all-one / single-spike rows): identical = {ok}")
if not ok:
    bad = np.argwhere(got != want)
    print(f"  {len(bad)} of {got.size} entries differ; first rows/cols: {bad[:8].tolist()}")
    print(f"  rows with errors: {np.unique(bad[:, 0])[:20].tolist()} ...; cols with errors (mod 128 histogram): "
          f"{np.bincount(bad[:, 1] % 128, minlength=128).tolist()}")
    for b, i in bad[:6]:
        print(f"  got[{b},{i}] = {got[b, i]}  want {want[b, i]}  ratio {got[b, i] / max(1, want[b, i]):.4f}")
    print(f"  single-spike row 2 (pre = 5): got nonzero at {np.nonzero(got[2])[0][:10].tolist()} want {np.nonzero(want[2])[0][:10].tolist()}")
    print(f"  all-one row: got[:8] {got[1, :8].tolist()} want {want[1, :8].tolist()}")
    sys.exit(1)

# ---- 2. whole simulation: rasters and all eight features, dense vs event-driven, several regimes
for mult in (0.6, 1.0):
    l2 = build_lsm(X[:500], mult, verbose=False)
    fe_, re_ = l2.simulate_batch(X[:140], nan_to_num=False, return_raster=True)
    l2.set_mode("dense")
    fd_, rd_ = l2.simulate_batch(X[:140], nan_to_num=False, return_raster=True)
    l2.set_mode("event")
    print(f"multiplier {mult}: rasters identical = {np.array_equal(re_, rd_)}, features identical = {np.array_equal(fe_, fd_, equal_nan=True)}, "
          f"mean spikes per step {re_.sum() / (140 * 400):.1f}")
    if not np.array_equal(re_, rd_):
        d = np.argwhere(re_ != rd_)
        print(f"  first differences (utt, t, neuron): {d[:6].tolist()}; first differing step {d[:, 1].min()}")
        sys.exit(1)
    l2.close()
if args.quick:
    sys.exit(0)

# ---- 3. timings at batch B, both arms, multiplier sweep and reservoir sizes
B = args.batch
spk = torch.from_numpy(np.concatenate([X] * (B // len(X) + 1))[:B]).cuda()
print(f"\n| N | multiplier | mean spikes/step | event-driven ms / {B} utt | dense ms / {B} utt | dense: digit planes | same features |")
print("|---|---|---|---|---|---|---|")
for N in [int(v) for v in args.sizes.split(",")]:
    for mult in [float(v) for v in (args.mults if N <= 4000 else "0.6,1.0").split(",")]:
        t0 = time.time()
        l4 = build_lsm(X[:500], mult, num_neurons=N, verbose=False)
        ms_e = timed(lambda: l4.simulate_batch(spk, keys))
        f_e = l4.simulate_batch(spk, keys)
        part, _, avg = l4.diagnostics(spk[:256])
        if N <= args.dense_max:
            l4.set_mode("dense")
            ms_d = timed(lambda: l4.simulate_batch(spk, keys), reps=1)
            f_d = l4.simulate_batch(spk, keys)
            same = bool(torch.equal(f_e, f_d))
            wmax = int(l4.reservoir.w_q.max())
            planes = 3 if wmax >= 65536 else (2 if wmax >= 256 else 1)
            print(f"| {N} | {mult:.1f} | {avg.mean() * N / 400:.1f} | {ms_e:.2f} | {ms_d:.1f} | {planes} | {same} |", flush=True)
        else:
            print(f"| {N} | {mult:.1f} | {avg.mean() * N / 400:.1f} | {ms_e:.2f} | - | - | - |", flush=True)
        l4.close()

# ---- 4. strict reservoirs (fp64 weights, ascending-order sums) against the quantised default, event-driven arm
from lsm_speech_classifier_b200.snn import SNN, SimulationParams  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import calculate_theoretical_w_critico  # noqa: E402
print(f"\n| N = 1000, multiplier | quantised ms / {B} utt | strict (fp64, ordered) ms / {B} utt |")
print("|---|---|---|")
for mult in (0.6, 1.0):
    row = []
    for q in (True, False):
        prm = SimulationParams(input_spike_times=X[0], quantize_weights=q)
        prm.mean_weight = calculate_theoretical_w_critico(prm, X[:500], verbose=False) * mult
        l5 = SNN(simulation_params=prm)
        row.append(timed(lambda: l5.simulate_batch(spk, keys)))
        l5.close()
    print(f"| {mult:.1f} | {row[0]:.2f} | {row[1]:.2f} |", flush=True)
