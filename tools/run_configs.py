"""BASELINE.json configs 2-4 on one B200, GPU vs CPU oracle.  Writes a markdown report to stdout.
    python tools/run_configs.py [--per-class 1000] [--skip-acc]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--per-class", type=int, default=1000)
ap.add_argument("--skip-acc", action="store_true")
ap.add_argument("--big-n", action="store_true", help="include N=16000 in config 4 (slow reservoir build)")
args = ap.parse_args()

t0 = time.time()
pcm, labels = synth.synth_dataset(12, args.per_class, workers=os.cpu_count() or 1)   # before CUDA init (forks)
print(f"<!-- synthesised {len(pcm)} utterances in {time.time() - t0:.1f}s -->")

import torch  # noqa: E402
from sklearn.linear_model import LogisticRegression  # noqa: E402
from sklearn.model_selection import train_test_split  # noqa: E402
from sklearn.preprocessing import StandardScaler  # noqa: E402
from lsm_speech_classifier_b200 import _lib, filterbank as fb  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402
from oracle import coracle  # noqa: E402

THR, GAP = [0.70, 0.80, 0.90, 0.95], 0.1
keys = FEATURE_SETS["original"]
mask = _lib.feature_mask(keys)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ------------------------------------------------------------------ config 2
print("## Config 2 - 12 classes, %d utterances (%d train / %d test), 128-ch gammatone, original, multiplier 0.6\n"
      % (len(pcm), int(len(pcm) * 0.8), len(pcm) - int(len(pcm) * 0.8)))
fe = Frontend(128, "gammatone")
t0 = time.time()
X = fe.encode(pcm)
t_enc = time.time() - t0
X_tr, X_te, y_tr, y_te = train_test_split(X, labels, test_size=0.2, random_state=42, stratify=labels)
lsm = build_lsm(X_tr, 0.6, verbose=False)
t0 = time.time()
F_tr = lsm.simulate_batch(X_tr, keys)
F_te = lsm.simulate_batch(X_te, keys)
t_sim = time.time() - t0
print(f"* GPU (host buffers): encode {len(pcm)} utterances {t_enc:.2f} s, reservoir+features {t_sim:.2f} s")
if not args.skip_acc:
    t0 = time.time()
    Xo = coracle.gammatone_encode(pcm, fe.table, fe.params.nwin, fe.params.hop, fe.time_bins, fe.zoom_i0, fe.zoom_f, THR, GAP)
    Xo_tr, Xo_te, _, _ = train_test_split(Xo, labels, test_size=0.2, random_state=42, stratify=labels)
    Fo_tr, _ = coracle.reservoir_run(lsm.reservoir, Xo_tr, mask, True, False)
    Fo_te, _ = coracle.reservoir_run(lsm.reservoir, Xo_te, mask, True, False)
    t_cpu = time.time() - t0
    spikes_equal = bool(np.array_equal(X, Xo))
    n_counts = F_tr[:, :400].size + F_te[:, :400].size
    counts_exact = (np.sum(F_tr[:, :400] == Fo_tr[:, :400]) + np.sum(F_te[:, :400] == Fo_te[:, :400])) / n_counts
    feats_equal = bool(np.array_equal(F_tr, Fo_tr) and np.array_equal(F_te, Fo_te))
    print(f"* CPU oracle, {coracle.num_threads()} threads: {t_cpu:.1f} s for the same {len(pcm)} utterances")
    print(f"* spike trains bit-identical: **{spikes_equal}**; neuron-utterance spike counts exact: **{100 * counts_exact:.4f} %** "
          f"({n_counts} counts); all {F_tr.shape[1]} features bit-identical: **{feats_equal}**")

    def accuracy(a_tr, a_te):
        sc = StandardScaler()
        clf = LogisticRegression(random_state=42, max_iter=1000)
        clf.fit(sc.fit_transform(a_tr), y_tr)
        return float((clf.predict(sc.transform(a_te)) == y_te).mean())

    t0 = time.time()
    acc_gpu = accuracy(F_tr, F_te)
    acc_cpu = acc_gpu if feats_equal else accuracy(Fo_tr, Fo_te)
    t_sk = time.time() - t0
    print(f"* test accuracy (StandardScaler + multinomial LogisticRegression, the reference's train_classifier.py settings): "
          f"GPU features **{100 * acc_gpu:.2f} %**, oracle features **{100 * acc_cpu:.2f} %** "
          f"(difference {100 * abs(acc_gpu - acc_cpu):.2f} pt; scikit-learn scaler + classifier fit {t_sk:.1f} s)")
    # the readout on the device: scaler (bit-exact with scikit-learn) + multinomial logistic regression (same objective)
    from lsm_speech_classifier_b200.readout import LogisticRegression as DevLR, StandardScaler as DevScaler
    t0 = time.time()
    dsc = DevScaler()
    d_tr = dsc.fit_transform(torch.from_numpy(F_tr).cuda())
    d_te = dsc.transform(torch.from_numpy(F_te).cuda())
    dclf = DevLR(random_state=42, max_iter=1000).fit(d_tr, y_tr)
    acc_dev = dclf.score(d_te, y_te)
    t_dev = time.time() - t0
    sk_sc = StandardScaler().fit(F_tr)
    print(f"* device readout (lsm_standardize_* + lsm_logreg_fit): accuracy **{100 * acc_dev:.2f} %** "
          f"(difference to scikit-learn {100 * abs(acc_dev - acc_gpu):.2f} pt), {dclf.n_iter_[0]} L-BFGS iterations, {t_dev:.2f} s including "
          f"the H2D copy of the feature matrix; scaled matrices bit-identical to scikit-learn's: "
          f"**{bool(np.array_equal(d_tr.cpu().numpy(), sk_sc.transform(F_tr)))}**\n")

# ------------------------------------------------------------------ config 3
print("## Config 3 - mel front end, n-filters 64/128/256, multiplier 0.4-1.0\n")
print("| n_filters | K1m ms / 2400 utt | utt/s | spikes == oracle (240 utt) | avg input density |")
print("|---|---|---|---|---|")
d_pcm = torch.from_numpy(pcm[:2400]).cuda()
sub = pcm[:: max(1, len(pcm) // 240)][:240]
mel_spikes = {}
for C in (64, 128, 256):
    fm = Frontend(C, "mel")
    ms = timed(lambda: fm.encode(d_pcm))
    got = fm.encode(sub)
    want = coracle.mel_encode(sub, fm.table, fm.window, fm.tw, fm.tw2, fb.pack_mel_basis(fm.table), fm.params.mel_hop,
                              fm.time_bins, fm.zoom_i0, fm.zoom_f, THR, GAP)
    mel_spikes[C] = got
    print(f"| {C} | {ms:.2f} | {2400 / ms * 1e3:,.0f} | {bool(np.array_equal(got, want))} | {got.mean():.4f} |")
print("\nParticipation (the reference's diagnostic, extract_lsm_features.py:121-125; mean % of the 1000 neurons that fire at least once, "
      "first 24 utterances), GPU raster vs oracle raster:\n")
print("| front end | multiplier | participation GPU % | participation oracle % | rasters identical | regime (reference bands) |")
print("|---|---|---|---|---|---|")
for name, spk in (("mel-128", mel_spikes[128]), ("gammatone-128", X[:: max(1, len(X) // 240)][:240])):
    for mult in (0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0):
        l2 = build_lsm(spk, mult, verbose=False)
        _, rg = l2.simulate_batch(spk[:24], ['spike_counts'], return_raster=True)
        _, ro = coracle.reservoir_run(l2.reservoir, spk[:24], 1, True, True)
        pg = np.mean([(r.sum(0) > 0).mean() * 100 for r in rg])
        po = np.mean([(r.sum(0) > 0).mean() * 100 for r in ro])
        regime = "sub-critical" if pg < 40 else ("saturated" if pg > 98 else "edge of chaos")
        print(f"| {name} | {mult:.1f} | {pg:.1f} | {po:.1f} | {bool(np.array_equal(rg, ro))} | {regime} |")
        l2.close()

# ------------------------------------------------------------------ config 4
print("\n## Config 4 - reservoir size sweep at batch 4096 (event-driven integer kernel; k = 0.2 N)\n")
print("| N | k | W plane MB | build s | K2 ms / 4096 utt | utt/s | G neuron-steps/s | mean spikes/step | raster == oracle (8 utt) |")
print("|---|---|---|---|---|---|---|---|---|")
spk4096 = torch.from_numpy(np.concatenate([X] * (4096 // len(X) + 1))[:4096]).cuda()
for N in (1000, 4000) + ((16000,) if args.big_n else ()):
    t0 = time.time()
    l4 = build_lsm(X[:500], 0.6, num_neurons=N, verbose=False)
    tb = time.time() - t0
    ms = timed(lambda: l4.simulate_batch(spk4096, keys), reps=2)
    _, rg = l4.simulate_batch(X[:8], ['spike_counts'], return_raster=True)
    _, ro = coracle.reservoir_run(l4.reservoir, X[:8], 1, True, True)
    npad = l4.ctx.lib and None
    print(f"| {N} | {int(0.2 * N)} | {(N + 1) * ((N + 127) // 128 * 128) * 4 / 1e6:.0f} | {tb:.1f} | {ms:.1f} | {4096 / ms * 1e3:,.0f} | "
          f"{4096 / ms * 1e3 * N * 400 / 1e9:.1f} | {rg.sum() / (8 * 400):.1f} | {bool(np.array_equal(rg, ro))} |")
    l4.close()
