"""Mint the committed golden vectors.  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden.py

What is pinned against the reference ITSELF (its code imported and run verbatim, with empty stub
modules standing in for the three packages it cannot import here - librosa, gammatone, snnpy):
  * convert_spectrogram_to_spikes_hysteresis, create_pure_redundancy  (create_dataset.py:81-104)
  * calculate_theoretical_w_critico, FEATURE_SETS                     (extract_lsm_features.py:19-60)
What is pinned against the reference's installed dependencies (scipy/numpy) through oracle/pyref.py:
  * lfilter cascade, window mean, dB, normalise, zoom  -> spec_norm and spikes for real PCM
What is only self-consistent (third-party source unavailable; "parity unpinned"):
  * gammatone filter design, mel front end, reservoir dynamics and features.
Nothing under tests/ reads /root/reference at test time; only this script does.
"""
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


def import_reference():
    for name in ("librosa", "gammatone", "gammatone.gtgram", "snnpy", "snnpy.snn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["gammatone"].gtgram = sys.modules["gammatone.gtgram"]
    sys.modules["snnpy.snn"].SNN = object
    sys.modules["snnpy.snn"].SimulationParams = object
    sys.path.insert(0, "/root/reference")
    import create_dataset as ref_cd
    import extract_lsm_features as ref_ex
    return ref_cd, ref_ex


def array_digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes())
    return h.hexdigest()


def main():
    ref_cd, ref_ex = import_reference()
    from oracle import pyref
    from lsm_speech_classifier_b200 import filterbank, synth
    from lsm_speech_classifier_b200.reservoir import SimulationParams, build_reservoir

    rng = np.random.default_rng(20261018)

    # ---- 1. encoder KATs: the reference's own encoder on crafted + random spectrograms
    thr = ref_cd.SPIKE_THRESHOLDS
    gap = ref_cd.HYSTERESIS_GAP
    specs64 = []
    ramp = np.linspace(0.0, 1.0, 100)
    specs64.append(np.stack([ramp, ramp[::-1], np.abs(np.sin(np.linspace(0, 9, 100))),
                             np.full(100, 0.95), np.full(100, 0.9500000000000001), np.full(100, 0.85),
                             np.where(np.arange(100) % 7 < 3, 0.96, 0.849999),
                             np.where(np.arange(100) % 5 < 2, 0.71, 0.6),
                             np.where(np.arange(100) % 5 < 2, 0.7000000000000001, 0.6000000000000001),
                             np.zeros(100), np.ones(100), np.full(100, 0.7)]))
    for _ in range(6):
        walk = np.cumsum(rng.normal(0, 0.06, size=(24, 100)), axis=1) + rng.uniform(0.4, 1.0, size=(24, 1))
        specs64.append(np.clip(walk, 0.0, 0.9999999))
    enc_in64 = np.concatenate(specs64, axis=0)
    enc_out64 = ref_cd.convert_spectrogram_to_spikes_hysteresis(enc_in64, thr, gap)
    enc_in32 = enc_in64.astype(np.float32)
    enc_out32 = ref_cd.convert_spectrogram_to_spikes_hysteresis(enc_in32, thr, gap)
    red3 = ref_cd.create_pure_redundancy(enc_out64[:5], 3)
    np.savez_compressed(os.path.join(HERE, "encoder_kats.npz"), spec64=enc_in64, spikes64=enc_out64,
                        spec32=enc_in32, spikes32=enc_out32, redundancy3=red3,
                        thresholds=np.array(thr), gap=np.array(gap))

    # ---- 2. w_critico KATs from the reference's own function
    class P:  # the three fields calculate_theoretical_w_critico reads back
        small_world_graph_k = 200
        membrane_threshold = 2.0
        refractory_period = 2
    sets, answers = [], []
    for dens, n in ((0.1, 7), (0.03, 3), (0.0, 2)):
        data = (rng.random((n, 16, 40)) < dens).astype(np.uint8)
        sets.append(data)
        answers.append(ref_ex.calculate_theoretical_w_critico(P, data))
    exact10 = np.zeros((10, 10, 10), np.uint8)
    exact10[:, :, 0] = 1  # density exactly 0.1 -> (2 - 0.4)/100
    answers.append(ref_ex.calculate_theoretical_w_critico(P, exact10))
    np.savez_compressed(os.path.join(HERE, "w_critico_kats.npz"), d0=sets[0], d1=sets[1], d2=sets[2],
                        d3=exact10, answers=np.array(answers),
                        feature_sets=np.array([f"{k}:{','.join(v)}" for k, v in ref_ex.FEATURE_SETS.items()]))

    # ---- 3. gammatone front end on real PCM: scipy/numpy restatement + the reference's encoder
    pcm = np.stack([synth.synth_utterance(0, 0), synth.synth_utterance(3, 1), synth.synth_utterance(7, 2),
                    np.zeros(16000, np.float32),                                  # silent clip -> zeros (:64-65)
                    np.concatenate([synth.synth_utterance(5, 3)[4000:9000], np.zeros(11000, np.float32)])])  # zero padded (:28-30)
    coefs = filterbank.gammatone_coefs(16000, 128, 50)
    spec_norm, spikes = [], []
    for a in pcm:
        s = pyref.audio_to_spectrogram(a, 128, "gammatone", coefs=coefs)
        spec_norm.append(np.asarray(s, dtype=np.float64))
        spikes.append(ref_cd.convert_spectrogram_to_spikes_hysteresis(s, thr, gap))
    np.savez_compressed(os.path.join(HERE, "frontend_gammatone.npz"), pcm=pcm, coefs=coefs,
                        spec_norm=np.stack(spec_norm)[:2], spikes_packed=np.packbits(np.stack(spikes), axis=-1))

    # ---- 4. mel front end (64 channels keeps the file small)
    basis = filterbank.mel_basis(16000, 2048, 64)
    mspec, mspikes = [], []
    for a in pcm[:3]:
        s = pyref.audio_to_spectrogram(a, 64, "mel", coefs=basis)
        mspec.append(s)
        mspikes.append(ref_cd.convert_spectrogram_to_spikes_hysteresis(s, thr, gap))
    np.savez_compressed(os.path.join(HERE, "frontend_mel64.npz"), spec_norm=np.stack(mspec)[:1],
                        spikes_packed=np.packbits(np.stack(mspikes), axis=-1))

    # ---- 5. reservoir: frozen spec, python restatement (parity unpinned vs snnpy)
    X = np.stack(spikes)
    wc = ref_ex.calculate_theoretical_w_critico(P, X)
    out = {}
    for tag, kw in (("n1000", dict()), ("n256_hetero", dict(num_neurons=256, small_world_graph_k=50,
                                                             num_output_neurons=100, leak_variance_divisor=4.0))):
        k = kw.get("small_world_graph_k", 200)
        p = SimulationParams(mean_weight=wc * 0.6 * 200 / k, input_spike_times=X[0], **kw)
        r = build_reservoir(p)
        rasters, feats = [], []
        for b in (0, 2, 3):
            ras = pyref.simulate(X[b], r.w_rowptr, r.w_col, r.w_q, r.w_shift, r.in_rowptr, r.in_col, r.in_val,
                                 r.leak, r.theta, r.refractory)
            fd = pyref.features_from_raster(ras, r.out_idx, r.refractory)
            rasters.append(np.packbits(ras, axis=-1))
            feats.append(np.concatenate([fd[key] for key in pyref.FEATURE_KEYS]))
        out[f"{tag}_raster_packed"] = np.stack(rasters)
        out[f"{tag}_features"] = np.stack(feats)
        out[f"{tag}_digest"] = np.array(array_digest(r.w_rowptr, r.w_col, r.w_q, r.in_rowptr, r.in_col, r.in_val,
                                                     r.out_idx, r.leak))
        out[f"{tag}_mean_weight"] = np.array(p.mean_weight)
    out["utt_index"] = np.array([0, 2, 3])
    np.savez_compressed(os.path.join(HERE, "reservoir.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
