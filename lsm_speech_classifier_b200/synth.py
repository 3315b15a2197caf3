"""Deterministic synthetic utterances (1 s, 16 kHz, float32 mono, exactly 16000 samples).

The reference reads Google Speech Commands WAVs through ``load_audio_file``
(/root/reference/create_dataset.py:22-36): float32, mono, resampled to 16 kHz and
zero-padded / truncated to exactly 16000 samples.  There is no dataset and no network
here, so the measured path starts from PCM that honours that contract.

Generator: a class-dependent voiced "syllable" - a 40-harmonic source shaped
by three gliding formant resonances - under a raised-cosine envelope with a random onset
and duration, plus white noise at -30 dBFS, peak-normalised to 0.5.  Seed = 1234 + class_id*100003 + utt_id, so any
utterance can be regenerated independently (and on any rank).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_SAMPLES = 16000


def _class_formants(class_id: int):
    """Per-class "vowel": three formant centres (200-3500 Hz), their glide over the
    syllable, and a pitch register.  Fixed by the class index alone."""
    rng = np.random.default_rng(977 + 31 * int(class_id))
    centres = np.array([rng.uniform(250.0, 900.0), rng.uniform(900.0, 2200.0), rng.uniform(2200.0, 3500.0)])
    glide = rng.uniform(-0.30, 0.30, size=3)
    pitch = rng.uniform(95.0, 240.0)
    return centres, glide, pitch


_N_HARM = 40


def synth_utterance(class_id: int, utt_id: int) -> np.ndarray:
    """One utterance -> float32[16000]: a harmonic (glottal-like) source whose harmonics are
    weighted by three gliding formant resonances, under a raised-cosine syllable envelope,
    plus white noise at -30 dBFS; peak-normalised to 0.5."""
    rng = np.random.default_rng(1234 + int(class_id) * 100003 + int(utt_id))
    centres, glide, pitch = _class_formants(class_id)
    t = np.arange(N_SAMPLES, dtype=np.float64) / SAMPLE_RATE
    onset = rng.uniform(0.1, 0.4)
    dur = rng.uniform(0.3, 0.5)
    f0 = pitch * rng.uniform(0.9, 1.1)
    bend = rng.uniform(-0.15, 0.15)
    jitter = rng.uniform(0.95, 1.05, size=3)
    amps = np.array([1.0, 0.7, 0.45]) * rng.uniform(0.7, 1.3, size=3)
    bw = np.array([90.0, 140.0, 220.0])
    # everything below is non-zero only inside the syllable: work on that span
    lo = max(0, int(np.floor(onset * SAMPLE_RATE)))
    hi = min(N_SAMPLES, int(np.ceil((onset + dur) * SAMPLE_RATE)) + 1)
    ts = t[lo:hi]
    u = np.clip((ts - onset) / dur, 0.0, 1.0)
    env = np.where((ts >= onset) & (ts <= onset + dur), 0.5 - 0.5 * np.cos(2 * np.pi * u), 0.0)
    # pitch contour f0*(1 + bend*u); phi is its running integral
    phi = 2 * np.pi * f0 * ((ts - onset) + 0.5 * bend * dur * u * u)
    h = np.arange(1, _N_HARM + 1, dtype=np.float64)[:, None]
    fh = h * (f0 * (1.0 + bend * u))[None, :]                       # [H, span] harmonic frequencies
    gain = np.zeros_like(fh)
    for k in range(3):
        fk = centres[k] * jitter[k] * (1.0 + glide[k] * u)
        gain += amps[k] / (1.0 + ((fh - fk[None, :]) / bw[k]) ** 2)
    gain *= (fh < 0.45 * SAMPLE_RATE)
    phase0 = rng.uniform(0, 2 * np.pi, size=(_N_HARM, 1))
    sig = np.zeros(N_SAMPLES, dtype=np.float64)
    sig[lo:hi] = np.sum(gain * np.sin(h * phi[None, :] + phase0), axis=0) * env
    peak = np.max(np.abs(sig))
    if peak > 0:
        sig *= 0.5 / peak
    noise = rng.standard_normal(N_SAMPLES) * (10.0 ** (-30.0 / 20.0))
    out = sig + noise
    out *= 0.5 / np.max(np.abs(out))
    return out.astype(np.float32)


def _synth_one(args):
    return synth_utterance(*args)


def synth_dataset(n_classes: int, per_class: int, start_utt: int = 0, workers: int = 1):
    """Class-major dataset like the reference's directory walk
    (/root/reference/create_dataset.py:130-162): label = class position, utterances
    in order inside a class.  Returns (pcm float32[S,16000], labels int32[S])."""
    S = n_classes * per_class
    pcm = np.empty((S, N_SAMPLES), dtype=np.float32)
    labels = np.empty(S, dtype=np.int32)
    jobs = [(c, start_utt + u) for c in range(n_classes) for u in range(per_class)]
    labels[:] = [c for c, _ in jobs]
    if workers > 1 and S >= 4 * workers:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            for i, w in enumerate(pool.imap(_synth_one, jobs, chunksize=16)):
                pcm[i] = w
    else:
        for i, j in enumerate(jobs):
            pcm[i] = synth_utterance(*j)
    return pcm, labels
