"""Stage 1 of the pipeline: audio -> spike-train dataset -> speech_spike_dataset_pure_redundancy.npz.

Same call surface as /root/reference/create_dataset.py (constants :10-17, load_audio_file :22-36,
audio_to_spectrogram :39-78, convert_spectrogram_to_spikes_hysteresis :81-98, create_pure_redundancy
:101-104, create_dataset :107-177, flags :183-192).  The per-file loop at :143 becomes one batched
kernel launch; the .npz has the reference's keys, dtypes, shapes and class-major sample order.
"""
from __future__ import annotations

import argparse
import ctypes as C
from pathlib import Path

import numpy as np

from . import _lib, npzio, timing
from .frontend import (DURATION, HYSTERESIS_GAP, REDUNDANCY_FACTOR, SAMPLE_RATE, SPIKE_THRESHOLDS,  # noqa: F401
                       TIME_BINS, Frontend)

MAX_SAMPLES_PER_CLASS = 1000
OUTPUT_FILE = "speech_spike_dataset_pure_redundancy.npz"
COMMANDS = ["yes", "no", "up", "visual", "backward", "stop", "bird", "cat", "nine", "eight", "zero", "follow"]

np.random.seed(42)

_frontends: dict = {}


def _frontend(n_filters: int, filterbank: str, redundancy: int = REDUNDANCY_FACTOR) -> Frontend:
    key = (int(n_filters), filterbank, int(redundancy))
    if key not in _frontends:
        _frontends[key] = Frontend(n_filters, filterbank, redundancy=redundancy)
    return _frontends[key]


def load_audio_file(filepath: Path):
    """reference :22-36 contract: float32 mono at 16 kHz, exactly 16000 samples (zero padded or truncated), None on failure
    (with the reference's message).  PCM and float WAV files of any rate and channel count (ingest.py: RIFF parser, channel
    mean, polyphase resampler on the GPU pinned to scipy.signal.resample_poly); other containers are reported and skipped."""
    from . import ingest
    try:
        return ingest.load_audio(filepath, SAMPLE_RATE, DURATION)
    except Exception as e:
        print(f"Error loading {filepath}: {e}")
        return None


def audio_to_spectrogram(audio: np.ndarray, n_filters: int, filterbank: str) -> np.ndarray:
    """reference :39-78 for one utterance: normalised, time-resampled spectrogram [n_filters, 100]
    (float64 for gammatone, float32 for mel, float32 zeros for a degenerate clip)."""
    fe = _frontend(n_filters, filterbank)
    _, spec = fe.encode(np.asarray(audio, dtype=np.float32)[None], return_spectrogram=True)
    spec = spec[0]
    if not spec.any():
        return np.zeros((n_filters, TIME_BINS), dtype=np.float32)
    return spec.astype(np.float32) if filterbank == "mel" else spec


def convert_spectrogram_to_spikes_hysteresis(spectrogram, thresholds, hysteresis_gap=0.05):
    """reference :81-98 on the GPU (lsm_hysteresis_encode): uint8[n_filters, n_time*len(thresholds)]."""
    import torch
    spec = np.ascontiguousarray(spectrogram)
    if spec.dtype not in (np.float32, np.float64):
        spec = spec.astype(np.float64)
    n_filters, n_time = spec.shape
    thr = np.array(sorted(thresholds, reverse=True), dtype=np.float64)
    lower = np.array([t - hysteresis_gap for t in thr], dtype=np.float64)
    ctx = _lib.context()
    d_spec = torch.from_numpy(spec).cuda(ctx.device)
    out = torch.empty((n_filters, n_time * len(thr)), dtype=torch.uint8, device=d_spec.device)
    ctx.set_stream(torch.cuda.current_stream(d_spec.device).cuda_stream)
    ctx.check(ctx.lib.lsm_hysteresis_encode(ctx.h, C.c_void_p(d_spec.data_ptr()), int(spec.dtype == np.float32), 1,
                                            n_filters, n_time, _lib._np_ptr(thr), _lib._np_ptr(lower), len(thr), 1,
                                            C.c_void_p(out.data_ptr())))
    return out.cpu().numpy()


def create_pure_redundancy(spike_train: np.ndarray, redundancy_factor: int) -> np.ndarray:
    """reference :101-104."""
    return np.repeat(spike_train, redundancy_factor, axis=0)


def encode_batch(pcm: np.ndarray, n_filters: int, filterbank: str, redundancy: int = REDUNDANCY_FACTOR) -> np.ndarray:
    """float32[S,16000] -> uint8[S, n_filters*redundancy, 400]; sharded across ranks under torchrun."""
    from .distributed import sharded_spikes
    return sharded_spikes(_frontend(n_filters, filterbank, redundancy), np.ascontiguousarray(pcm, dtype=np.float32))


def _collect_wavs():
    """reference :121-146: class-major, sorted file names, first 1000 per class, unreadable files skipped."""
    base = Path("speech_commands_v0.02")
    clips, labels = [], []
    for label_idx, command in enumerate(COMMANDS):
        print(f"Processing '{command}'...")
        command_dir = base / command
        if not command_dir.is_dir():
            print(f"  Warning: Directory not found, skipping: {command_dir}")
            continue
        audio_files = sorted(command_dir.glob("*.wav"))[:MAX_SAMPLES_PER_CLASS]
        if not audio_files:
            print(f"  Warning: No files found for '{command}'")
            continue
        for f in audio_files:
            a = load_audio_file(f)
            if a is None:
                continue
            clips.append(a)
            labels.append(label_idx)
    return clips, labels


PACKED_FILE = "speech_spike_dataset_packed.npz"


def save_packed_spikes(filename, X_spikes: np.ndarray, y_labels: np.ndarray):
    """(extension) the spike trains as bits: X_spikes_bits uint8[S, C, ceil(T/8)] (np.packbits along the
    time axis), n_steps, y_labels - an eighth of the reference file before compression.  `load_spike_dataset` reads it back to
    exactly the reference's arrays when the reference-schema file is absent."""
    X = np.asarray(X_spikes)
    if X.dtype != np.uint8 or X.ndim != 3 or (X > 1).any():
        raise ValueError("X_spikes must be uint8[S, C, T] of zeros and ones")
    npzio.savez_compressed(filename, X_spikes_bits=np.packbits(X, axis=2), n_steps=np.int32(X.shape[2]),
                        y_labels=np.asarray(y_labels, dtype=np.int32))


def load_packed_spikes(filename):
    data = np.load(filename)
    T = int(data["n_steps"])
    return np.unpackbits(data["X_spikes_bits"], axis=2, count=T), data["y_labels"]


def collect_pcm(synthetic: tuple | None = None):
    """The utterances of a run as float32[S, 16000] + labels: the synthetic generator or the reference's directory walk."""
    if synthetic is not None:
        from . import synth
        import os
        return synth.synth_dataset(int(synthetic[0]), int(synthetic[1]), workers=os.cpu_count() or 1)
    clips, labels = _collect_wavs()
    if not clips:
        print("\nERROR: No audio files were successfully processed.")
        return None, None
    return np.stack(clips), np.array(labels, dtype=np.int32)


def create_dataset(n_filters: int, filterbank: str, synthetic: tuple | None = None, packed: bool = False):
    """reference :107-177.  `synthetic=(n_classes, per_class)` replaces the directory walk with the
    deterministic generator (synth.py) - there is no Speech Commands copy in this environment.
    `packed=True` writes the bit-packed file (save_packed_spikes) instead of the reference-schema one."""
    from .distributed import init_from_env, is_main
    init_from_env()             # under torchrun (also when this stage is run on its own): one rank per GPU, utterances sharded
    if is_main():
        print(f"Creating dataset with filterbank: {filterbank}, filters: {n_filters}")
    with timing.stage("audio: synthesis / WAV decode"):
        pcm, labels = collect_pcm(synthetic)
    if pcm is None:
        return
    if timing.enabled:
        # attribute the one-time costs (CUDA context, library load, filter tables, pinned staging ring, host copy threads) to their
        # own line: a warm-up call on silence sized like one piece of the real run
        with timing.stage("GPU init (one-time): CUDA context, tables, pinned staging"):
            _frontend(n_filters, filterbank).encode(np.zeros((min(len(pcm), 768), pcm.shape[1]), np.float32))
    with timing.stage("stage 1 compute: audio -> spike trains (GPU, host buffers)", len(pcm)):
        X_spikes = encode_batch(pcm, n_filters, filterbank)
    y_labels = np.array(labels, dtype=np.int32)
    if not is_main():
        return
    counts = X_spikes.reshape(len(X_spikes), -1).sum(axis=1, dtype=np.int64)
    print("\nDataset created successfully.")
    print(f"  Shape: {X_spikes.shape}")
    print(f"  Avg spikes per sample: {np.mean(counts):.1f}")
    with timing.stage("stage 1 write: spike file (parallel deflate)"):
        if packed:
            save_packed_spikes(PACKED_FILE, X_spikes, y_labels)
        else:
            npzio.savez_compressed(OUTPUT_FILE, X_spikes=X_spikes, y_labels=y_labels)
    print(f"Saved to '{PACKED_FILE}' (bit-packed)" if packed else f"Saved to '{OUTPUT_FILE}'")


def _cli(argv=None):
    parser = argparse.ArgumentParser(description="Create a spike train dataset from audio files.")
    parser.add_argument("--n-filters", type=int, default=128, help="Number of filters for the filterbank.")
    parser.add_argument("--filterbank", type=str, default="gammatone", choices=["mel", "gammatone"],
                        help="Type of filterbank to use.")
    parser.add_argument("--synthetic", type=int, nargs=2, metavar=("CLASSES", "PER_CLASS"), default=None,
                        help="(extension) use the deterministic synthetic generator instead of speech_commands_v0.02/")
    parser.add_argument("--packed", action="store_true", help="(extension) write the bit-packed spike file instead")
    args = parser.parse_args(argv)
    create_dataset(n_filters=args.n_filters, filterbank=args.filterbank, synthetic=args.synthetic, packed=args.packed)


if __name__ == "__main__":
    _cli()
