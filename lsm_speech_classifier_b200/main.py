"""Orchestrator with the reference's flags (/root/reference/main.py:35-57) and stage order (:19-27).
The stages run in-process instead of through three os.system calls; the two .npz hand-off files are
still written, so the reference's train_classifier.py consumes the result unchanged."""
from __future__ import annotations

import argparse


def run_pipeline(n_filters: int, filterbank: str, feature_set: str, multiplier: float, synthetic=None,
                 train: bool = True, readout: str = "sklearn", fused: bool = False, packed: bool = False, show_timing: bool = False):
    from . import create_dataset as cd, extract_lsm_features as ex, timing
    from .distributed import init_from_env, is_main
    init_from_env()
    timing.enabled = bool(show_timing)
    if is_main():
        print("--- Running Pipeline ---")
    if fused:
        # (extension) no spike file between the stages: audio -> features in one pass
        print("\n--- Steps 1+2: Audio -> LSM Features (fused) ---")
        with timing.stage("audio: synthesis / WAV decode"):
            pcm, labels = cd.collect_pcm(synthetic)
        if pcm is not None:
            ex.main_fused(pcm, labels, n_filters, filterbank, feature_set, multiplier)
    else:
        if is_main():
            print("\n--- Step 1: Creating Spike Train Dataset ---")
        cd.create_dataset(n_filters=n_filters, filterbank=filterbank, synthetic=synthetic, packed=packed)
        _barrier()
        if is_main():
            print("\n--- Step 2: Extracting LSM Features ---")
        ex.main(feature_set=feature_set, multiplier=multiplier)
    _barrier()
    if train and is_main():
        print("\n--- Step 3: Training and Evaluating Classifier ---")
        from .train_classifier import train_and_evaluate_classifier
        train_and_evaluate_classifier(readout=readout)
    if is_main():
        print("\n--- Pipeline Finished ---")
        if show_timing:
            print("\nWall-clock breakdown:\n" + timing.report())


def _barrier():
    from .distributed import _dist
    d = _dist()
    if d is not None:
        d.barrier()


def _cli(argv=None):
    parser = argparse.ArgumentParser(description="Run the entire speech recognition pipeline.")
    parser.add_argument("--n-filters", type=int, default=128, help="Number of filters for the filterbank.")
    parser.add_argument("--filterbank", type=str, default="gammatone", choices=["mel", "gammatone"],
                        help="Type of filterbank to use.")
    parser.add_argument("--feature-set", type=str, default="original",
                        choices=['all', 'rate', 'timing', 'rhythm', 'original'], help="The set of features to extract.")
    parser.add_argument("--multiplier", type=float, default=0.6, help="Multiplier for w_critico.")
    parser.add_argument("--synthetic", type=int, nargs=2, metavar=("CLASSES", "PER_CLASS"), default=None,
                        help="(extension) synthetic utterances instead of speech_commands_v0.02/")
    parser.add_argument("--no-train", action="store_true", help="(extension) stop after the feature file")
    parser.add_argument("--readout", type=str, default="sklearn", choices=["sklearn", "device"],
                        help="(extension) fit the logistic-regression readout with scikit-learn (reference) or on the GPU")
    parser.add_argument("--fused", action="store_true",
                        help="(extension) audio -> features in one pass, no spike file between the stages (same feature file)")
    parser.add_argument("--packed", action="store_true", help="(extension) bit-packed spike file between the stages")
    parser.add_argument("--timing", action="store_true", help="(extension) print a wall-clock breakdown of the stages")
    args = parser.parse_args(argv)
    run_pipeline(n_filters=args.n_filters, filterbank=args.filterbank, feature_set=args.feature_set,
                 multiplier=args.multiplier, synthetic=args.synthetic, train=not args.no_train, readout=args.readout,
                 fused=args.fused, packed=args.packed, show_timing=args.timing)


if __name__ == "__main__":
    _cli()
