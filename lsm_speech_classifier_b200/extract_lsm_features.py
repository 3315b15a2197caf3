"""Stage 2 of the pipeline: spike-train dataset -> LSM features -> lsm_features_larger.npz.

Same call surface as /root/reference/extract_lsm_features.py (constants :10-16, FEATURE_SETS :19-28,
calculate_theoretical_w_critico :33-60, load_spike_dataset :63-73, extract_all_features :76-89,
run_network_diagnostics :92-152, main :155-214, flags :219-221), but each per-sample loop is one
batched kernel launch on the B200, and with --gpus/torchrun the samples shard across ranks with a
single all-gather of the raw feature rows.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np

from . import npzio, timing
from .snn import SNN, SimulationParams

NUM_NEURONS = 1000
NUM_OUTPUT_NEURONS = 400
LEAK_COEFFICIENT = 1 / 100
REFRACTORY_PERIOD = 2
MEMBRANE_THRESHOLD = 2.0
SMALL_WORLD_P = 0.1
SMALL_WORLD_K = int(0.10 * NUM_NEURONS * 2)

FEATURE_SETS = {
    'all': ['spike_counts', 'spike_variances', 'mean_spike_times', 'first_spike_times',
            'last_spike_times', 'mean_isi', 'isi_variances', 'burst_counts'],
    'rate': ['spike_counts', 'spike_variances', 'burst_counts'],
    'timing': ['mean_spike_times', 'first_spike_times', 'last_spike_times'],
    'rhythm': ['mean_isi', 'isi_variances'],
    'original': ['spike_counts', 'spike_variances', 'mean_spike_times', 'mean_isi', 'isi_variances'],
}

SPIKE_FILE = "speech_spike_dataset_pure_redundancy.npz"
FEATURE_FILE = "lsm_features_larger.npz"

np.random.seed(42)


def calculate_theoretical_w_critico(lsm_params, input_data, verbose: bool = True):
    """Mean-field critical weight (reference :33-60): w = (theta - 2*avg_I*R) / (k/2) with avg_I the
    input spike density over the first <= 500 samples.  Exact integer sums; 0.007 fallbacks kept."""
    num_samples = min(500, len(input_data))
    head = np.asarray(input_data[:num_samples])
    total_spikes = int(head.sum(dtype=np.int64)) if head.size else 0
    total_elements = int(head.size)
    if total_elements == 0:
        return 0.007
    avg_I = total_spikes / total_elements
    beta = lsm_params.small_world_graph_k / 2
    if beta == 0:
        return 0.007
    w_critico = (lsm_params.membrane_threshold - 2 * avg_I * lsm_params.refractory_period) / beta
    if verbose:
        print(f"Theoretical w_critico: {w_critico:.8f}")
    return w_critico


PACKED_SPIKE_FILE = "speech_spike_dataset_packed.npz"


def load_spike_dataset(filename=SPIKE_FILE):
    if not Path(filename).exists():
        if filename == SPIKE_FILE and Path(PACKED_SPIKE_FILE).exists():
            # (extension) the bit-packed form written by create_dataset --packed: same arrays after unpacking
            from .create_dataset import load_packed_spikes
            X_spikes, y_labels = load_packed_spikes(PACKED_SPIKE_FILE)
            print(f"Loaded {len(X_spikes)} samples from '{PACKED_SPIKE_FILE}' (bit-packed)")
            return X_spikes, y_labels
        print(f"Error: Dataset not found at '{filename}'")
        return None, None
    data = np.load(filename)
    X_spikes, y_labels = data['X_spikes'], data['y_labels']
    print(f"Loaded {len(X_spikes)} samples from '{filename}'")
    return X_spikes, y_labels


def extract_all_features(lsm: SNN, spike_data, feature_keys, desc: str = ""):
    """reference :76-89 for the whole array at once -> float64[S, len(keys)*N_out], NaNs zeroed."""
    spike_data = np.asarray(spike_data, dtype=np.uint8)
    if len(spike_data) == 0:
        return np.zeros((0, len(feature_keys) * lsm.num_output_neurons))
    from .distributed import sharded_features
    return sharded_features(lsm, spike_data, list(feature_keys))


def run_network_diagnostics(lsm: SNN, X_sample_batch):
    """reference :92-152: participation / dead neurons / mean activity on the first 5 samples."""
    print("\n" + "=" * 40)
    print("RUNNING NETWORK DIAGNOSTICS")
    print("=" * 40)
    subset = np.asarray(X_sample_batch[:5], dtype=np.uint8)
    if len(subset) == 0:
        return None
    # reduced in the kernel epilogue; no raster leaves the device
    participation, dead, avg_spikes = lsm.diagnostics(subset)
    rates = []
    for i in range(len(subset)):
        rates.append(float(participation[i]))
        print(f"Sample {i + 1}: Active: {participation[i]:.1f}% | Dead: {int(dead[i])} | Avg Spikes/Neuron: {avg_spikes[i]:.2f}")
    avg_part = float(np.mean(rates))
    print("-" * 40)
    print("DIAGNOSTIC RESULT:")
    print(f"   Average Participation: {avg_part:.1f}%")
    if avg_part < 40:
        print("   STATUS: SUB-CRITICAL (Too Silent)")
        print("   Recommendation: INCREASE multiplier or DECREASE threshold.")
    elif avg_part > 98:
        print("   STATUS: SUPER-CRITICAL (Epileptic/Saturated)")
        print("   Recommendation: DECREASE multiplier.")
    else:
        print("   STATUS: EDGE OF CHAOS (Healthy)")
        print("   (Ideal is 80-95% participation with low firing rates)")
    print("=" * 40 + "\n")
    return avg_part


def build_lsm(X_train, multiplier: float, leak_variance_divisor=None, num_neurons: int = NUM_NEURONS, verbose=True,
              strict_weights: bool = False) -> SNN:
    """reference :164-188: parameter record, w_critico, weight, then the one reservoir.  strict_weights: fp64 weights summed in
    ascending presynaptic order (SURVEY.md 8c S3/S6 as written) instead of the order-free 2^-24 quantisation."""
    k = int(0.10 * num_neurons * 2)
    base_params = SimulationParams(
        num_neurons=num_neurons, mean_weight=0.0, num_output_neurons=NUM_OUTPUT_NEURONS,
        membrane_threshold=MEMBRANE_THRESHOLD, leak_coefficient=LEAK_COEFFICIENT,
        refractory_period=REFRACTORY_PERIOD, small_world_graph_p=SMALL_WORLD_P, small_world_graph_k=k,
        input_spike_times=X_train[0], leak_variance_divisor=leak_variance_divisor, quantize_weights=not strict_weights)
    w = calculate_theoretical_w_critico(base_params, X_train, verbose=verbose)
    optimal_weight = w * multiplier
    if verbose:
        print(f"Using weight: {optimal_weight:.8f} (multiplier: {multiplier:.2f})")
        if leak_variance_divisor:
            print(f"Using Heterogeneous Leak. Divisor: {leak_variance_divisor}")
    base_params.mean_weight = optimal_weight
    base_params.weight_variance = 10
    return SNN(simulation_params=base_params)


def main(feature_set: str, multiplier: float, leak_variance_divisor: float = None, num_neurons: int = NUM_NEURONS,
         strict_weights: bool = False):
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import StandardScaler
    from .distributed import _dist, init_from_env, is_main

    init_from_env()             # under torchrun (also when this stage is run on its own): one rank per GPU, samples sharded
    if _dist() is not None:
        _dist().barrier()       # the spike file is written by rank 0 of the previous stage
    with timing.stage("stage 2 read: spike file (inflate, one core)"):
        X_spikes, y_labels = load_spike_dataset()
    if X_spikes is None:
        return
    with timing.stage("stage 2 split (train_test_split copies)"):
        X_train, X_test, y_train, y_test = train_test_split(
            X_spikes, y_labels, test_size=0.2, random_state=42, stratify=y_labels)
    with timing.stage("stage 2 setup: w_critico, reservoir build + upload, diagnostics"):
        lsm = build_lsm(X_train, multiplier, leak_variance_divisor, num_neurons, verbose=is_main(), strict_weights=strict_weights)
        if is_main():
            run_network_diagnostics(lsm, X_train)
    feature_keys = FEATURE_SETS[feature_set]
    if is_main():
        print(f"Extracting feature set: '{feature_set}'")
    if timing.enabled:
        with timing.stage("stage 2 GPU init (one-time when run on its own): staging"):
            lsm.simulate_batch(np.zeros((min(len(X_train), 1024),) + X_train.shape[1:], np.uint8), feature_keys)
    with timing.stage("stage 2 compute: spike trains -> features (GPU, host buffers)", len(X_train) + len(X_test)):
        X_train_feat = extract_all_features(lsm, X_train, feature_keys, "Training")
        X_test_feat = extract_all_features(lsm, X_test, feature_keys, "Testing")
    if not is_main():
        return
    with timing.stage("stage 2 scaler (sklearn, host)"):
        scaler = StandardScaler()
        X_train_scaled = scaler.fit_transform(X_train_feat)
        X_test_scaled = scaler.transform(X_test_feat)
    with timing.stage("stage 2 write: feature file (parallel deflate)"):
        npzio.savez_compressed(FEATURE_FILE, X_train_features=X_train_scaled, y_train=y_train,
                               X_test_features=X_test_scaled, y_test=y_test, feature_set=feature_set,
                               leak_variance_divisor=leak_variance_divisor)
    print(f"Extraction complete. Features saved to '{FEATURE_FILE}'")


def main_fused(pcm, y_labels, n_filters: int, filterbank: str, feature_set: str, multiplier: float,
               leak_variance_divisor: float = None, num_neurons: int = NUM_NEURONS):
    """(extension) stages 1 and 2 without the spike file in between: audio -> features through the fused
    path, same split, same reservoir, same feature file as create_dataset + main (bit for bit; tests/test_gpu_cli.py).  Only
    the first <= 500 training utterances are encoded on their own, for w_critico (reference :40-49, :160-162)."""
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import StandardScaler
    from .distributed import world
    from .frontend import Frontend
    from .snn import AudioToFeatures
    if world()[1] > 1:
        raise RuntimeError("main_fused is single-process; under torchrun use create_dataset + main (utterance-sharded)")
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    y_labels = np.asarray(y_labels, dtype=np.int32)
    idx = np.arange(len(pcm))
    tr, te, y_train, y_test = train_test_split(idx, y_labels, test_size=0.2, random_state=42, stratify=y_labels)
    with timing.stage("fused setup: w_critico head, reservoir build + upload, diagnostics"):
        fe = Frontend(n_filters, filterbank)
        head = fe.encode(pcm[tr[:500]])
        lsm = build_lsm(head, multiplier, leak_variance_divisor, num_neurons)
        run_network_diagnostics(lsm, head)
    feature_keys = FEATURE_SETS[feature_set]
    print(f"Extracting feature set: '{feature_set}' (fused audio -> features)")
    path = AudioToFeatures(fe, lsm)
    if timing.enabled:
        with timing.stage("GPU init (one-time): pinned staging ring, host copy threads"):
            path.run_host(np.zeros((min(len(pcm), 1024), pcm.shape[1]), np.float32), feature_keys)
    with timing.stage("fused gather of the split rows (numpy fancy indexing)"):
        pcm_tr, pcm_te = np.ascontiguousarray(pcm[tr]), np.ascontiguousarray(pcm[te])
    with timing.stage("fused compute: audio -> features (GPU, host buffers)", len(pcm)):
        X_train_feat = path.run_host(pcm_tr, feature_keys)
        X_test_feat = path.run_host(pcm_te, feature_keys)
    with timing.stage("scaler (sklearn, host)"):
        scaler = StandardScaler()
        X_train_scaled = scaler.fit_transform(X_train_feat)
        X_test_scaled = scaler.transform(X_test_feat)
    with timing.stage("write: feature file (parallel deflate)"):
        npzio.savez_compressed(FEATURE_FILE, X_train_features=X_train_scaled, y_train=y_train,
                               X_test_features=X_test_scaled, y_test=y_test, feature_set=feature_set,
                               leak_variance_divisor=leak_variance_divisor)
    print(f"Extraction complete. Features saved to '{FEATURE_FILE}'")


def _cli(argv=None):
    parser = argparse.ArgumentParser(description="Extract features from a spike train dataset using an LSM.")
    parser.add_argument("--feature-set", type=str, default="original", choices=FEATURE_SETS.keys())
    parser.add_argument("--multiplier", type=float, default=0.6)
    parser.add_argument("--leak-variance-divisor", type=float, default=None)
    parser.add_argument("--n-neurons", type=int, default=NUM_NEURONS, help="(extension) reservoir size; k = 0.2*N")
    parser.add_argument("--strict-weights", action="store_true",
                        help="(extension) fp64 weights summed in ascending presynaptic order instead of order-free 2^-24 multiples")
    args = parser.parse_args(argv)
    main(feature_set=args.feature_set, multiplier=args.multiplier, leak_variance_divisor=args.leak_variance_divisor,
         num_neurons=args.n_neurons, strict_weights=args.strict_weights)


if __name__ == "__main__":
    _cli()
