"""GPU, more than one device: the utterance-sharded run is bit-identical for every world size.

Launches `python -m torch.distributed.run --nproc-per-node G tests/multirank_worker.py` for G in {2, 4, 8} (as many as the box has)
and compares the all-gathered feature matrix of every G with the single-GPU matrix computed here.  Skipped on a one-GPU box
(the driver's round-end GPU tier); the N = 8 run is recorded in profiles/r2_multirank.txt."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_features_identical_for_every_world_size(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("one GPU: nothing to shard across")
    from lsm_speech_classifier_b200 import synth
    from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm
    from lsm_speech_classifier_b200.frontend import Frontend
    from lsm_speech_classifier_b200.snn import AudioToFeatures
    pcm, _ = synth.synth_dataset(12, 25, workers=1)                 # 300 utterances: ragged blocks for G = 8 (38, 38, ..., 34)
    np.save(tmp_path / "pcm.npy", pcm)
    fe = Frontend(128, "gammatone")
    X = fe.encode(pcm)
    lsm = build_lsm(X, 0.6, verbose=False)
    keys = FEATURE_SETS["original"]
    want = AudioToFeatures(fe, lsm).run_host(pcm, keys)
    np.save(tmp_path / "want.npy", want)
    np.save(tmp_path / "spikes.npy", X)
    for G in [g for g in (2, 4, 8) if g <= n_gpu]:
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={G}",
                              "--master-addr", "127.0.0.1", "--master-port", str(29600 + G),
                              os.path.join(ROOT, "tests", "multirank_worker.py"), str(tmp_path)],
                             capture_output=True, text=True, timeout=600)
        print(out.stdout[-2000:])
        assert out.returncode == 0, out.stderr[-3000:]
        assert f"G={G} identical: spikes True features True fused-gather True" in out.stdout
