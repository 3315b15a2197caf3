"""Worker of tests/test_gpu_multirank.py (one rank per GPU under torchrun): stage 1 and stage 2 through the reference-named sharded
calls (NCCL all-gather) and the fused audio -> features path with the all-gather fused into the readout epilogue; rank 0 compares
everything with the single-GPU results the test wrote."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lsm_speech_classifier_b200.distributed import PeerAllGather, init_from_env, shard_bounds, sharded_features, sharded_spikes  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

d = sys.argv[1]
rank, world, local_rank = init_from_env()
pcm, want, want_spk = np.load(os.path.join(d, "pcm.npy")), np.load(os.path.join(d, "want.npy")), np.load(os.path.join(d, "spikes.npy"))
keys = FEATURE_SETS["original"]
fe = Frontend(128, "gammatone")
X = sharded_spikes(fe, pcm)                                   # create_dataset's path: every rank ends up with all spike trains
lsm = build_lsm(X, 0.6, verbose=False)                        # the same reservoir on every rank
feats = sharded_features(lsm, X, keys)                        # extract_all_features' path: NCCL all-gather of device rows
# the fused path with the all-gather in the kernel's epilogue (what bench.py times)
lo, hi, per = shard_bounds(len(pcm), rank, world)
pag = PeerAllGather(per, want.shape[1], torch.float64, torch.device("cuda", local_rank), n_buffers=1, ctx=fe.ctx)
pag.bufs[0].zero_()
dist.barrier()
lsm.set_gather(pag.pointers(0), rank * per)
AudioToFeatures(fe, lsm).run(torch.from_numpy(pcm[lo:hi]).cuda(), keys, want_spikes=False)
torch.cuda.synchronize()
lsm.set_gather([], 0)
dist.barrier()
fused = pag.bufs[0].cpu().numpy()[:len(pcm)]
ok = torch.tensor([float(np.array_equal(X, want_spk)), float(np.array_equal(feats, want)), float(np.array_equal(fused, want))], device="cuda")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"G={world} identical: spikes {bool(ok[0])} features {bool(ok[1])} fused-gather {bool(ok[2])}")
pag.close()
dist.destroy_process_group()
sys.exit(0 if bool(ok.min()) else 1)
