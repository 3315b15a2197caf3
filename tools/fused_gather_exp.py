"""Experiment (torchrun, >= 2 GPUs): the feature all-gather fused into the readout epilogue (SNN.set_gather) with the peers'
gather matrices mapped through CUDA IPC on the local device, checked against NCCL's all-gather and timed beside it."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
B = 2400
base, _ = synth.synth_dataset(12, 20, start_utt=20 * rank, workers=2)
pcm = np.concatenate([base] * (B // len(base) + 1))[:B]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
from lsm_speech_classifier_b200.distributed import PeerAllGather  # noqa: E402
from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

keys = FEATURE_SETS["original"]
d_pcm = torch.from_numpy(pcm).cuda()
fe = Frontend(128, "gammatone")
lsm = build_lsm(fe.encode(d_pcm[:240]).cpu().numpy(), 0.6, verbose=False)
path = AudioToFeatures(fe, lsm)
pag = PeerAllGather(B, 2000, torch.float64, torch.device("cuda", local_rank), map_on_local_device=True)
d_feat = [torch.empty((B, 2000), dtype=torch.float64, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def steps(n, fused):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    pend = [None, None]
    for i in range(n):
        k = i & 1
        with torch.cuda.stream(streams[k]):
            if i < 2:
                streams[k].wait_event(a)
            if pend[k] is not None:
                pend[k].wait()
            if fused:
                lsm.set_gather(pag.pointers(k), rank * B)
            path.run(d_pcm, keys, out=d_feat[k], want_spikes=False)
            if not fused:
                pend[k] = dist.all_gather_into_tensor(nccl_all[k], d_feat[k], async_op=True)
    for k in (0, 1):
        with torch.cuda.stream(streams[k]):
            if pend[k] is not None:
                pend[k].wait()
        torch.cuda.current_stream().wait_stream(streams[k])
    b.record()
    torch.cuda.synchronize()
    lsm.set_gather([], 0)
    dist.barrier()
    t = torch.tensor([a.elapsed_time(b) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


nccl_all = [torch.empty((world * B, 2000), dtype=torch.float64, device="cuda") for _ in range(2)]
steps(4, False); steps(4, True)
t_nccl = steps(10, False)
t_fused = steps(10, True)
torch.cuda.synchronize(); dist.barrier()
ok = all(torch.equal(pag.bufs[k], nccl_all[k]) for k in (0, 1))
if rank == 0:
    print(f"N={world}: NCCL all-gather {t_nccl:.3f} ms/step, fused gather {t_fused:.3f} ms/step, matrices identical: {ok}")
dist.destroy_process_group()
