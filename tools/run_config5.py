"""BASELINE.json config 5: a Speech-Commands-scale set (105 000 synthetic utterances, 35 classes x 3000) sharded in contiguous
blocks of the class-major sample order across the ranks of one box, audio -> raw LSM features on every rank, one all-gather of
the feature rows (NCCL), and a bit-exact spot check of other ranks' rows on rank 0.  Markdown report on rank 0's stdout.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config5.py [--total 105000]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsm_speech_classifier_b200 import synth  # noqa: E402
from lsm_speech_classifier_b200.distributed import shard_bounds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--total", type=int, default=105000)
ap.add_argument("--classes", type=int, default=35)
ap.add_argument("--nccl", action="store_true", help="round-1 form: zero-copy feed, one NCCL all-gather after the kernels")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
per_class = args.total // args.classes
S = per_class * args.classes
lo, hi, per = shard_bounds(S, rank, world)


def jobs(a, b):
    return [(i // per_class, i % per_class) for i in range(a, b)]          # class-major global order


# synthesise this rank's block before CUDA is initialised (fork pool)
t0 = time.time()
workers = max(1, (os.cpu_count() or 8) // world)
import multiprocessing as mp  # noqa: E402

pcm = np.empty((hi - lo, synth.N_SAMPLES), dtype=np.float32)
with mp.get_context("fork").Pool(workers) as pool:
    for i, w in enumerate(pool.imap(synth._synth_one, jobs(lo, hi), chunksize=32)):
        pcm[i] = w
t_synth = time.time() - t0

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

from lsm_speech_classifier_b200.extract_lsm_features import FEATURE_SETS, build_lsm  # noqa: E402
from lsm_speech_classifier_b200.frontend import Frontend  # noqa: E402
from lsm_speech_classifier_b200.snn import AudioToFeatures  # noqa: E402

keys = FEATURE_SETS["original"]
fe = Frontend(128, "gammatone")
# w_critico from the first <= 500 utterances of the global order: every rank computes the same head redundantly
head = np.stack([synth.synth_utterance(*j) for j in jobs(0, 64)])
lsm = build_lsm(fe.encode(head), 0.6, verbose=False)
path = AudioToFeatures(fe, lsm)
F = len(keys) * lsm.num_output_neurons

h_pcm = torch.from_numpy(pcm).pin_memory()
d_local = torch.zeros((per, F), dtype=torch.float64, device="cuda")
path.run_host(h_pcm[:256].numpy(), keys, out=d_local[:256])            # warm-up
torch.cuda.synchronize()
fused = world > 1 and not args.nccl
if fused:
    # round-2 form: PCM by the copy engine in batches of 2400 on alternating launch lanes, and the all-gather fused into the
    # readout epilogue (every rank's kernel stores its rows into all ranks' matrices over NVLink)
    from lsm_speech_classifier_b200.distributed import PeerAllGather
    pag = PeerAllGather(per, F, torch.float64, torch.device("cuda", local_rank), n_buffers=1, ctx=fe.ctx)
    pag.bufs[0].zero_()
    fe.ctx.set_host_feed("copy_engine")
    lsm.set_gather(pag.pointers(0), rank * per)
    w = min(2400, hi - lo)
    for ln in (0, 1):                                                  # staging buffers of this feed, both lanes, full batch size
        path.run_host_async(h_pcm[:w], keys, out=d_local[:w], lane=ln)
    fe.ctx.sync_all()
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
if fused:
    step = 2400
    for k, off in enumerate(range(0, hi - lo, step)):
        n = min(step, hi - lo - off)
        lsm.set_gather(pag.pointers(0), rank * per + off)
        path.run_host_async(h_pcm[off:off + n], keys, out=d_local[off:off + n], lane=k & 1)
    fe.ctx.sync_all()
else:
    path.run_host(h_pcm.numpy(), keys, out=d_local[:hi - lo])          # pinned PCM in (zero-copy), feature rows stay on the device
torch.cuda.synchronize()
t_compute = time.perf_counter() - t0
if fused:
    lsm.set_gather([], 0)
    dist.barrier()                                                     # every rank's stores have landed
    d_all = pag.bufs[0]
    t_gather = 0.0
else:
    d_all = torch.empty((world * per, F), dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world > 1:
        dist.all_gather_into_tensor(d_all, d_local)
    else:
        d_all.copy_(d_local)
    torch.cuda.synchronize()
    t_gather = time.perf_counter() - t0
times = torch.tensor([t_synth, t_compute, t_gather], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
t_synth, t_compute, t_gather = [float(x) for x in times]

if rank == 0:
    # spot check: recompute 16 utterances out of every rank's block here and compare with the gathered rows, bit for bit
    ok, checked = True, 0
    for r in range(world):
        rlo, rhi, _ = shard_bounds(S, r, world)
        pick = [rlo + k * max(1, (rhi - rlo) // 16) for k in range(16) if rlo + k * max(1, (rhi - rlo) // 16) < rhi]
        sub = np.stack([synth.synth_utterance(*j) for j in [(i // per_class, i % per_class) for i in pick]])
        want = path.run_host(sub, keys)
        got = d_all[[r * per + (i - rlo) for i in pick]].cpu().numpy()
        ok = ok and np.array_equal(got, want)
        checked += len(pick)
    gb = d_all.numel() * 8 / 1e9
    print(f"## Config 5 - {S} synthetic utterances ({args.classes} classes x {per_class}), {world} x B200, contiguous blocks of {per} utterances per rank\n")
    print(f"* synthesis on the host (not part of the path): {t_synth:.1f} s per rank, {workers} worker processes each")
    print(f"* audio -> raw features, pinned host PCM in, feature rows in device memory: **{t_compute * 1e3:.1f} ms** for the slowest rank "
          f"= **{S / t_compute / 1e6:.2f} M utterances/s** over the box ({S / t_compute / world / 1e3:.0f} k per GPU)")
    if fused:
        print(f"* the all-gather of the float64[{per}, {F}] blocks -> float64[{world * per}, {F}] ({gb:.2f} GB on every rank) is inside that time: the "
              f"readout epilogue of every rank's kernel stores its rows into all {world} matrices over NVLink (PCM by the copy engine in batches of 2400)")
    else:
        print(f"* one NCCL all-gather of the float64[{per}, {F}] blocks -> float64[{world * per}, {F}] ({gb:.2f} GB on every rank): "
              f"{t_gather * 1e3:.1f} ms ({gb / max(t_gather, 1e-9):.0f} GB/s into each GPU)")
    print(f"* spot check on rank 0: {checked} rows out of all {world} blocks recomputed locally and compared with the gathered matrix: "
          f"bit-identical **{ok}**; exact re-executions of the speculative filter on this rank: {fe.reruns()}")
if world > 1:
    dist.destroy_process_group()
