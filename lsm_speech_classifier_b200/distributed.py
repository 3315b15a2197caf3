"""Utterance-sharded data parallelism: one process per GPU, no data-path collective except one
all-gather of the raw feature rows.

Every utterance is independent in both stages (/root/reference/create_dataset.py:143-162,
extract_lsm_features.py:78-87 with reset() per sample at :79) and the reservoir is read-only
(:188), so ranks take contiguous blocks of the sample order, padded to equal length, and the
result is bit-identical for any world size.  Works under torchrun with NCCL on GPUs and, for the
CPU test-suite, with gloo (the compute callback is injected there).
"""
from __future__ import annotations

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def is_main() -> bool:
    return world()[0] == 0


def shard_bounds(n: int, rank: int, world_size: int):
    """Contiguous block [lo, hi) of rank `rank`; blocks are ceil(n/world) long, the tail ranks short/empty."""
    per = -(-n // world_size) if n else 0
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def all_gather_rows(local: np.ndarray, n_total: int, per: int, device=None) -> np.ndarray:
    """All-gather equal-length (padded) row blocks and trim to n_total rows.  NCCL if the process
    group is NCCL (rows go through the GPU), gloo otherwise."""
    d = _dist()
    if d is None:
        return local[:n_total]
    import torch
    rank, ws = world()
    backend = d.get_backend()
    width = local.shape[1]
    pad = np.zeros((per, width), dtype=local.dtype)
    pad[:len(local)] = local
    t = torch.from_numpy(pad)
    if backend == "nccl":
        t = t.cuda(device if device is not None else torch.cuda.current_device())
    out = torch.empty((ws * per, width), dtype=t.dtype, device=t.device)
    d.all_gather_into_tensor(out, t)
    return out.cpu().numpy()[:n_total]


def sharded_features(lsm, spike_data: np.ndarray, feature_keys, compute=None) -> np.ndarray:
    """Each rank simulates its block; every rank returns the full [S, F] matrix in sample order."""
    rank, ws = world()
    n = len(spike_data)
    lo, hi, per = shard_bounds(n, rank, ws)
    if compute is None:
        def compute(block):
            return lsm.simulate_batch(block, feature_keys, nan_to_num=True)
    width = len(feature_keys) * lsm.num_output_neurons
    local = compute(spike_data[lo:hi]) if hi > lo else np.zeros((0, width))
    local = np.ascontiguousarray(local, dtype=np.float64)
    if ws == 1:
        return local
    return all_gather_rows(local, n, per, getattr(getattr(lsm, "ctx", None), "device", None))


def sharded_spikes(frontend, pcm: np.ndarray, compute=None) -> np.ndarray:
    """Stage 1 under the same sharding: uint8 spike trains for all S utterances on every rank."""
    rank, ws = world()
    n = len(pcm)
    lo, hi, per = shard_bounds(n, rank, ws)
    if compute is None:
        compute = frontend.encode
    rows, steps = frontend.rows, frontend.steps
    local = compute(pcm[lo:hi]) if hi > lo else np.zeros((0, rows, steps), np.uint8)
    if ws == 1:
        return local
    flat = all_gather_rows(np.ascontiguousarray(local).reshape(len(local), rows * steps), n, per,
                           getattr(getattr(frontend, "ctx", None), "device", None))
    return flat.reshape(n, rows, steps)


def init_from_env():
    """Initialise torch.distributed from torchrun's environment (no-op without it). Returns (rank, world, local_rank)."""
    import os
    if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return 0, 1, int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group("gloo")
    return dist.get_rank(), dist.get_world_size(), local_rank


class PeerAllGather:
    """All-gather of equal row blocks by peer-to-peer copies over NVLink (copy engines, no SM use).

    The compute kernels of this path are persistent and fill every SM, so an NCCL all-gather kernel has to find CTA slots
    between them.  Here every rank maps the gather buffers of all ranks (CUDA IPC, exchanged once through the process
    group) and writes its block straight into each of them with cudaMemcpyPeerAsync on a side stream; the SMs never see
    the collective.  Measured at N = 2: 6.68 ms per step against 6.72 ms with NCCL's asynchronous all-gather - the
    collective was not what separates N = 2 from N = 1 (6.3 ms; the step time is the maximum over ranks).  At N = 8 this
    simple form is far worse than NCCL (25.5 vs 7.0 ms per step: seven serial 38 MB peer copies per rank and step through
    IPC-mapped buffers), so bench.py keeps NCCL and uses this class only with LSM_BENCH_P2P=1; both give identical matrices
    (checked in bench.py).
    `n_buffers` gather buffers per rank alternate between steps; the caller synchronises ranks (a barrier) before it reads."""

    def __init__(self, rows: int, width: int, dtype, device, n_buffers: int = 2, map_on_local_device: bool = False):
        import torch
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.rows = rows
        self.bufs = [torch.empty((self.world * rows, width), dtype=dtype, device=device) for _ in range(n_buffers)]
        mine = [reduce_tensor(b) for b in self.bufs]
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        # peer[r][k]: rank r's gather buffer k, addressable from this process (this rank's own buffers are used directly).
        # map_on_local_device: open the IPC handles with THIS rank's device current (rebuild_cuda_tensor's storage_device
        # argument), so that the mapping belongs to the context kernels of this rank run in - what kernels that store into the
        # peers' buffers (SNN.set_gather) need; torch then labels those tensors with the local device.
        def rebuild(fn, args):
            if map_on_local_device:
                args = list(args)
                args[6] = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
                args = tuple(args)
            return fn(*args)
        self.peer = [[(self.bufs[k] if r == self.rank else rebuild(fn, args)) for k, (fn, args) in enumerate(everyone[r])]
                     for r in range(self.world)]
        self.streams = [torch.cuda.Stream(device=device) for _ in range(n_buffers)]

    def gather_async(self, k: int, local, after_stream):
        """Write `local` (this rank's [rows, width] block) into slot `rank` of every rank's buffer k, ordered after the
        work already enqueued on `after_stream`.  Returns the side stream the copies run on."""
        import torch
        st = self.streams[k]
        st.wait_stream(after_stream)
        lo = self.rank * self.rows
        with torch.cuda.stream(st):
            for d in range(self.world):
                r = (self.rank + d) % self.world              # start with the own buffer, then round the ring
                self.peer[r][k][lo:lo + self.rows].copy_(local, non_blocking=True)
        return st

    def pointers(self, k: int):
        """Device addresses of every rank's gather buffer k (rank order), for SNN.set_gather: the fused all-gather, where the
        readout epilogue of the kernel stores the rows into all of them itself."""
        return [self.peer[r][k].data_ptr() for r in range(self.world)]

    def wait(self, k: int, stream=None):
        """Make `stream` (default: the current one) wait for this rank's outstanding copies into buffers k."""
        import torch
        (stream or torch.cuda.current_stream()).wait_stream(self.streams[k])
