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
    default_compute = compute is None
    if compute is None:
        def compute(block):
            return lsm.simulate_batch(block, feature_keys, nan_to_num=True)
    width = len(feature_keys) * lsm.num_output_neurons
    d = _dist()
    if default_compute and ws > 1 and d.get_backend() == "nccl":
        # GPU ranks: the block goes to the device once, the feature rows stay there for the all-gather, and only the gathered
        # matrix comes back (no numpy round trip between the kernel and the collective)
        import torch
        dev = torch.device("cuda", lsm.ctx.device)
        out = torch.zeros((ws * per, width), dtype=torch.float64, device=dev)
        mine = out[rank * per: rank * per + (hi - lo)]
        if hi > lo:
            block = torch.from_numpy(np.ascontiguousarray(spike_data[lo:hi], dtype=np.uint8)).to(dev)
            mine.copy_(lsm.simulate_batch(block, feature_keys, nan_to_num=True))
        d.all_gather_into_tensor(out, out[rank * per: (rank + 1) * per].clone())
        return out.cpu().numpy()[:n]
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


class _DeviceBlock:
    """A device allocation owned by the C library, viewable as a torch tensor (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerAllGather:
    """The feature all-gather without a collective kernel.

    The compute kernels of this path are persistent and fill every SM, so an NCCL all-gather kernel has to find CTA slots
    between them (round 1: 7.0 vs 5.8 ms per step at N = 8).  Here every rank owns `n_buffers` gather matrices allocated by
    the C library (lsm_peer_buffer_create), exports them as CUDA IPC handles, and maps the other ranks' matrices into its own
    process with its own device current (lsm_peer_buffer_open).  `pointers(k)` is then what `SNN.set_gather` takes: the
    readout epilogue of the kernel stores every feature row into all ranks' matrices itself, as NVLink stores.
    `gather_async` is the copy-engine form of the same collective (cudaMemcpyPeerAsync through the same mappings).
    The caller synchronises ranks (a barrier) before it reads `bufs[k]`."""

    def __init__(self, rows: int, width: int, dtype, device, n_buffers: int = 2, ctx=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        if dtype != torch.float64:
            raise ValueError("PeerAllGather holds float64 feature rows")
        self.ctx = ctx or _lib.context(torch.device(device).index)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.rows, self.width = rows, width
        nbytes = self.world * rows * width * 8
        # Every rank walks through the same collectives whatever fails locally; the outcome is agreed on at the end, so that
        # either all ranks hold a working mapping or all of them raise.
        self._own, self._opened, handles, err = [], [], [], None
        try:
            for _ in range(n_buffers):
                ptr, h = C.c_void_p(), C.create_string_buffer(64)
                self.ctx.check(self.ctx.lib.lsm_peer_buffer_create(self.ctx.h, nbytes, C.byref(ptr), h))
                self._own.append(int(ptr.value))
                handles.append(h.raw)
        except Exception as e:
            err, handles = e, None
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles)
        self._ptr = [[0] * n_buffers for _ in range(self.world)]
        if err is None and all(h is not None for h in everyone):
            try:
                for r in range(self.world):
                    for k in range(n_buffers):
                        if r == self.rank:
                            self._ptr[r][k] = self._own[k]
                        else:
                            ptr = C.c_void_p()
                            self.ctx.check(self.ctx.lib.lsm_peer_buffer_open(self.ctx.h, C.c_char_p(everyone[r][k]), C.byref(ptr)))
                            self._ptr[r][k] = int(ptr.value)
                            self._opened.append(int(ptr.value))
            except Exception as e:
                err = e
        elif err is None:
            err = RuntimeError("another rank could not create its gather buffers")
        ok = torch.tensor([0.0 if err is not None else 1.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # doubles as the barrier: every rank has mapped every buffer
        if ok.item() < 1.0:
            for p in self._opened:
                self.ctx.lib.lsm_peer_buffer_close(self.ctx.h, p)
            dist.barrier()
            for p in self._own:
                self.ctx.lib.lsm_peer_buffer_destroy(self.ctx.h, p)
            self._own = None
            raise RuntimeError(f"PeerAllGather: peer mapping unavailable on some rank ({err})")
        self.bufs = [torch.as_tensor(_DeviceBlock(p, (self.world * rows, width), "<f8"), device=device) for p in self._own]
        self.peer = [[(self.bufs[k] if r == self.rank else
                       torch.as_tensor(_DeviceBlock(self._ptr[r][k], (self.world * rows, width), "<f8"), device=device))
                      for k in range(n_buffers)] for r in range(self.world)]
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
        return [self._ptr[r][k] for r in range(self.world)]

    def wait(self, k: int, stream=None):
        """Make `stream` (default: the current one) wait for this rank's outstanding copies into buffers k."""
        import torch
        (stream or torch.cuda.current_stream()).wait_stream(self.streams[k])

    def close(self):
        """Unmap the peers' matrices and free this rank's own (collective: every rank calls it)."""
        import torch.distributed as dist
        if getattr(self, "_own", None) is None:
            return
        dist.barrier()
        for p in self._opened:
            self.ctx.lib.lsm_peer_buffer_close(self.ctx.h, p)
        dist.barrier()
        for p in self._own:
            self.ctx.lib.lsm_peer_buffer_destroy(self.ctx.h, p)
        self._own = None
        self.bufs = self.peer = []
