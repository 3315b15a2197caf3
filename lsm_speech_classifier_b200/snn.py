"""Stage 2+3 host side: the liquid state machine on the GPU (K2 + K3).

`SNN` honours the object protocol the reference uses from the un-vendored ``snnpy.snn``
(/root/reference/extract_lsm_features.py):

    lsm = SNN(simulation_params=base_params)                  # :188
    lsm.reset(); lsm.set_input_spike_times(sample); lsm.simulate()          # :79-81, :109-111
    feature_dict = lsm.extract_features_from_spikes()         # :83
    lsm.num_neurons, lsm.spike_matrix (Time x Neurons)        # :100, :115-116

so the reference's `extract_all_features` / `run_network_diagnostics` loops run against it
unchanged (tests/test_gpu_parity.py does exactly that), plus the batched calls the fast path uses:

    lsm.simulate_batch(spikes[B,C,T], feature_keys) -> features[B,F]      one kernel for the whole batch
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .reservoir import ReservoirDef, SimulationParams, build_reservoir  # noqa: F401  (re-exported)

FEATURE_KEYS = _lib.FEATURE_KEYS


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _is_int16(x) -> bool:
    return str(x.dtype) in ("int16", "torch.int16")


class SNN:
    def __init__(self, simulation_params: SimulationParams | None = None, reservoir: ReservoirDef | None = None,
                 ctx: _lib.Context | None = None, device: int | None = None):
        if reservoir is None:
            if simulation_params is None:
                raise ValueError("SNN needs simulation_params or a prebuilt ReservoirDef")
            reservoir = build_reservoir(simulation_params)
        self.params = simulation_params
        self.reservoir = r = reservoir
        self.ctx = ctx or _lib.context(device)
        self.num_neurons = r.num_neurons
        self.num_output_neurons = len(r.out_idx)
        p = _lib.ReservoirParams()
        p.num_neurons, p.num_inputs, p.num_steps = r.num_neurons, r.num_inputs, r.num_steps
        p.refractory, p.w_shift, p.n_out, p.theta = r.refractory, r.w_shift, len(r.out_idx), r.theta
        self._p = p
        strict = getattr(r, "w_val", None) is not None        # fp64 weights, ordered sums (SimulationParams.quantize_weights=False)
        arrs = [_lib.as_host(r.w_rowptr, np.int32), _lib.as_host(r.w_col, np.int32),
                _lib.as_host(r.w_val, np.float64) if strict else _lib.as_host(r.w_q, np.int32),
                _lib.as_host(r.in_rowptr, np.int32), _lib.as_host(r.in_col, np.int32), _lib.as_host(r.in_val, np.float64),
                _lib.as_host(r.leak, np.float64), _lib.as_host(r.out_idx, np.int32)]
        h = C.c_void_p()
        create = self.ctx.lib.lsm_reservoir_create_f64 if strict else self.ctx.lib.lsm_reservoir_create
        self.ctx.check(create(self.ctx.h, C.byref(p), *[_lib._np_ptr(a) for a in arrs], C.byref(h)))
        self.h = h
        self._sample = None
        self.spike_matrix = None
        self._stats = None

    # ------------------------------------------------------------------ batched fast path
    def simulate_batch(self, spikes, feature_keys=None, nan_to_num: bool = True, return_raster: bool = False):
        """spikes uint8[B, C, T] -> raw (un-standardised) features float64[B, len(keys)*N_out],
        key-major like extract_lsm_features.py:85-87.  torch CUDA in -> torch CUDA out (async);
        numpy in -> numpy out."""
        keys = list(FEATURE_KEYS if feature_keys is None else feature_keys)
        mask = _lib.feature_mask(keys)
        order = _lib.mask_keys(mask)
        r = self.reservoir
        n_out = self.num_output_neurons
        if _is_torch(spikes):
            import torch
            if not spikes.is_cuda:
                raise _lib.LsmError("simulate_batch(torch tensor) needs a CUDA tensor; pass numpy for host buffers")
            spikes = spikes.contiguous()
            if spikes.dtype != torch.uint8 or spikes.dim() != 3 or tuple(spikes.shape[1:]) != (r.num_inputs, r.num_steps):
                raise ValueError(f"spikes must be uint8[B,{r.num_inputs},{r.num_steps}]")
            B = spikes.shape[0]
            feats = torch.empty((B, len(order) * n_out), dtype=torch.float64, device=spikes.device)
            raster = torch.empty((B, r.num_steps, r.num_neurons), dtype=torch.uint8, device=spikes.device) if return_raster else None
            self.ctx.set_stream(torch.cuda.current_stream(spikes.device).cuda_stream)
            self.ctx.check(self.ctx.lib.lsm_reservoir_run(
                self.ctx.h, self.h, C.c_void_p(spikes.data_ptr()), B, mask, int(nan_to_num), C.c_void_p(feats.data_ptr()),
                C.c_void_p(raster.data_ptr()) if raster is not None else None))
            feats = self._reorder(feats, order, keys, n_out)
            return (feats, raster) if return_raster else feats
        spikes = _lib.as_host(spikes, np.uint8)
        if spikes.ndim != 3 or spikes.shape[1:] != (r.num_inputs, r.num_steps):
            raise ValueError(f"spikes must be uint8[B,{r.num_inputs},{r.num_steps}]")
        B = spikes.shape[0]
        feats = np.empty((B, len(order) * n_out), dtype=np.float64)
        raster = np.empty((B, r.num_steps, r.num_neurons), dtype=np.uint8) if return_raster else None
        self.ctx.set_stream(None)
        self.ctx.check(self.ctx.lib.lsm_reservoir_run_host(
            self.ctx.h, self.h, _lib._np_ptr(spikes), B, mask, int(nan_to_num), _lib._np_ptr(feats), _lib._np_ptr(raster)))
        feats = self._reorder(feats, order, keys, n_out)
        return (feats, raster) if return_raster else feats

    def diagnostics(self, spikes):
        """Per utterance (participation %, dead neurons, average spikes per neuron) over all N neurons, reduced
        on the device (lsm_reservoir_diagnostics) - what run_network_diagnostics derives from spike_matrix."""
        import torch
        r = self.reservoir
        if not _is_torch(spikes):
            spikes = torch.from_numpy(_lib.as_host(spikes, np.uint8)).cuda(self.ctx.device)
        spikes = spikes.contiguous()
        if spikes.dtype != torch.uint8 or spikes.dim() != 3 or tuple(spikes.shape[1:]) != (r.num_inputs, r.num_steps):
            raise ValueError(f"spikes must be uint8[B,{r.num_inputs},{r.num_steps}]")
        B = spikes.shape[0]
        diag = torch.zeros((B, 2), dtype=torch.int32, device=spikes.device)
        self.ctx.set_stream(torch.cuda.current_stream(spikes.device).cuda_stream)
        self.ctx.check(self.ctx.lib.lsm_reservoir_diagnostics(self.ctx.h, self.h, C.c_void_p(spikes.data_ptr()), B,
                                                              C.c_void_p(diag.data_ptr())))
        d = diag.cpu().numpy().astype(np.int64)
        n = self.num_neurons
        return d[:, 0] / n * 100.0, n - d[:, 0], d[:, 1] / n

    RESERVOIR_MODES = {"event": 0, "dense": 1}

    def set_mode(self, mode: str = "event"):
        """Which arm forms the recurrent current (lsm_reservoir_set_mode): "event" - the event-driven gather over the neurons
        that fired (default) - or "dense" - spikes[B,N] . W[N,N] on the integer tensor cores, one launch per time step.
        Same rasters and features bit for bit."""
        if mode not in self.RESERVOIR_MODES:
            raise ValueError(f"mode must be one of {list(self.RESERVOIR_MODES)}")
        self.ctx.check(self.ctx.lib.lsm_reservoir_set_mode(self.ctx.h, self.h, self.RESERVOIR_MODES[mode]))

    def dense_probe(self, s):
        """Diagnostic (lsm_reservoir_dense_probe): s uint8[B, N] spike bytes -> int32[B, N] recurrent sums W . s of the dense arm."""
        import torch
        s = torch.as_tensor(s, dtype=torch.uint8).cuda(self.ctx.device).contiguous()
        if s.dim() != 2 or s.shape[1] != self.num_neurons:
            raise ValueError(f"s must be uint8[B,{self.num_neurons}]")
        acc = torch.zeros(s.shape, dtype=torch.int32, device=s.device)
        self.ctx.set_stream(torch.cuda.current_stream(s.device).cuda_stream)
        self.ctx.check(self.ctx.lib.lsm_reservoir_dense_probe(self.ctx.h, self.h, C.c_void_p(s.data_ptr()), s.shape[0],
                                                              C.c_void_p(acc.data_ptr())))
        return acc

    def set_gather(self, pointers, row0: int = 0):
        """Fused all-gather (lsm_reservoir_set_gather): launches enqueued from now on also store utterance u's feature row at
        row row0 + u of every matrix in `pointers` (device addresses, e.g. `distributed.PeerAllGather.pointers(k)`); an empty
        list switches it off."""
        ptrs = [int(p) for p in pointers]
        arr = (C.c_void_p * max(1, len(ptrs)))(*ptrs)
        self.ctx.check(self.ctx.lib.lsm_reservoir_set_gather(self.ctx.h, self.h, arr, len(ptrs), int(row0)))

    @staticmethod
    def _reorder(feats, order, keys, n_out):
        """The kernel emits keys in bit order; FEATURE_SETS lists are already in that order, but honour any."""
        if order == keys:
            return feats
        idx = np.concatenate([np.arange(order.index(k) * n_out, (order.index(k) + 1) * n_out) for k in keys])
        if _is_torch(feats):
            import torch
            return feats[:, torch.as_tensor(idx, device=feats.device)]
        return feats[:, idx]

    # ------------------------------------------------------------------ per-sample protocol (snnpy)
    def reset(self):
        self.spike_matrix = None
        self._stats = None

    def set_input_spike_times(self, sample):
        self._sample = _lib.as_host(sample, np.uint8)

    def simulate(self):
        if self._sample is None:
            raise _lib.LsmError("simulate() before set_input_spike_times()")
        feats, raster = self.simulate_batch(self._sample[None], FEATURE_KEYS, nan_to_num=False, return_raster=True)
        self.spike_matrix = raster[0]
        self._stats = feats[0].reshape(len(FEATURE_KEYS), self.num_output_neurons)

    def extract_features_from_spikes(self):
        if self._stats is None:
            raise _lib.LsmError("extract_features_from_spikes() before simulate()")
        return {k: self._stats[i].copy() for i, k in enumerate(FEATURE_KEYS)}

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.lsm_reservoir_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AudioToFeatures:
    """The whole hot path behind one call: PCM -> raw LSM features."""

    def __init__(self, frontend, snn: SNN):
        if frontend.ctx is not snn.ctx:
            raise ValueError("front end and reservoir must share a context (same device)")
        self.frontend, self.snn, self.ctx = frontend, snn, snn.ctx

    @property
    def fused(self) -> bool:
        """True when the pair runs as one kernel (spikes handed over in shared memory)."""
        return bool(self.ctx.lib.lsm_pipeline_is_fused(self.frontend.h, self.snn.h))

    def run_host(self, pcm, feature_keys, nan_to_num: bool = True, out=None, spikes_out: np.ndarray | None = None):
        """Host buffers in, host buffers out (lsm_pipeline_run_host), synchronous.  With pinned buffers (e.g.
        `torch.Tensor.pin_memory().numpy()`) the fused kernel reads the PCM and writes the feature rows directly
        across PCIe; pageable buffers go through a chunked copy/compute pipeline.  `out` may also be a CUDA tensor
        (features stay on the device, e.g. for an all-gather)."""
        keys = list(feature_keys)
        mask = _lib.feature_mask(keys)
        if _lib.mask_keys(mask) != keys:
            raise ValueError("feature keys must be in FEATURE_SETS order")
        B = pcm.shape[0]
        F = len(keys) * self.snn.num_output_neurons
        if out is None:
            out = np.empty((B, F), dtype=np.float64)
        ptr = (lambda a: a.data_ptr() if _is_torch(a) else a.ctypes.data)
        self.ctx.set_stream(None)
        self.ctx.check(self.ctx.lib.lsm_pipeline_run_host(
            self.ctx.h, self.frontend.h, self.snn.h, C.c_void_p(ptr(pcm)), B, mask, int(nan_to_num),
            C.c_void_p(ptr(out)), C.c_void_p(spikes_out.ctypes.data) if spikes_out is not None else None))
        return out

    def run_host_async(self, pcm, feature_keys, out, lane: int = 0, nan_to_num: bool = True):
        """Pinned host buffers only: enqueue and return (lsm_pipeline_run_host_async).  Alternate `lane` 0/1 between
        consecutive batches and finish with `self.ctx.sync_all()`; `out` holds the feature rows after that.
        `pcm` may be float32[B, 16000] or int16[B, 16000] (PCM16 as stored in a WAV file; converted in the kernel)."""
        keys = list(feature_keys)
        mask = _lib.feature_mask(keys)
        if _lib.mask_keys(mask) != keys:
            raise ValueError("feature keys must be in FEATURE_SETS order")
        ptr = (lambda a: a.data_ptr() if _is_torch(a) else a.ctypes.data)
        fn = self.ctx.lib.lsm_pipeline_run_host_async_i16 if _is_int16(pcm) else self.ctx.lib.lsm_pipeline_run_host_async
        self.ctx.check(fn(
            self.ctx.h, self.frontend.h, self.snn.h, C.c_void_p(ptr(pcm)), pcm.shape[0], mask, int(nan_to_num),
            C.c_void_p(ptr(out)), int(lane)))
        return out

    def run(self, pcm, feature_keys, nan_to_num: bool = True, spikes=None, out=None, want_spikes: bool = True):
        """torch CUDA tensors, asynchronous on the current stream.  One fused kernel when the pair allows it;
        with want_spikes=False the spike trains then never leave the SM."""
        import torch
        keys = list(feature_keys)
        mask = _lib.feature_mask(keys)
        if _lib.mask_keys(mask) != keys:
            raise ValueError("feature keys must be in FEATURE_SETS order")
        B = pcm.shape[0]
        fe = self.frontend
        if spikes is None and (want_spikes or not self.fused):
            spikes = torch.empty((B, fe.rows, fe.steps), dtype=torch.uint8, device=pcm.device)
        if out is None:
            out = torch.empty((B, len(keys) * self.snn.num_output_neurons), dtype=torch.float64, device=pcm.device)
        self.ctx.set_stream(torch.cuda.current_stream(pcm.device).cuda_stream)
        fn = self.ctx.lib.lsm_pipeline_run_i16 if _is_int16(pcm) else self.ctx.lib.lsm_pipeline_run
        self.ctx.check(fn(
            self.ctx.h, fe.h, self.snn.h, C.c_void_p(pcm.data_ptr()), B, mask, int(nan_to_num),
            C.c_void_p(spikes.data_ptr()) if spikes is not None else None, C.c_void_p(out.data_ptr())))
        return out, spikes
