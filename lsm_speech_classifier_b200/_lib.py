"""ctypes binding of liblsmb200.so (include/lsm_b200.h).  There is no CPU fallback: if the
shared library is missing or no CUDA device is present, calls raise."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_PKG, os.environ.get("LSM_SO_NAME", "liblsmb200.so"))      # LSM_SO_NAME: build experiments only

LSM_OK = 0
FILTERBANK_KINDS = {"gammatone": 0, "mel": 1}

# order of FEATURE_SETS['all'] (/root/reference/extract_lsm_features.py:20-22) = bit order of feature_mask
FEATURE_KEYS = ['spike_counts', 'spike_variances', 'mean_spike_times', 'first_spike_times',
                'last_spike_times', 'mean_isi', 'isi_variances', 'burst_counts']

EXPORTS = [
    "lsm_ctx_create", "lsm_ctx_destroy", "lsm_last_error", "lsm_set_stream", "lsm_reset_stream", "lsm_sync", "lsm_launch_count",
    "lsm_sm_count", "lsm_frontend_create", "lsm_frontend_destroy", "lsm_frontend_encode",
    "lsm_frontend_encode_host", "lsm_reservoir_create", "lsm_reservoir_destroy", "lsm_reservoir_run",
    "lsm_reservoir_run_host", "lsm_pipeline_run_host", "lsm_pipeline_run", "lsm_spike_density",
    "lsm_hysteresis_encode", "lsm_fp64_peak_gops", "lsm_pipeline_is_fused", "lsm_frontend_mel_tables", "lsm_reservoir_diagnostics", "lsm_gammatone_design", "lsm_zoom_table", "lsm_standardize_fit", "lsm_standardize_transform",
    "lsm_frontend_set_mode", "lsm_frontend_reruns", "lsm_pipeline_run_host_async", "lsm_sync_all", "lsm_lane_stream", "lsm_logreg_fit", "lsm_logreg_predict", "lsm_pipeline_run_i16", "lsm_pipeline_run_host_async_i16", "lsm_reservoir_set_gather",
    "lsm_frontend_set_bound_scale", "lsm_frontend_audit", "lsm_gammatone_error_bound",
    "lsm_reservoir_set_mode", "lsm_reservoir_dense_probe", "lsm_resample_poly", "lsm_reservoir_create_f64",
    "lsm_ctx_set_host_feed", "lsm_peer_buffer_create", "lsm_peer_buffer_open", "lsm_peer_buffer_close", "lsm_peer_buffer_destroy",
]


class FrontendParams(C.Structure):
    _fields_ = [("kind", C.c_int32), ("channels", C.c_int32), ("n_samples", C.c_int32),
                ("nwin", C.c_int32), ("hop", C.c_int32), ("n_fft", C.c_int32), ("mel_hop", C.c_int32),
                ("n_bins", C.c_int32), ("n_thresholds", C.c_int32), ("redundancy", C.c_int32),
                ("thresholds_desc", C.c_double * 8), ("lower_bounds", C.c_double * 8)]


class ReservoirParams(C.Structure):
    _fields_ = [("num_neurons", C.c_int32), ("num_inputs", C.c_int32), ("num_steps", C.c_int32),
                ("refractory", C.c_int32), ("w_shift", C.c_int32), ("n_out", C.c_int32),
                ("theta", C.c_double)]


class LsmError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen liblsmb200.so and declare signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise LsmError(f"{SO_PATH} is missing: build it with `python -m lsm_speech_classifier_b200.build` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(SO_PATH)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    lib.lsm_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.lsm_ctx_destroy.argtypes = [vp]
    lib.lsm_ctx_destroy.restype = None
    lib.lsm_last_error.argtypes = [vp]
    lib.lsm_last_error.restype = C.c_char_p
    lib.lsm_set_stream.argtypes = [vp, vp]
    lib.lsm_reset_stream.argtypes = [vp]
    lib.lsm_sync.argtypes = [vp]
    lib.lsm_launch_count.argtypes = [vp]
    lib.lsm_launch_count.restype = i64
    lib.lsm_sm_count.argtypes = [vp]
    lib.lsm_frontend_create.argtypes = [vp, C.POINTER(FrontendParams), vp, vp, vp, C.POINTER(vp)]
    lib.lsm_frontend_destroy.argtypes = [vp]
    lib.lsm_frontend_destroy.restype = None
    lib.lsm_frontend_encode.argtypes = [vp, vp, vp, i32, vp, vp]
    lib.lsm_frontend_encode_host.argtypes = [vp, vp, vp, i32, vp]
    lib.lsm_reservoir_create.argtypes = [vp, C.POINTER(ReservoirParams)] + [vp] * 8 + [C.POINTER(vp)]
    lib.lsm_reservoir_create_f64.argtypes = [vp, C.POINTER(ReservoirParams)] + [vp] * 8 + [C.POINTER(vp)]
    lib.lsm_reservoir_destroy.argtypes = [vp]
    lib.lsm_reservoir_destroy.restype = None
    lib.lsm_reservoir_run.argtypes = [vp, vp, vp, i32, u32, i32, vp, vp]
    lib.lsm_reservoir_run_host.argtypes = [vp, vp, vp, i32, u32, i32, vp, vp]
    lib.lsm_pipeline_run_host.argtypes = [vp, vp, vp, vp, i32, u32, i32, vp, vp]
    lib.lsm_pipeline_run.argtypes = [vp, vp, vp, vp, i32, u32, i32, vp, vp]
    lib.lsm_spike_density.argtypes = [vp, vp, i64, vp]
    lib.lsm_hysteresis_encode.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, i32, i32, vp]
    lib.lsm_fp64_peak_gops.argtypes = [vp, vp]
    lib.lsm_pipeline_is_fused.argtypes = [vp, vp]
    lib.lsm_frontend_mel_tables.argtypes = [vp, vp, vp, vp, vp]
    lib.lsm_reservoir_diagnostics.argtypes = [vp, vp, vp, i32, vp]
    lib.lsm_gammatone_design.argtypes = [C.c_double, i32, C.c_double, vp]
    lib.lsm_zoom_table.argtypes = [i32, i32, vp, vp]
    lib.lsm_pipeline_run_host_async.argtypes = [vp, vp, vp, vp, i32, u32, i32, vp, i32]
    lib.lsm_sync_all.argtypes = [vp]
    lib.lsm_reservoir_set_gather.argtypes = [vp, vp, vp, i32, i64]
    lib.lsm_pipeline_run_i16.argtypes = [vp, vp, vp, vp, i32, u32, i32, vp, vp]
    lib.lsm_pipeline_run_host_async_i16.argtypes = [vp, vp, vp, vp, i32, u32, i32, vp, i32]
    lib.lsm_logreg_fit.argtypes = [vp, vp, vp, i32, i32, i32, C.c_double, i32, C.c_double, vp, vp, vp]
    lib.lsm_logreg_predict.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.lsm_lane_stream.argtypes = [vp, i32]
    lib.lsm_lane_stream.restype = vp
    lib.lsm_frontend_set_mode.argtypes = [vp, vp, i32, C.c_double]
    lib.lsm_frontend_reruns.argtypes = [vp, vp, vp, i32]
    lib.lsm_frontend_set_bound_scale.argtypes = [vp, vp, C.c_double]
    lib.lsm_frontend_audit.argtypes = [vp, vp, vp, i32, vp]
    lib.lsm_gammatone_error_bound.argtypes = [vp, i32, i32, vp]
    lib.lsm_ctx_set_host_feed.argtypes = [vp, i32]
    lib.lsm_peer_buffer_create.argtypes = [vp, i64, C.POINTER(vp), vp]
    lib.lsm_peer_buffer_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.lsm_peer_buffer_close.argtypes = [vp, vp]
    lib.lsm_peer_buffer_destroy.argtypes = [vp, vp]
    lib.lsm_resample_poly.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, i32, vp]
    lib.lsm_reservoir_set_mode.argtypes = [vp, vp, i32]
    lib.lsm_reservoir_dense_probe.argtypes = [vp, vp, vp, i32, vp]
    lib.lsm_standardize_fit.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    lib.lsm_standardize_transform.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    _lib = lib
    return lib


def feature_mask(keys) -> int:
    m = 0
    for k in keys:
        m |= 1 << FEATURE_KEYS.index(k)
    return m


def mask_keys(mask: int):
    return [k for i, k in enumerate(FEATURE_KEYS) if mask & (1 << i)]


def _np_ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One per (process, device).  Owns the lsm_ctx handle."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.lsm_ctx_create(C.byref(h), int(device))
        if rc != LSM_OK:
            raise LsmError(f"lsm_ctx_create(device={device}) failed with status {rc}: a B200-class CUDA device "
                           "is required (there is no CPU fallback)")
        self.h = h
        self.device = int(device)

    def check(self, rc: int):
        if rc != LSM_OK:
            raise LsmError(f"status {rc}: {self.lib.lsm_last_error(self.h).decode(errors='replace')}")

    def set_stream(self, cuda_stream_handle: int | None):
        """An int (0 = CUDA's default stream, what torch uses unless told otherwise) enqueues there;
        None goes back to the ctx's own stream (used by the synchronous *_host calls)."""
        if cuda_stream_handle is None:
            self.check(self.lib.lsm_reset_stream(self.h))
        else:
            self.check(self.lib.lsm_set_stream(self.h, C.c_void_p(int(cuda_stream_handle))))

    def sync(self):
        self.check(self.lib.lsm_sync(self.h))

    def set_host_feed(self, mode: str = "zero_copy"):
        """How run_host_async brings pinned PCM to the kernel: "zero_copy" (the kernel reads the host buffer itself) or
        "copy_engine" (one cudaMemcpyAsync into a device staging buffer per call, overlapped with the other lane's kernel)."""
        modes = {"zero_copy": 0, "copy_engine": 1}
        if mode not in modes:
            raise ValueError(f"mode must be one of {list(modes)}")
        self.check(self.lib.lsm_ctx_set_host_feed(self.h, modes[mode]))

    def lane_stream(self, lane: int) -> int:
        """cudaStream_t handle of launch lane 0 / 1 (wrap with torch.cuda.ExternalStream to order other work after it)."""
        h = self.lib.lsm_lane_stream(self.h, int(lane))
        if not h:
            raise LsmError("lane must be 0 or 1")
        return int(h)

    def sync_all(self):
        """Wait for everything enqueued through this ctx (both launch lanes of the asynchronous host calls)."""
        self.check(self.lib.lsm_sync_all(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.lsm_launch_count(self.h))

    @property
    def sm_count(self) -> int:
        return int(self.lib.lsm_sm_count(self.h))

    def fp64_peak_gops(self) -> float:
        out = C.c_double(0.0)
        self.check(self.lib.lsm_fp64_peak_gops(self.h, C.byref(out)))
        return float(out.value)

    def close(self):
        if getattr(self, "h", None):
            self.lib.lsm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts: dict[int, Context] = {}


def context(device: int | None = None) -> Context:
    """Process-wide ctx for a device (default: torch's current device)."""
    if device is None:
        import torch
        if not torch.cuda.is_available():
            raise LsmError("no CUDA device: liblsmb200 has no CPU fallback")
        device = torch.cuda.current_device()
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


def as_host(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)
