"""Stage 1 host side: PCM -> spike trains on the GPU (K1 / K1m).

Batched counterpart of the per-utterance body of the reference's dataset loop,
/root/reference/create_dataset.py:148-158:

    spectrogram = audio_to_spectrogram(audio_data, n_filters, filterbank)       # :39-78
    spikes      = convert_spectrogram_to_spikes_hysteresis(spectrogram, SPIKE_THRESHOLDS, HYSTERESIS_GAP)  # :81-98
    spikes      = create_pure_redundancy(spikes, REDUNDANCY_FACTOR)              # :101-104

The same names are kept as thin wrappers (`audio_to_spectrogram`, `convert_spectrogram_to_spikes_hysteresis`
live in create_dataset.py of this package); the work is one CUDA kernel per batch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, filterbank as fb

SAMPLE_RATE = 16000          # create_dataset.py:10
DURATION = 1.0               # create_dataset.py:11
TIME_BINS = 100              # create_dataset.py:12
SPIKE_THRESHOLDS = [0.70, 0.80, 0.90, 0.95]  # create_dataset.py:13
HYSTERESIS_GAP = 0.1         # create_dataset.py:14
REDUNDANCY_FACTOR = 1        # create_dataset.py:17


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class Frontend:
    """PCM float32[B,16000] -> spikes uint8[B, n_filters*redundancy, 400]."""

    def __init__(self, n_filters: int = 128, filterbank: str = "gammatone", thresholds=None, hysteresis_gap=HYSTERESIS_GAP,
                 redundancy: int = REDUNDANCY_FACTOR, n_samples: int = int(SAMPLE_RATE * DURATION),
                 time_bins: int = TIME_BINS, ctx: _lib.Context | None = None, device: int | None = None):
        if filterbank not in _lib.FILTERBANK_KINDS:
            raise ValueError(f"filterbank must be one of {list(_lib.FILTERBANK_KINDS)}")
        self.ctx = ctx or _lib.context(device)
        self.n_filters, self.filterbank, self.redundancy = int(n_filters), filterbank, int(redundancy)
        self.n_samples, self.time_bins = int(n_samples), int(time_bins)
        thresholds = SPIKE_THRESHOLDS if thresholds is None else list(thresholds)
        # create_dataset.py:87,89: descending thresholds; lower bound = threshold - gap, in fp64
        thr = sorted(thresholds, reverse=True)
        lower = [t - hysteresis_gap for t in thr]
        self.n_thresholds = len(thr)
        p = _lib.FrontendParams()
        p.kind = _lib.FILTERBANK_KINDS[filterbank]
        p.channels, p.n_samples, p.n_bins = self.n_filters, self.n_samples, self.time_bins
        p.n_thresholds, p.redundancy = self.n_thresholds, self.redundancy
        for k in range(self.n_thresholds):
            p.thresholds_desc[k] = thr[k]
            p.lower_bounds[k] = lower[k]
        if filterbank == "gammatone":
            hop_time = self.n_samples / (SAMPLE_RATE * self.time_bins)            # create_dataset.py:50
            nwin, hop, ncols = fb.gtgram_strides(SAMPLE_RATE, 0.025, hop_time, self.n_samples)
            p.nwin, p.hop = nwin, hop
            self.table = fb.gammatone_coefs(SAMPLE_RATE, self.n_filters, 50)      # create_dataset.py:51-58
        else:
            hop_length = max(1, int(self.n_samples / self.time_bins))             # create_dataset.py:44
            p.n_fft, p.mel_hop = 2048, hop_length
            ncols = 1 + self.n_samples // hop_length
            self.table = np.ascontiguousarray(fb.mel_basis(SAMPLE_RATE, 2048, self.n_filters), dtype=np.float32)
        self.ncols = ncols
        self.zoom_i0, self.zoom_f = fb.zoom_table(ncols, self.time_bins)
        self.params = p
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.lsm_frontend_create(
            self.ctx.h, C.byref(p), _lib._np_ptr(self.table), _lib._np_ptr(self.zoom_i0), _lib._np_ptr(self.zoom_f), C.byref(h)))
        self.h = h
        if filterbank == "mel":
            # hand the STFT tables over as data too, so the CPU oracle and the GPU transform with identical bits
            self.window = fb.hann_periodic(2048)
            self.tw, self.tw2 = fb.fft_tables(2048)
            self.ctx.check(self.ctx.lib.lsm_frontend_mel_tables(self.ctx.h, self.h, _lib._np_ptr(self.window),
                                                                _lib._np_ptr(self.tw), _lib._np_ptr(self.tw2)))

    @property
    def rows(self) -> int:
        return self.n_filters * self.redundancy

    @property
    def steps(self) -> int:
        return self.time_bins * self.n_thresholds

    def encode(self, pcm, return_spectrogram: bool = False):
        """torch CUDA tensor in -> torch CUDA tensors out (async on the current stream);
        numpy in -> numpy out (H2D, kernel, D2H inside the library, synchronous)."""
        if _is_torch(pcm):
            import torch
            if not pcm.is_cuda:
                raise _lib.LsmError("encode(torch tensor) needs a CUDA tensor; pass numpy for host buffers")
            pcm = pcm.contiguous()
            if pcm.dtype != torch.float32 or pcm.dim() != 2 or pcm.shape[1] != self.n_samples:
                raise ValueError(f"pcm must be float32[B,{self.n_samples}]")
            B = pcm.shape[0]
            spikes = torch.empty((B, self.rows, self.steps), dtype=torch.uint8, device=pcm.device)
            spec = torch.empty((B, self.n_filters, self.time_bins), dtype=torch.float64, device=pcm.device) if return_spectrogram else None
            self.ctx.set_stream(torch.cuda.current_stream(pcm.device).cuda_stream)
            self.ctx.check(self.ctx.lib.lsm_frontend_encode(
                self.ctx.h, self.h, C.c_void_p(pcm.data_ptr()), B, C.c_void_p(spikes.data_ptr()),
                C.c_void_p(spec.data_ptr()) if spec is not None else None))
            return (spikes, spec) if return_spectrogram else spikes
        pcm = _lib.as_host(pcm, np.float32)
        if pcm.ndim == 1:
            pcm = pcm[None, :]
        if pcm.ndim != 2 or pcm.shape[1] != self.n_samples:
            raise ValueError(f"pcm must be float32[B,{self.n_samples}]")
        if return_spectrogram:
            import torch
            s, sp = self.encode(torch.from_numpy(pcm).cuda(self.ctx.device), True)
            return s.cpu().numpy(), sp.cpu().numpy()
        B = pcm.shape[0]
        spikes = np.empty((B, self.rows, self.steps), dtype=np.uint8)
        self.ctx.set_stream(None)
        self.ctx.check(self.ctx.lib.lsm_frontend_encode_host(self.ctx.h, self.h, _lib._np_ptr(pcm), B, _lib._np_ptr(spikes)))
        return spikes

    # ---- gammatone only: how the filter bank is evaluated (include/lsm_b200.h, lsm_frontend_set_mode)
    def set_mode(self, mode: str = "speculative", delta_db: float = 0.0):
        """"exact": every fp64 operation in the reference's order.  "speculative" (default): a cheaper equivalent
        arrangement; utterances in which the derived distance bound between the two arrangements (plus `delta_db`
        decibels) could change any encoder comparison are filtered again exactly - the spike trains are the exact
        path's either way."""
        modes = {"exact": 0, "speculative": 1}
        if mode not in modes:
            raise ValueError(f"mode must be one of {list(modes)}")
        self.ctx.check(self.ctx.lib.lsm_frontend_set_mode(self.ctx.h, self.h, modes[mode], float(delta_db)))

    def set_bound_scale(self, scale: float = 1.0):
        """Diagnostic: multiply the derived error bound of the speculative mode (1 = the guarantee, 0 = off)."""
        self.ctx.check(self.ctx.lib.lsm_frontend_set_bound_scale(self.ctx.h, self.h, float(scale)))

    def error_bound(self) -> np.ndarray:
        """Per channel: bound on |window amplitude (speculative) - window amplitude (reference order)| for max |sample| = 1
        (lsm_gammatone_error_bound; gammatone only)."""
        out = np.zeros(self.n_filters, dtype=np.float64)
        rc = self.ctx.lib.lsm_gammatone_error_bound(_lib._np_ptr(self.table), self.n_filters, self.n_samples, _lib._np_ptr(out))
        if rc != 0:
            raise _lib.LsmError(f"lsm_gammatone_error_bound failed with status {rc}")
        return out

    def audit(self, pcm):
        """Both filter arrangements on every utterance (lsm_frontend_audit): float64[B, 8] =
        [max |dB difference|, max difference / bound, same for the plane's maximum, for its minimum, bound on the maximum,
        bound on the minimum, max |sample|, range in dB]."""
        import torch
        if not _is_torch(pcm):
            pcm = torch.from_numpy(_lib.as_host(pcm, np.float32)).cuda(self.ctx.device)
        pcm = pcm.contiguous()
        out = torch.zeros((pcm.shape[0], 8), dtype=torch.float64, device=pcm.device)
        self.ctx.set_stream(torch.cuda.current_stream(pcm.device).cuda_stream)
        self.ctx.check(self.ctx.lib.lsm_frontend_audit(self.ctx.h, self.h, C.c_void_p(pcm.data_ptr()), pcm.shape[0],
                                                       C.c_void_p(out.data_ptr())))
        return out.cpu().numpy()

    def reruns(self, reset: bool = False) -> int:
        """Utterances the speculative mode had to filter twice (since creation / the last reset)."""
        out = C.c_int64(0)
        self.ctx.check(self.ctx.lib.lsm_frontend_reruns(self.ctx.h, self.h, C.byref(out), int(reset)))
        return int(out.value)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.lsm_frontend_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
