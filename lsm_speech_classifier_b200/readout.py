"""Downstream of the path: the standardisation the reference applies to the raw
feature matrix, /root/reference/extract_lsm_features.py:199-201

    scaler = StandardScaler(); X_train_scaled = scaler.fit_transform(X_train_feat); X_test_scaled = scaler.transform(X_test_feat)

on the device, so that after the feature all-gather the [S, F] matrix does not have to visit the host before
it is standardised.  `StandardScaler` here has scikit-learn's attribute names (`mean_`, `var_`, `scale_`,
`n_samples_seen_`) and gives scikit-learn's dense float64 result bit for bit (tests/test_gpu_readout.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class StandardScaler:
    """fit / transform / fit_transform on float64[n, F]; torch CUDA in -> torch CUDA out, numpy in -> numpy out."""

    def __init__(self, ctx: _lib.Context | None = None, device: int | None = None):
        self.ctx = ctx or _lib.context(device)
        self.mean_ = self.var_ = self.scale_ = None
        self.n_samples_seen_ = 0
        self._d = None

    def _dev(self, X):
        import torch
        if _is_torch(X):
            if not X.is_cuda or X.dtype != torch.float64 or X.dim() != 2:
                raise ValueError("X must be a float64[n, F] CUDA tensor (or a numpy array)")
            return X.contiguous(), True
        X = _lib.as_host(X, np.float64)
        if X.ndim != 2:
            raise ValueError("X must be float64[n, F]")
        return torch.from_numpy(X).cuda(self.ctx.device), False

    def fit(self, X, y=None):
        import torch
        Xd, _ = self._dev(X)
        n, F = Xd.shape
        if n == 0:
            raise ValueError("StandardScaler.fit needs at least one sample")
        stats = torch.empty((3, F), dtype=torch.float64, device=Xd.device)
        self.ctx.set_stream(torch.cuda.current_stream(Xd.device).cuda_stream)
        self.ctx.check(self.ctx.lib.lsm_standardize_fit(
            self.ctx.h, C.c_void_p(Xd.data_ptr()), n, F, C.c_void_p(stats[0].data_ptr()),
            C.c_void_p(stats[1].data_ptr()), C.c_void_p(stats[2].data_ptr())))
        self._d = stats
        host = stats.cpu().numpy()
        self.mean_, self.var_, self.scale_ = host[0].copy(), host[1].copy(), host[2].copy()
        self.n_samples_seen_ = int(n)
        return self

    def transform(self, X):
        import torch
        if self._d is None:
            raise _lib.LsmError("StandardScaler.transform() before fit()")
        Xd, was_torch = self._dev(X)
        n, F = Xd.shape
        if F != self._d.shape[1]:
            raise ValueError(f"X has {F} features, the scaler was fitted with {self._d.shape[1]}")
        out = torch.empty_like(Xd)
        self.ctx.set_stream(torch.cuda.current_stream(Xd.device).cuda_stream)
        self.ctx.check(self.ctx.lib.lsm_standardize_transform(
            self.ctx.h, C.c_void_p(Xd.data_ptr()), n, F, C.c_void_p(self._d[0].data_ptr()),
            C.c_void_p(self._d[2].data_ptr()), C.c_void_p(out.data_ptr())))
        return out if was_torch else out.cpu().numpy()

    def fit_transform(self, X, y=None):
        return self.fit(X).transform(X)
