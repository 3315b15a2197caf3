"""Downstream of the path: the standardisation the reference applies to the raw
feature matrix, /root/reference/extract_lsm_features.py:199-201

    scaler = StandardScaler(); X_train_scaled = scaler.fit_transform(X_train_feat); X_test_scaled = scaler.transform(X_test_feat)

and the classifier the reference then fits, /root/reference/train_classifier.py:36-47

    LogisticRegression(random_state=42, max_iter=1000).fit(X_train, y_train); .predict(X_test)

on the device, so that after the feature all-gather the [S, F] matrix does not have to visit the host before
it is standardised and read out.  `StandardScaler` here has scikit-learn's attribute names (`mean_`, `var_`, `scale_`,
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


class LogisticRegression:
    """Multinomial logistic regression with scikit-learn's lbfgs objective and attribute names (`coef_`, `intercept_`,
    `classes_`, `n_iter_`), fitted on the device (lsm_logreg_fit).  The optimum is unique; agreement with scikit-learn is to
    solver tolerance (same predictions on >= 99 % of samples, accuracy within 0.5 points - tests/test_gpu_readout.py)."""

    def __init__(self, C: float = 1.0, max_iter: int = 100, tol: float = 1e-4, random_state=None,
                 ctx: _lib.Context | None = None, device: int | None = None):
        self.C, self.max_iter, self.tol, self.random_state = float(C), int(max_iter), float(tol), random_state
        self.ctx = ctx or _lib.context(device)
        self.coef_ = self.intercept_ = self.classes_ = None
        self.n_iter_ = None

    def _dev(self, X):
        import torch
        if _is_torch(X):
            if not X.is_cuda or X.dtype != torch.float64 or X.dim() != 2:
                raise ValueError("X must be a float64[n, F] CUDA tensor (or a numpy array)")
            return X.contiguous()
        X = _lib.as_host(X, np.float64)
        if X.ndim != 2:
            raise ValueError("X must be float64[n, F]")
        return torch.from_numpy(X).cuda(self.ctx.device)

    def fit(self, X, y):
        import torch
        Xd = self._dev(X)
        y = np.asarray(y.cpu().numpy() if _is_torch(y) else y)
        self.classes_, yi = np.unique(y, return_inverse=True)
        K = len(self.classes_)
        if K < 3:
            # scikit-learn fits two classes with the binary (single-row, sigmoid) objective, whose L2 term differs from the
            # symmetric multinomial one implemented here; the reference has 12 classes
            raise ValueError("the device readout implements the multinomial objective: at least three classes")
        n, F = Xd.shape
        if len(yi) != n:
            raise ValueError("X and y have different lengths")
        yd = torch.from_numpy(np.ascontiguousarray(yi, dtype=np.int32)).cuda(Xd.device)
        coef = np.zeros((K, F), dtype=np.float64)
        icpt = np.zeros(K, dtype=np.float64)
        n_iter = C.c_int32(0)
        torch.cuda.current_stream(Xd.device).synchronize()
        self.ctx.set_stream(None)
        self.ctx.check(self.ctx.lib.lsm_logreg_fit(
            self.ctx.h, C.c_void_p(Xd.data_ptr()), C.c_void_p(yd.data_ptr()), n, F, K, self.C, self.max_iter, self.tol,
            _lib._np_ptr(coef), _lib._np_ptr(icpt), C.byref(n_iter)))
        self.coef_, self.intercept_, self.n_iter_ = coef, icpt, np.array([n_iter.value], dtype=np.int32)
        return self

    def predict(self, X):
        import torch
        if self.coef_ is None:
            raise _lib.LsmError("LogisticRegression.predict() before fit()")
        Xd = self._dev(X)
        n, F = Xd.shape
        if F != self.coef_.shape[1]:
            raise ValueError(f"X has {F} features, the model was fitted with {self.coef_.shape[1]}")
        pred = torch.empty((n,), dtype=torch.int32, device=Xd.device)
        torch.cuda.current_stream(Xd.device).synchronize()
        self.ctx.set_stream(None)
        self.ctx.check(self.ctx.lib.lsm_logreg_predict(
            self.ctx.h, C.c_void_p(Xd.data_ptr()), n, F, len(self.classes_), _lib._np_ptr(self.coef_),
            _lib._np_ptr(self.intercept_), C.c_void_p(pred.data_ptr())))
        return self.classes_[pred.cpu().numpy()]

    def score(self, X, y):
        y = np.asarray(y.cpu().numpy() if _is_torch(y) else y)
        return float(np.mean(self.predict(X) == y))
