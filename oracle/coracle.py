"""TEST INFRASTRUCTURE — ctypes binding of oracle/liblsm_oracle.so (the plain-C CPU oracle).
Only tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__ may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liblsm_oracle.so")
    src = os.path.join(_HERE, "lsm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_log10.restype = C.c_double
        _LIB.oracle_log10.argtypes = [C.c_double]
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def log10(x):
    f = lib().oracle_log10
    return np.array([f(float(v)) for v in np.ravel(x)]).reshape(np.shape(x))


def encoder_tables(thresholds, gap):
    """Descending thresholds and their lower bounds, computed in fp64 exactly as
    create_dataset.py:87-89 does (sorted(reverse=True); threshold - hysteresis_gap)."""
    thr = np.array(sorted(thresholds, reverse=True), dtype=np.float64)
    lower = np.array([t - gap for t in thr], dtype=np.float64)
    return thr, lower


def check_const_division(x, g):
    x = np.ascontiguousarray(x, dtype=np.float64)
    f = lib().oracle_check_const_division
    f.restype = C.c_int64
    return int(f(_p(x, C.c_double), C.c_int64(len(x)), C.c_double(float(g))))


def hysteresis_encode(norm, thresholds, gap, redundancy=1):
    norm = np.ascontiguousarray(norm, dtype=np.float64)
    Cc, nb = norm.shape
    thr, lower = encoder_tables(thresholds, gap)
    out = np.zeros((Cc * redundancy, nb * len(thr)), dtype=np.uint8)
    lib().oracle_hysteresis_encode(_p(norm, C.c_double), Cc, nb, _p(thr, C.c_double), _p(lower, C.c_double),
                                   len(thr), redundancy, _p(out, C.c_uint8))
    return out


def gammatone_encode(pcm, coefs, nwin, hop, nbins, zi0, zf, thresholds, gap, redundancy=1,
                     want_spec=False, nthreads=0):
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    B, L = pcm.shape
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    Cc = coefs.shape[0]
    thr, lower = encoder_tables(thresholds, gap)
    K = len(thr)
    zi0 = np.ascontiguousarray(zi0, dtype=np.int32)
    zf = np.ascontiguousarray(zf, dtype=np.float64)
    spikes = np.zeros((B, Cc * redundancy, nbins * K), dtype=np.uint8)
    spec = np.zeros((B, Cc, nbins), dtype=np.float64) if want_spec else None
    rc = lib().oracle_gammatone_encode(
        _p(pcm, C.c_float), B, L, _p(coefs, C.c_double), Cc, nwin, hop, nbins,
        _p(zi0, C.c_int32), _p(zf, C.c_double), _p(thr, C.c_double), _p(lower, C.c_double),
        K, redundancy, _p(spikes, C.c_uint8), _p(spec, C.c_double), int(nthreads))
    assert rc == 0
    return (spikes, spec) if want_spec else spikes


def mel_encode(pcm, basis, win, tw, tw2, packed, hop, nbins, zi0, zf, thresholds, gap, redundancy=1,
               want_spec=False, nthreads=0):
    """packed = (weights f32, lo, n, off) from filterbank.pack_mel_basis(basis)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    B, L = pcm.shape
    Cc = basis.shape[0]
    thr, lower = encoder_tables(thresholds, gap)
    K = len(thr)
    w, lo, n, off = packed
    zi0 = np.ascontiguousarray(zi0, dtype=np.int32)
    zf = np.ascontiguousarray(zf, dtype=np.float64)
    spikes = np.zeros((B, Cc * redundancy, nbins * K), dtype=np.uint8)
    spec = np.zeros((B, Cc, nbins), dtype=np.float32) if want_spec else None
    f = lib().oracle_mel_encode
    f.argtypes = None
    rc = f(_p(pcm, C.c_float), C.c_int(B), C.c_int(L), C.c_int(2048), C.c_int(hop), C.c_int(Cc), C.c_int(nbins),
           _p(win, C.c_double), _p(tw, C.c_double), _p(tw2, C.c_double), _p(w, C.c_float), _p(lo, C.c_int32),
           _p(n, C.c_int32), _p(off, C.c_int32), _p(zi0, C.c_int32), _p(zf, C.c_double), _p(thr, C.c_double),
           _p(lower, C.c_double), C.c_int(K), C.c_int(redundancy), _p(spikes, C.c_uint8), _p(spec, C.c_float),
           C.c_int(int(nthreads)))
    assert rc == 0
    return (spikes, spec) if want_spec else spikes


def transpose_csr(rowptr, col, val, n):
    """CSR over postsynaptic rows -> CSR over presynaptic neuron (outgoing edges)."""
    post = np.repeat(np.arange(n, dtype=np.int32), np.diff(rowptr))
    order = np.lexsort((post, col))
    t_rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(col, minlength=n), out=t_rowptr[1:])
    return t_rowptr, np.ascontiguousarray(post[order]), np.ascontiguousarray(val[order])


def reservoir_run(r, spikes, feature_mask=0xFF, nan_to_num=False, want_raster=False, nthreads=0):
    """r: lsm_speech_classifier_b200.reservoir.ReservoirDef (plain arrays).  spikes uint8[B,C,T].
    Returns (features float64[B, nkeys*n_out], raster uint8[B,T,N] or None)."""
    spikes = np.ascontiguousarray(spikes, dtype=np.uint8)
    B, Cc, T = spikes.shape
    N = r.num_neurons
    t_rowptr, t_col, t_q = transpose_csr(r.w_rowptr, r.w_col, r.w_q, N)
    w_val = getattr(r, "w_val", None)
    t_val = transpose_csr(r.w_rowptr, r.w_col, np.ascontiguousarray(w_val, dtype=np.float64), N)[2] if w_val is not None else None
    nkeys = bin(feature_mask & 0xFF).count("1")
    n_out = len(r.out_idx)
    feats = np.zeros((B, nkeys * n_out), dtype=np.float64)
    raster = np.zeros((B, T, N), dtype=np.uint8) if want_raster else None
    f = lib().oracle_reservoir_run
    f.argtypes = None
    rc = f(C.c_int(N), C.c_int(Cc), C.c_int(T), C.c_double(r.theta), C.c_int(r.refractory), C.c_int(r.w_shift),
           _p(t_rowptr, C.c_int32), _p(t_col, C.c_int32), _p(t_q, C.c_int32), _p(t_val, C.c_double),
           _p(r.in_rowptr, C.c_int32), _p(r.in_col, C.c_int32), _p(r.in_val, C.c_double),
           _p(r.leak, C.c_double), _p(r.out_idx, C.c_int32), C.c_int(n_out),
           _p(spikes, C.c_uint8), C.c_int(B), C.c_uint32(feature_mask), C.c_int(int(nan_to_num)),
           _p(feats, C.c_double), _p(raster, C.c_uint8), C.c_int(int(nthreads)))
    assert rc == 0
    return feats, raster


def num_threads():
    return int(lib().oracle_num_threads())


def two_arrangements(pcm, coefs, nwin, hop, ncols):
    """One utterance -> (amp_exact, amp_fast), float64[C, ncols] each: the window amplitudes of the reference-order
    cascade and of the product's speculative arrangement restated with the same FMAs (oracle_gammatone_two_arrangements)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    coefs = np.ascontiguousarray(coefs, dtype=np.float64)
    Cc = coefs.shape[0]
    a = np.zeros((Cc, ncols), dtype=np.float64)
    b = np.zeros((Cc, ncols), dtype=np.float64)
    f = lib().oracle_gammatone_two_arrangements
    f.restype = None
    f(_p(pcm, C.c_float), C.c_int(len(pcm)), _p(coefs, C.c_double), C.c_int(Cc), C.c_int(nwin), C.c_int(hop),
      C.c_int(ncols), _p(a, C.c_double), _p(b, C.c_double))
    return a, b
