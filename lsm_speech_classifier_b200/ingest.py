"""Audio ingest: WAV decode, mono mix-down, resampling to 16 kHz.

The reference reads every clip with `librosa.load(filepath, sr=SAMPLE_RATE, mono=True)` (/root/reference/create_dataset.py:26):
decode to float32 (integer PCM scaled by 2^-(bits-1), as libsndfile does), average the channels, resample to 16 kHz.  Here:

* `read_wav` parses RIFF/WAVE itself (8/16/24/32-bit PCM, 32/64-bit IEEE float, WAVE_FORMAT_EXTENSIBLE) - the standard
  library's `wave` refuses float files;
* `resample_poly` runs the sample-rate conversion on the GPU (lsm_resample_poly) with the arithmetic of
  `scipy.signal.resample_poly`, i.e. librosa's `res_type="polyphase"`, bit for bit.  librosa's *default* resampler (soxr_hq) is
  a closed recipe that cannot be restated, so non-16 kHz files are "parity unpinned" against the reference and pinned against
  scipy.  16 kHz files (all of Speech Commands) never reach the resampler.
"""
from __future__ import annotations

import ctypes as C
import math
import struct

import numpy as np

from . import _lib


def read_wav(path):
    """-> (float32[n_frames, n_channels], sample_rate).  Raises ValueError for anything that is not a PCM / float WAV file."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt, body = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        chunk = data[pos + 8: pos + 8 + size]
        if cid == b"fmt ":
            fmt = chunk
        elif cid == b"data":
            body = chunk
            if fmt is not None:
                break
        pos += 8 + size + (size & 1)
    if fmt is None or body is None or len(fmt) < 16:
        raise ValueError("missing fmt or data chunk")
    tag, ch, rate, _, align, bits = struct.unpack_from("<HHIIHH", fmt, 0)
    if tag == 0xFFFE and len(fmt) >= 26:                      # WAVE_FORMAT_EXTENSIBLE: the real tag leads the sub-format GUID
        tag = struct.unpack_from("<H", fmt, 24)[0]
    if ch < 1 or rate < 1:
        raise ValueError("bad channel count or sample rate")
    n = len(body) // (ch * (bits // 8)) * ch
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(body, np.uint8, n).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(body, "<i2", n).astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(body, np.uint8, n * 3).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v & 0x800000, v - (1 << 24), v)
            x = v.astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(body, "<i4", n).astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError(f"{bits}-bit PCM is not supported")
    elif tag == 3:
        if bits == 32:
            x = np.frombuffer(body, "<f4", n).astype(np.float32)
        elif bits == 64:
            x = np.frombuffer(body, "<f8", n).astype(np.float32)
        else:
            raise ValueError(f"{bits}-bit float is not supported")
    else:
        raise ValueError(f"WAVE format tag {tag} is not PCM or IEEE float")
    return x.reshape(-1, ch), int(rate)


def polyphase_design(n_in: int, up: int, down: int, dtype=np.float32):
    """The arguments scipy.signal.resample_poly derives before it calls upfirdn (default window ("kaiser", 5.0), padtype
    "constant"): reduced ratio, per-phase taps (transposed and flipped as scipy.signal._upfirdn._pad_h does), leading outputs
    to remove and the output length."""
    from scipy.signal import firwin
    g = math.gcd(int(up), int(down))
    up, down = int(up) // g, int(down) // g
    n_out = n_in * up
    n_out = n_out // down + bool(n_out % down)
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)).astype(dtype)
    h *= up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down

    def out_len(len_h):        # scipy.signal._upfirdn._output_len
        return (((n_in - 1) * up + len_h) - 1) // down + 1

    n_post_pad = 0
    while out_len(len(h) + n_pre_pad + n_post_pad) < n_out + n_pre_remove:
        n_post_pad += 1
    h = np.concatenate((np.zeros(n_pre_pad, dtype), h, np.zeros(n_post_pad, dtype)))
    hpp = -(-len(h) // up)
    h_full = np.zeros(hpp * up, dtype)
    h_full[:len(h)] = h
    taps = np.ascontiguousarray(h_full.reshape(hpp, up).T[:, ::-1])        # [up][hpp]
    return up, down, taps, hpp, n_pre_remove, n_out


def resample_poly(x, orig_sr: int, target_sr: int, ctx: _lib.Context | None = None):
    """float32[B, n] (numpy or torch CUDA) at orig_sr -> float32[B, ceil(n * target / orig)] at target_sr on the GPU; the values
    of scipy.signal.resample_poly(x, target_sr, orig_sr, axis=-1)."""
    import torch
    is_torch = type(x).__module__.startswith("torch")
    single = (x.dim() if is_torch else np.ndim(x)) == 1
    if orig_sr == target_sr:
        return x
    ctx = ctx or _lib.context()
    xd = x if is_torch else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda(ctx.device)
    if single:
        xd = xd[None]
    xd = xd.contiguous().float()
    B, n_in = xd.shape
    up, down, taps, hpp, n_pre_remove, n_out = polyphase_design(n_in, target_sr, orig_sr)
    out = torch.empty((B, n_out), dtype=torch.float32, device=xd.device)
    ctx.set_stream(torch.cuda.current_stream(xd.device).cuda_stream)
    for lo in range(0, B, 65535):
        hi = min(B, lo + 65535)
        ctx.check(ctx.lib.lsm_resample_poly(ctx.h, C.c_void_p(xd[lo:hi].data_ptr()), hi - lo, n_in, up, down, _lib._np_ptr(taps), hpp,
                                            n_pre_remove, n_out, C.c_void_p(out[lo:hi].data_ptr())))
    if single:
        out = out[0]
    return out if is_torch else out.cpu().numpy()


def load_audio(path, sample_rate: int = 16000, duration: float = 1.0):
    """The reference's load_audio_file contract (create_dataset.py:22-36): float32 mono at `sample_rate`, exactly
    sample_rate * duration samples (zero padded or truncated)."""
    x, rate = read_wav(path)
    y = x[:, 0] if x.shape[1] == 1 else x.mean(axis=1, dtype=np.float32)        # librosa.to_mono
    if rate != sample_rate:
        # only the part that can reach the first `duration` seconds needs converting (plus the filter's reach)
        keep = int(math.ceil((duration + 0.05) * rate))
        y = resample_poly(np.ascontiguousarray(y[:keep], dtype=np.float32), rate, sample_rate)
    target = int(sample_rate * duration)
    if len(y) < target:
        y = np.pad(y, (0, target - len(y)))
    return np.ascontiguousarray(y[:target], dtype=np.float32)
