"""The two `.npz` files of the pipeline, written with many host threads.

/root/reference/create_dataset.py:176 and extract_lsm_features.py:204 call `np.savez_compressed`, which deflates every array on
one core: for `uint8[12000, 128, 400]` that is seconds, against milliseconds for the kernels that produced it.  `savez_compressed`
here writes the same container - a ZIP archive of `.npy` members, method "deflated", readable by `np.load` and by the
reference's own readers (`load_spike_dataset` :63-73, train_classifier.py:27-31) - but compresses each member as independent
raw-deflate segments on a thread pool (zlib releases the GIL) and joins them into one valid deflate stream: every segment but the
last ends in a sync flush (a byte-aligned empty stored block), the last one carries the final-block bit.  Keys, dtypes, shapes and
array bytes are those of `np.savez_compressed`; only the compressed bytes differ (as they do between zlib versions).
"""
from __future__ import annotations

import io
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

SEGMENT = 8 << 20          # bytes of input per deflate segment
LEVEL = 6                  # zipfile's default for ZIP_DEFLATED, what np.savez_compressed uses


def _npy_buffers(arr):
    """Header and payload of one `.npy` member, as np.lib.format writes it, without copying a large contiguous payload."""
    arr = np.asanyarray(arr)
    if arr.dtype.hasobject or not (arr.flags.c_contiguous or arr.flags.f_contiguous) or arr.nbytes < (1 << 20):
        buf = io.BytesIO()
        np.lib.format.write_array(buf, arr, allow_pickle=True)
        return [buf.getbuffer()]
    head = io.BytesIO()
    np.lib.format.write_array_header_1_0(head, np.lib.format.header_data_from_array_1_0(arr))
    payload = arr if arr.flags.c_contiguous else arr.T          # Fortran order: the header says so, the bytes are the transpose's
    return [head.getbuffer(), memoryview(payload.reshape(-1).view(np.uint8))]


def _deflate_segment(args):
    view, last, level = args
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    out = c.compress(view)
    out += c.flush(zlib.Z_FINISH if last else zlib.Z_SYNC_FLUSH)
    return out, zlib.crc32(view)


def _crc32_combine(crc1: int, crc2: int, len2: int) -> int:
    """zlib's crc32_combine (GF(2) matrix method): CRC of A+B from CRC(A), CRC(B), len(B)."""
    def times(mat, vec):
        s, i = 0, 0
        while vec:
            if vec & 1:
                s ^= mat[i]
            vec >>= 1
            i += 1
        return s

    def square(mat):
        return [times(mat, mat[n]) for n in range(32)]

    if len2 <= 0:
        return crc1
    odd = [0xEDB88320] + [1 << n for n in range(31)]
    even = square(odd)
    odd = square(even)
    while True:
        even = square(odd)
        if len2 & 1:
            crc1 = times(even, crc1)
        len2 >>= 1
        if not len2:
            break
        odd = square(even)
        if len2 & 1:
            crc1 = times(odd, crc1)
        len2 >>= 1
        if not len2:
            break
    return crc1 ^ crc2


def deflate_parallel(buffers, pool, level: int = LEVEL, segment: int = SEGMENT):
    """Raw deflate stream of the concatenation of `buffers` (bytes-like objects) + its CRC-32 and length, compressed as
    independent segments on `pool`."""
    views = []
    for b in buffers:
        mv = memoryview(b).cast("B")
        views += [mv[lo:lo + segment] for lo in range(0, len(mv), segment)]
    if not views:
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        return [c.flush(zlib.Z_FINISH)], 0, 0
    jobs = [(v, i == len(views) - 1, level) for i, v in enumerate(views)]
    parts, crc, total = [], 0, 0
    for (out, c), v in zip(pool.map(_deflate_segment, jobs), views):
        crc = _crc32_combine(crc, c, len(v)) if total else c
        total += len(v)
        parts.append(out)
    return parts, crc & 0xFFFFFFFF, total


def savez_compressed(file, threads: int | None = None, **arrays):
    """Drop-in for `np.savez_compressed(file, **arrays)` (same `.npz` extension rule, keys and member names)."""
    if isinstance(file, (str, os.PathLike)):
        file = os.fspath(file)
        if not file.endswith(".npz"):
            file += ".npz"
        fh, own = open(file, "wb"), True
    else:
        fh, own = file, False
    threads = threads or min(32, os.cpu_count() or 1)
    central = []
    try:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            for key, arr in arrays.items():
                name = (key + ".npy").encode()
                parts, crc, usize = deflate_parallel(_npy_buffers(arr), pool)
                csize = sum(len(p) for p in parts)
                offset = fh.tell()
                # local header: version 45 (zip64), method 8 (deflate), sizes in the zip64 extra field
                extra = struct.pack("<HHQQ", 0x0001, 16, usize, csize)
                fh.write(struct.pack("<IHHHHHIIIHH", 0x04034B50, 45, 0, 8, 0, 0x21, crc, 0xFFFFFFFF, 0xFFFFFFFF, len(name), len(extra)))
                fh.write(name)
                fh.write(extra)
                for p in parts:
                    fh.write(p)
                central.append((name, crc, csize, usize, offset))
        cd_start = fh.tell()
        for name, crc, csize, usize, offset in central:
            extra = struct.pack("<HHQQQ", 0x0001, 24, usize, csize, offset)
            fh.write(struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 45, 45, 0, 8, 0, 0x21, crc, 0xFFFFFFFF, 0xFFFFFFFF,
                                 len(name), len(extra), 0, 0, 0, 0, 0xFFFFFFFF))
            fh.write(name)
            fh.write(extra)
        cd_size = fh.tell() - cd_start
        n = len(central)
        # zip64 end of central directory + locator + classic end record
        fh.write(struct.pack("<IQHHIIQQQQ", 0x06064B50, 44, 45, 45, 0, 0, n, n, cd_size, cd_start))
        fh.write(struct.pack("<IIQI", 0x07064B50, 0, cd_start + cd_size, 1))
        fh.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, min(n, 0xFFFF), min(n, 0xFFFF), 0xFFFFFFFF, 0xFFFFFFFF, 0))
    finally:
        if own:
            fh.close()
