"""ctypes binding of the C synthetic-ACM generator (csrc/acmgen.c).

The reference ships no encoder or sample files, so this is
where every input stream comes from.  Pure host code; no GPU needed.
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_lib", "libacmgen.so")

DIST_FALLOUT, DIST_STRESS, DIST_SINGLE = 0, 1, 2
INJECT_NONE, INJECT_BAD_IND, INJECT_BAD_TCODE = 0, 1, 2
VALID_INDS = (0,) + tuple(range(3, 17)) + (17, 18, 19, 20, 21, 22, 23, 24, 26, 27, 29)
BAD_INDS = (1, 2, 25, 28, 30, 31)


class GenParams(C.Structure):
    _fields_ = [
        ("level", C.c_uint32), ("rows", C.c_uint32), ("channels", C.c_uint32),
        ("rate", C.c_uint32), ("total_values", C.c_uint32), ("wavc", C.c_uint32),
        ("dist", C.c_uint32), ("single_ind", C.c_uint32), ("pzero", C.c_uint32),
        ("inject", C.c_uint32), ("inject_block", C.c_uint32), ("inject_col", C.c_uint32),
        ("inject_value", C.c_uint32), ("reserved", C.c_uint32), ("seed", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(_LIB_PATH)
        _lib.acmgen_bound.restype = C.c_size_t
        _lib.acmgen_bound.argtypes = [C.POINTER(GenParams)]
        _lib.acmgen_write.restype = C.c_size_t
        _lib.acmgen_write.argtypes = [C.POINTER(GenParams), C.c_void_p, C.c_size_t]
        _lib.acmgen_write_many.restype = C.c_size_t
        _lib.acmgen_write_many.argtypes = [C.POINTER(GenParams), C.c_size_t, C.c_void_p,
                                           C.c_size_t, C.c_void_p, C.c_void_p]
    return _lib


def params(level=7, rows=16, channels=1, rate=22050, total_values=2048, wavc=0,
           dist=DIST_FALLOUT, single_ind=0, pzero=128, seed=1, inject=INJECT_NONE,
           inject_block=0, inject_col=0, inject_value=0) -> GenParams:
    return GenParams(level, rows, channels, rate, total_values, int(wavc), dist, single_ind,
                     pzero, inject, inject_block, inject_col, inject_value, 0, seed)


def make_stream(**kw) -> bytes:
    """One ACM file image as bytes."""
    p = params(**kw)
    cap = lib().acmgen_bound(C.byref(p))
    buf = (C.c_uint8 * cap)()
    n = lib().acmgen_write(C.byref(p), buf, cap)
    if n == 0:
        raise RuntimeError("acmgen_write overflowed its bound")
    return bytes(buf[:n])


def make_batch(plist, threads: int | None = None):
    """Pack many streams into one blob.

    Returns (blob uint8[nbytes], offs uint64[n], lens uint32[n]); every image
    starts on a 16-byte boundary.  Generation is split over `threads` host
    threads (ctypes releases the GIL); the result does not depend on the split.
    """
    n = len(plist)
    arr = (GenParams * n)(*plist)
    bounds = np.array([(lib().acmgen_bound(C.byref(arr[i])) + 15) & ~15 for i in range(n)],
                      dtype=np.int64)
    threads = threads or min(32, os.cpu_count() or 1)
    nchunk = max(1, min(n, threads * 4))
    edges = np.linspace(0, n, nchunk + 1).astype(np.int64)
    pieces = []

    def work(k):
        lo, hi = int(edges[k]), int(edges[k + 1])
        cap = int(bounds[lo:hi].sum()) + 16
        blob = np.empty(cap, dtype=np.uint8)
        offs = np.empty(hi - lo, dtype=np.uint64)
        lens = np.empty(hi - lo, dtype=np.uint32)
        sub = C.cast(C.byref(arr, lo * C.sizeof(GenParams)), C.POINTER(GenParams))
        used = lib().acmgen_write_many(sub, hi - lo, blob.ctypes.data, cap,
                                       offs.ctypes.data, lens.ctypes.data)
        if used == 0 and hi > lo:
            raise RuntimeError("acmgen_write_many overflowed")
        return blob[:used], offs, lens

    with ThreadPoolExecutor(max_workers=threads) as ex:
        pieces = list(ex.map(work, range(nchunk)))
    total = 0
    starts = []
    for blob, _, _ in pieces:
        total = (total + 15) & ~15
        starts.append(total)
        total += len(blob)
    out = np.zeros(total + 64, dtype=np.uint8)  # tail padding for vector loads
    offs = np.empty(n, dtype=np.uint64)
    lens = np.empty(n, dtype=np.uint32)
    at = 0
    for (blob, o, l), s in zip(pieces, starts):
        out[s:s + len(blob)] = blob
        offs[at:at + len(o)] = o + np.uint64(s)
        lens[at:at + len(o)] = l
        at += len(o)
    return out, offs, lens
