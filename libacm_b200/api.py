"""ctypes binding of libacm_b200.so -- the C ABI of include/acm_gpu.h and include/libacm.h.

This module is plumbing: it hands pointers (numpy host arrays or torch CUDA
tensors) to the C library.  There is no Python or CPU decode path; if the CUDA
library is missing or no device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACM_B200_LIB") or os.path.join(_HERE, "_lib", "libacm_b200.so")
# (ACM_B200_LIB: a tuning build of the same library, tools/build_variants.py)

ACM_OK, ACM_ERR_OTHER, ACM_ERR_OPEN, ACM_ERR_NOT_ACM = 0, -1, -2, -3
ACM_ERR_READ_ERR, ACM_ERR_BADFMT, ACM_ERR_CORRUPT = -4, -5, -6
ACM_ERR_UNEXPECTED_EOF, ACM_ERR_NOT_SEEKABLE = -7, -8

# numpy mirror of struct acm_gpu_stream (include/acm_gpu.h)
STREAM_DTYPE = np.dtype([
    ("in_off", "<u8"), ("in_len", "<u4"), ("reserved0", "<u4"), ("out_off", "<u8"),
    ("total_values", "<u4"), ("channels", "<u4"), ("acm_channels", "<u4"), ("rate", "<u4"),
    ("level", "<u4"), ("rows", "<u4"), ("wavc", "<u4"),
    ("status", "<i4"), ("words", "<u4"), ("reserved1", "<u4"), ("checksum", "<u8"),
], align=True)
assert STREAM_DTYPE.itemsize == 72


class Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("bigendianp", C.c_int32), ("wordlen", C.c_int32),
                ("sgned", C.c_int32), ("force_chans", C.c_int32), ("want_checksums", C.c_int32),
                ("pad_tail", C.c_int32), ("kernel", C.c_int32), ("device_mask", C.c_uint32),
                ("reserved", C.c_int32 * 7)]


class Batch(C.Structure):
    _fields_ = [("blob", C.c_void_p), ("blob_len", C.c_uint64), ("blob_on_device", C.c_int32),
                ("out_on_device", C.c_int32), ("out", C.c_void_p), ("out_len", C.c_uint64),
                ("streams", C.c_void_p), ("n", C.c_uint64)]


_lib = None


def lib():
    """Load the CUDA library; fail loudly (there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing -- build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.acm_gpu_opts_init.argtypes = [C.POINTER(Opts)]
        L.acm_gpu_opts_init.restype = None
        L.acm_gpu_probe.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64,
                                    C.POINTER(Opts)]
        L.acm_gpu_probe.restype = C.c_int64
        L.acm_gpu_layout.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.acm_gpu_layout.restype = C.c_uint64
        L.acm_gpu_decode_batch.argtypes = [C.POINTER(Batch), C.POINTER(Opts)]
        L.acm_gpu_decode_batch.restype = C.c_int
        L.acm_gpu_plan_create.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Opts), C.POINTER(C.c_int)]
        L.acm_gpu_plan_create.restype = C.c_void_p
        L.acm_gpu_plan_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.acm_gpu_plan_run.restype = C.c_int
        L.acm_gpu_plan_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.acm_gpu_plan_fetch.restype = C.c_int
        L.acm_gpu_plan_launches.argtypes = [C.c_void_p]
        L.acm_gpu_plan_launches.restype = C.c_int
        L.acm_gpu_plan_split.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.acm_gpu_plan_split.restype = None
        L.acm_gpu_plan_routes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.acm_gpu_plan_routes.restype = None
        L.acm_gpu_plan_gen_groups.argtypes = [C.c_void_p]
        L.acm_gpu_plan_gen_groups.restype = C.c_int
        L.acm_gpu_plan_last_ms.argtypes = [C.c_void_p]
        L.acm_gpu_plan_last_ms.restype = C.c_float
        L.acm_gpu_plan_destroy.argtypes = [C.c_void_p]
        L.acm_gpu_plan_destroy.restype = None
        L.acm_gpu_generate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_uint64), C.c_int]
        L.acm_gpu_generate.restype = C.c_int
        L.acm_gpu_last_error.restype = C.c_char_p
        L.acm_gpu_abi_version.restype = C.c_int
        _lib = L
    return _lib


def last_error() -> str:
    return lib().acm_gpu_last_error().decode()


class AcmGpuError(RuntimeError):
    pass


def make_opts(device=-1, bigendianp=0, wordlen=2, sgned=1, force_chans=0, want_checksums=0,
              pad_tail=1, kernel=0, device_mask=0) -> Opts:
    o = Opts()
    lib().acm_gpu_opts_init(C.byref(o))
    o.device, o.bigendianp, o.wordlen, o.sgned = device, bigendianp, wordlen, sgned
    o.force_chans, o.want_checksums, o.pad_tail, o.kernel = force_chans, want_checksums, pad_tail, kernel
    o.device_mask = device_mask
    return o


def _ptr(buf):
    """(address, nbytes, on_device) of a numpy array or a torch tensor."""
    if isinstance(buf, np.ndarray):
        assert buf.flags["C_CONTIGUOUS"]
        return buf.ctypes.data, buf.nbytes, 0
    # torch tensor (imported lazily: torch is plumbing, not a dependency of the ABI)
    assert buf.is_contiguous()
    return buf.data_ptr(), buf.numel() * buf.element_size(), 1 if buf.is_cuda else 0


def new_streams(offs, lens) -> np.ndarray:
    s = np.zeros(len(offs), dtype=STREAM_DTYPE)
    s["in_off"] = offs
    s["in_len"] = lens
    return s


def probe(blob, streams: np.ndarray, opts: Opts | None = None) -> int:
    p, n, dev = _ptr(blob)
    ok = lib().acm_gpu_probe(p, n, dev, streams.ctypes.data, len(streams),
                             C.byref(opts) if opts is not None else None)
    if ok < 0:
        raise AcmGpuError(f"acm_gpu_probe failed: {last_error()}")
    return ok


def layout(streams: np.ndarray, wordlen=2) -> int:
    return lib().acm_gpu_layout(streams.ctypes.data, len(streams), wordlen)


def decode_batch(blob, streams: np.ndarray, out, opts: Opts) -> None:
    """acm_gpu_decode_batch: blob/out are numpy (host) or torch CUDA tensors (device)."""
    b = Batch()
    b.blob, b.blob_len, b.blob_on_device = _ptr(blob)
    b.out, b.out_len, b.out_on_device = _ptr(out)
    b.streams, b.n = streams.ctypes.data, len(streams)
    err = lib().acm_gpu_decode_batch(C.byref(b), C.byref(opts))
    if err < 0:
        raise AcmGpuError(f"acm_gpu_decode_batch -> {err}: {last_error()}")


class Plan:
    """Resident-path handle (acm_gpu_plan_*): descriptor tables built once, kernels
    launched on device-resident blob/out as often as wanted."""

    def __init__(self, streams: np.ndarray, opts: Opts):
        err = C.c_int(0)
        self._h = lib().acm_gpu_plan_create(streams.ctypes.data, len(streams), C.byref(opts),
                                            C.byref(err))
        if not self._h:
            raise AcmGpuError(f"acm_gpu_plan_create -> {err.value}: {last_error()}")
        self.n = len(streams)

    def run(self, d_blob, d_out, cuda_stream: int = 0) -> None:
        pb, _, devb = _ptr(d_blob)
        po, _, devo = _ptr(d_out)
        assert devb and devo, "Plan.run needs device-resident tensors"
        if lib().acm_gpu_plan_run(self._h, pb, po, cuda_stream) < 0:
            raise AcmGpuError(f"acm_gpu_plan_run: {last_error()}")

    def fetch(self, streams: np.ndarray, cuda_stream: int = 0) -> None:
        if lib().acm_gpu_plan_fetch(self._h, streams.ctypes.data, cuda_stream) < 0:
            raise AcmGpuError(f"acm_gpu_plan_fetch: {last_error()}")

    @property
    def launches(self) -> int:
        return lib().acm_gpu_plan_launches(self._h)

    def split(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        lib().acm_gpu_plan_split(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def routes(self):
        """(fused level-7 kernel, split path, general throughput path, backstop) stream counts"""
        out = (C.c_uint64 * 4)()
        lib().acm_gpu_plan_routes(self._h, out)
        return tuple(int(v) for v in out)

    def gen_groups(self) -> int:
        """Stream groups of the general path, each on its own CUDA stream (acm_gpu_plan_gen_groups)."""
        return int(lib().acm_gpu_plan_gen_groups(self._h))

    def last_ms(self) -> float:
        return lib().acm_gpu_plan_last_ms(self._h)

    def counters(self) -> np.ndarray:
        """The plan's 64 in-kernel counters (acm_gpu_plan_debug_counters)."""
        buf = np.zeros(64, np.uint64)
        lib().acm_gpu_plan_debug_counters.argtypes = [C.c_void_p, C.c_void_p]
        lib().acm_gpu_plan_debug_counters.restype = C.c_int
        if lib().acm_gpu_plan_debug_counters(self._h, buf.ctypes.data) < 0:
            raise AcmGpuError(f"acm_gpu_plan_debug_counters: {last_error()}")
        return buf

    def close(self):
        if self._h:
            lib().acm_gpu_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate_on_device(plist, d_blob_ptr: int | None, cap: int = 0, device: int = -1):
    """Synthetic corpus generated on the GPU (acm_gpu_generate): plist = libacm_b200.gen.params(...)
    records.  With d_blob_ptr None the call only sizes the corpus.  Returns (offs uint64[n],
    lens uint32[n], bytes used); the images are byte-identical to gen.make_batch's."""
    from . import gen
    n = len(plist)
    arr = (gen.GenParams * n)(*plist)
    offs = np.zeros(n, np.uint64)
    lens = np.zeros(n, np.uint32)
    used = C.c_uint64(0)
    rc = lib().acm_gpu_generate(C.cast(arr, C.c_void_p), n, C.c_void_p(d_blob_ptr) if d_blob_ptr else None, cap,
                                offs.ctypes.data, lens.ctypes.data, C.byref(used), device)
    if rc != 0:
        raise RuntimeError(f"acm_gpu_generate: {rc}: {last_error()}")
    return offs, lens, int(used.value)


def checksum_ref(pcm_bytes: np.ndarray, words: int, wordlen=2, be=0) -> int:
    """Host restatement of acm_gpu_stream.checksum for tests: sum (i+1)*(u_i+1) mod 2^64."""
    b = np.asarray(pcm_bytes, dtype=np.uint8)[: words * wordlen].reshape(words, wordlen).astype(np.uint64)
    if be:
        b = b[:, ::-1]
    u = np.zeros(words, dtype=np.uint64)
    for k in range(wordlen):
        u |= b[:, k] << np.uint64(8 * k)
    i = np.arange(1, words + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(np.sum(i * (u + np.uint64(1)), dtype=np.uint64))
