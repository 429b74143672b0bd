"""Loads oracle/ref_driver.c twice: linked with the unmodified reference (oracle/_ref/libacm_ref.so)
and, compiled against include/libacm.h, linked with libacm_b200.so.  The driver only uses the public
libacm.h API, so the very same harness exercises both libraries (API-parity tier).  TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINE_SO = os.path.join(ROOT, "tests", "_build", "libacm_b200_driver.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libacm_ref.so")
LIBDIR = os.path.join(ROOT, "libacm_b200", "_lib")


class ACMInfo(C.Structure):
    _fields_ = [(n, C.c_uint) for n in ("channels", "rate", "acm_id", "acm_version", "acm_channels",
                                        "acm_level", "acm_cols", "acm_rows")]


class IoCallbacks(C.Structure):
    _fields_ = [("read_func", C.c_void_p), ("seek_func", C.c_void_p), ("close_func", C.c_void_p),
                ("get_length_func", C.c_void_p)]


class ACMStream(C.Structure):
    """public part of struct ACMStream (reference libacm.h:71-100); identical in include/libacm.h"""
    _fields_ = [("info", ACMInfo), ("total_values", C.c_uint), ("io_arg", C.c_void_p), ("io", IoCallbacks),
                ("data_len", C.c_uint), ("buf", C.c_void_p), ("buf_max", C.c_uint), ("buf_size", C.c_uint),
                ("buf_pos", C.c_uint), ("bit_avail", C.c_uint), ("bit_data", C.c_uint),
                ("buf_start_ofs", C.c_uint), ("block_len", C.c_uint), ("wrapbuf_len", C.c_uint),
                ("block", C.c_void_p), ("wrapbuf", C.c_void_p), ("ampbuf", C.c_void_p), ("midbuf", C.c_void_p),
                ("flags", C.c_uint), ("stream_pos", C.c_uint), ("block_pos", C.c_uint)]


def _build_mine():
    src = os.path.join(ROOT, "oracle", "ref_driver.c")
    lib = os.path.join(LIBDIR, "libacm_b200.so")
    if (not os.path.exists(MINE_SO) or os.path.getmtime(MINE_SO) < os.path.getmtime(src)
            or os.path.getmtime(MINE_SO) < os.path.getmtime(lib)):
        os.makedirs(os.path.dirname(MINE_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", MINE_SO,
                               src, "-L", LIBDIR, "-lacm_b200", "-Wl,-rpath," + LIBDIR])
    return MINE_SO


class Lib:
    def __init__(self, path, name):
        self.name = name
        L = self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                               C.POINTER(C.c_int)]
        L.ref_stream.restype = C.POINTER(ACMStream)
        L.ref_stream.argtypes = [C.c_void_p]
        for f in ("ref_n_seek", "ref_n_close"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_close.restype = None
        P = C.POINTER(ACMStream)
        L.acm_read.argtypes = [P, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int]
        L.acm_read_loop.argtypes = [P, C.c_void_p, C.c_uint, C.c_int, C.c_int, C.c_int]
        L.acm_seek_pcm.argtypes = [P, C.c_uint]
        L.acm_seek_time.argtypes = [P, C.c_uint]
        for f in ("acm_bitrate", "acm_rate", "acm_channels", "acm_raw_total", "acm_raw_tell", "acm_pcm_total",
                  "acm_pcm_tell", "acm_time_total", "acm_time_tell"):
            getattr(L, f).restype = C.c_uint
            getattr(L, f).argtypes = [P]
        L.acm_seekable.argtypes = [P]
        L.acm_strerror.restype = C.c_char_p
        L.acm_strerror.argtypes = [C.c_int]
        L.acm_open_file.argtypes = [C.POINTER(P), C.c_char_p, C.c_int]
        L.acm_close.argtypes = [P]
        L.acm_close.restype = None


class Handle:
    """One opened stream on one library, through ref_driver.c's memory data source."""

    def __init__(self, lib: Lib, img, force_chans=0, seekable=1, with_seek=1, chunk=0):
        self.L = lib.lib
        self.buf = np.frombuffer(bytes(img), np.uint8).copy()
        err, closed = C.c_int(0), C.c_int(0)
        self.h = self.L.ref_open(self.buf.ctypes.data, self.buf.size, force_chans, seekable, with_seek, chunk,
                                 C.byref(err), C.byref(closed))
        self.err, self.closed_on_fail = err.value, closed.value
        self.s = self.L.ref_stream(self.h) if self.h else None

    def _stream(self):
        if not self.s:
            raise RuntimeError(f"stream is not open (open returned {self.err})")
        return self.s

    def read(self, n, be=0, wordlen=2, sgned=1, loop=False, null=False):
        self._stream()
        out = np.zeros(max(n, 1), np.uint8)
        f = self.L.acm_read_loop if loop else self.L.acm_read
        r = f(self.s, None if null else out.ctypes.data, n, be, wordlen, sgned)
        return r, (bytes(out[:r]) if r > 0 and not null else b"")

    def seek(self, pcm):
        return self.L.acm_seek_pcm(self._stream(), pcm)

    def seek_time(self, ms):
        return self.L.acm_seek_time(self._stream(), ms)

    def state(self):
        s = self._stream().contents
        return dict(stream_pos=s.stream_pos, block_pos=s.block_pos, block_len=s.block_len,
                    total_values=s.total_values, data_len=s.data_len, wrapbuf_len=s.wrapbuf_len,
                    info={n: getattr(s.info, n) for n, _ in ACMInfo._fields_})

    def getters(self):
        L, s = self.L, self._stream()
        return dict(rate=L.acm_rate(s), channels=L.acm_channels(s), raw_total=L.acm_raw_total(s),
                    pcm_total=L.acm_pcm_total(s), pcm_tell=L.acm_pcm_tell(s), time_total=L.acm_time_total(s),
                    time_tell=L.acm_time_tell(s), bitrate=L.acm_bitrate(s), seekable=L.acm_seekable(s))

    def counts(self):
        return self.L.ref_n_seek(self.h), self.L.ref_n_close(self.h)

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None


def mine() -> Lib:
    return Lib(_build_mine(), "libacm_b200")


def ref() -> Lib:
    return Lib(REF_SO, "reference")


def have_ref() -> bool:
    return os.path.exists(REF_SO)
