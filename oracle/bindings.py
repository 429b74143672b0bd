"""ctypes access to the checkers -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Two checkers, same call shape:

  Ref     oracle/_ref/libacm_ref.so : the UNMODIFIED reference decoder (built from
          /root/reference by oracle/Makefile; travels to the GPU box prebuilt)
  Oracle  oracle/libacm_oracle.so   : our C restatement (acm_oracle.c), pinned
          against Ref by tests/test_oracle.py

`best()` returns Ref when it is available and Oracle otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libacm_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libacm_ref.so")
REF_ACMTOOL = os.path.join(_HERE, "_ref", "acmtool")


class Info(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "channels", "rate", "acm_channels", "acm_level", "acm_cols", "acm_rows",
        "total_values", "block_len", "wavc")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Result:
    __slots__ = ("open_err", "status", "words", "pcm", "info")

    def __init__(self, open_err, status, words, pcm, info):
        self.open_err, self.status, self.words, self.pcm, self.info = open_err, status, words, pcm, info


def _as_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


class Oracle:
    kind = "port"

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            raise RuntimeError(f"{ORACLE_SO} missing: run make -C oracle")
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.ora_decode.restype = C.c_int
        self.lib.ora_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_int), C.POINTER(Info)]
        self.lib.ora_parse.restype = C.c_int
        self.lib.ora_parse.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(Info)]
        self.lib.ora_time_decode.restype = C.c_double
        self.lib.ora_time_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int,
                                             C.POINTER(C.c_ulonglong)]

    def parse(self, data, force_chans=0):
        a = _as_u8(data)
        info = Info()
        err = self.lib.ora_parse(a.ctypes.data, a.size, force_chans, C.byref(info))
        return err, info

    def decode(self, data, force_chans=0, be=0, wordlen=2, sgned=1, out_words=None) -> Result:
        """Decode a whole file image; PCM is zero-padded to total_values words
        (what `acmtool -d -r` writes, acmtool.c:293-310) unless out_words is given."""
        a = _as_u8(data)
        err, info = self.parse(a, force_chans)
        if err < 0:
            return Result(err, 0, 0, np.zeros(0, np.uint8), info)
        nwords = info.total_values if out_words is None else out_words
        out = np.zeros(nwords * wordlen, dtype=np.uint8)
        words, status = C.c_uint32(0), C.c_int(0)
        err = self.lib.ora_decode(a.ctypes.data, a.size, force_chans, be, wordlen, sgned,
                                  out.ctypes.data, out.size, C.byref(words), C.byref(status),
                                  C.byref(info))
        return Result(err, status.value, words.value, out, info)

    def time_decode(self, data, reps=1):
        a = _as_u8(data)
        words = C.c_ulonglong(0)
        secs = self.lib.ora_time_decode(a.ctypes.data, a.size, reps, C.byref(words))
        return secs, words.value


class Ref:
    kind = "reference"

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise RuntimeError(f"{REF_SO} missing: run make -C oracle (needs /root/reference)")
        self.lib = C.CDLL(REF_SO)
        self.lib.ref_decode.restype = C.c_int
        self.lib.ref_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_int), C.POINTER(Info)]
        self.lib.ref_time_decode.restype = C.c_double
        self.lib.ref_time_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int,
                                             C.POINTER(C.c_ulonglong)]

    def decode(self, data, force_chans=0, be=0, wordlen=2, sgned=1, out_words=None) -> Result:
        if wordlen != 2:
            raise ValueError("the reference only implements wordlen 2 (decode.c:832-835)")
        a = _as_u8(data)
        info = Info()
        # first pass with a 0-byte buffer just opens and reports the header
        words, status = C.c_uint32(0), C.c_int(0)
        err = self.lib.ref_decode(a.ctypes.data, a.size, force_chans, be, sgned, None, 0,
                                  C.byref(words), C.byref(status), C.byref(info))
        if err < 0:
            return Result(err, 0, 0, np.zeros(0, np.uint8), info)
        nwords = info.total_values if out_words is None else out_words
        out = np.zeros(nwords * 2, dtype=np.uint8)
        err = self.lib.ref_decode(a.ctypes.data, a.size, force_chans, be, sgned,
                                  out.ctypes.data, out.size, C.byref(words), C.byref(status),
                                  C.byref(info))
        return Result(err, status.value, words.value, out, info)

    def time_decode(self, data, reps=1):
        a = _as_u8(data)
        words = C.c_ulonglong(0)
        secs = self.lib.ref_time_decode(a.ctypes.data, a.size, reps, C.byref(words))
        return secs, words.value


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def best():
    return Ref() if have_ref() else Oracle()
