"""Builds and loads tests/emu/emu_generic.cpp (CPU emulation of the generic kernel's
control flow around the product's __host__ __device__ code).  TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "_build", "libemu.so")
SRCS = [os.path.join(ROOT, "tests", "emu", "emu_generic.cpp"),
        os.path.join(ROOT, "tests", "emu", "emu_fast2.cpp"),
        os.path.join(ROOT, "libacm_b200", "csrc", "acm_hostlogic.cpp"),
        os.path.join(ROOT, "libacm_b200", "csrc", "acm_tables.c")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        csrc = os.path.join(ROOT, "libacm_b200", "csrc")
        deps = SRCS + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".h", ".cuh"))]
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
            os.makedirs(os.path.dirname(SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-Wall", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                                   "-I", csrc, "-o", SO] + SRCS)
        _lib = C.CDLL(SO)
        _lib.emu_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        _lib.emu_fast2_check.argtypes = [C.c_void_p, C.c_uint32]
        _lib.emu_fast2_check.restype = C.c_long
        _lib.emu_fast2_stats.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.emu_fast2_stats.restype = C.c_long
    return _lib


def fast2_check(img):
    """Runs the fast kernel's lane-local code (acm_fast2_core.cuh: uni16 walk, column unpackers)
    over every block of a level-7 / 16-row image on the CPU and compares it with the generic
    building blocks.  Returns the number of blocks checked (negative: first difference)."""
    raw = np.frombuffer(bytes(img), np.uint8).copy()
    return lib().emu_fast2_check(raw.ctypes.data, raw.size)


def fast2_stats(images):
    """Walk statistics of level-7 / 16-row images (tests/emu/emu_fast2.cpp::emu_fast2_stats): dict with
    blocks, steps, big_steps (> 31 bits in one step), capped_steps (if no step could advance more than
    31 bits), columns, zero / linear / k / t columns, bits."""
    out = np.zeros(16, np.uint64)
    for img in images:
        raw = np.frombuffer(bytes(img), np.uint8).copy()
        lib().emu_fast2_stats(raw.ctypes.data, raw.size, out.ctypes.data)
    names = ("blocks", "steps", "big_steps", "capped_steps", "columns", "zero", "linear", "k", "t", "bits")
    return {n: int(out[i]) for i, n in enumerate(names)}


def decode(img, be=0, sgned=1, wordlen=2, force_chans=0, lead=0, nthreads=64, trail_fill=0xFF):
    """Returns (open_err, status, words, pcm[total_values*wordlen], checksum).  The image is
    placed `lead` bytes into a buffer whose tail is `trail_fill` garbage, so that the
    end-of-file masking is exercised."""
    raw = np.frombuffer(bytes(img), np.uint8)
    blob = np.full(lead + raw.size + 64, trail_fill, np.uint8)
    blob[lead:lead + raw.size] = raw
    w, s, k, tv = C.c_uint32(), C.c_int(), C.c_uint64(), C.c_uint32()
    out = np.full(max(16, (int(_peek_total(raw)) + 8) * wordlen), 0xAA, np.uint8)
    err = lib().emu_decode(blob.ctypes.data, blob.size, lead, raw.size, force_chans, be, wordlen, sgned, 1,
                           nthreads, out.ctypes.data, C.byref(w), C.byref(s), C.byref(k), C.byref(tv))
    return err, s.value, w.value, out[: tv.value * wordlen].copy(), k.value


def _peek_total(raw):
    # total_values sits at bytes 4..7 (plain) or 32..35 (WAVC); over-estimate safely
    vals = [0]
    for off in (4, 32):
        if raw.size >= off + 4:
            vals.append(int(raw[off]) | int(raw[off + 1]) << 8 | int(raw[off + 2]) << 16 | int(raw[off + 3]) << 24)
    return min(max(vals), 1 << 26)
