"""CPU tier: the C-ABI library loads and exports every symbol the headers declare
(no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

from libacm_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(acm_[a-z0-9_]+)\s*\(", txt))
    return {n for n in names if not n.endswith("_func")}


def test_exports_every_declared_symbol():
    lib = C.CDLL(api.LIB_PATH)
    declared = _declared("acm_gpu.h") | _declared("libacm.h")
    assert {"acm_gpu_decode_batch", "acm_open_decoder", "acm_read", "acm_read_loop", "acm_seek_pcm",
            "acm_pcm_total", "acm_close"} <= declared
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing


def test_abi_version_and_struct_sizes():
    assert api.lib().acm_gpu_abi_version() == 2
    assert C.sizeof(api.Opts) == 64
    assert api.STREAM_DTYPE.itemsize == 72


def test_sm100a_cubin_present():
    """the library carries sm_100a code (and nothing older)"""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "sm_90" not in out and "sm_80" not in out


def test_fast_kernel_geometry_rules():
    """Host logic of the scan-CTA / decode-CTA launch: all CTAs of a launch have to be resident
    together (they wait for each other), so the grid never exceeds its SM budget, there is at least one
    CTA of each kind, every slot has an owner with room for it, and small batches use small grids."""
    lib = C.CDLL(api.LIB_PATH)
    lib.acm_gpu_debug_geometry.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
    lib.acm_gpu_debug_geometry.restype = None
    out = (C.c_uint32 * 3)()
    for sms in (2, 8, 37, 132, 148):
        for budget in sorted({2, sms // 4 or 2, sms}):
            for n in (1, 31, 32, 33, 255, 256, 257, 1250, 10_000, 125_000, 5_000_000):
                lib.acm_gpu_debug_geometry(n, sms, budget, out)
                n_scan, n_work, n_slots = out[0], out[1], out[2]
                assert n_scan >= 1 and n_work >= 1
                assert n_scan + n_work <= max(2, min(budget, sms))
                assert 32 <= n_slots <= n_scan * 256 and n_slots % 32 == 0
                assert n_slots <= n_work * 128            # MAXOWN slots per decode CTA
                assert n_slots <= (n + 31) // 32 * 32      # never more slots than streams (whole warps)
                assert n_work == 1 or n_work % 2 == 1      # slots 0, 32, 64, .. (the longest streams) spread over all owners
    lib.acm_gpu_debug_geometry(10_000, 148, 148, out)
    assert tuple(out) == (33, 115, 8448)                  # a decode-bound launch of that size
    # walk-bound launches (few, long streams): a lane for every stream if the scan share allows it, and ONE
    # scan warp per sub-partition (4 of a scan CTA's 8) when that still fits
    lib.acm_gpu_debug_geometry_walk.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
    lib.acm_gpu_debug_geometry_walk.restype = None
    out4 = (C.c_uint32 * 4)()
    for sms in (8, 37, 132, 148):
        for budget in sorted({sms // 4 or 2, sms}):
            for n in (1, 31, 33, 300, 1250, 2500, 10_000, 125_000):
                lib.acm_gpu_debug_geometry_walk(n, sms, budget, out4)
                n_scan, n_work, n_slots, sw = tuple(out4)
                assert sw in (4, 8) and n_scan >= 1 and n_work >= 1
                assert n_scan + n_work <= max(2, min(budget, sms))
                assert 32 <= n_slots <= n_scan * 32 * sw and n_slots % 32 == 0
                assert n_slots <= n_work * 128 and n_slots <= (n + 31) // 32 * 32
    lib.acm_gpu_debug_geometry_walk(10_000, 148, 148, out4)
    assert tuple(out4) == (40, 107, 10016, 8)             # the bench's full-batch launch: every clip has a lane
    lib.acm_gpu_debug_geometry_walk(1250, 148, 37, out4)
    assert tuple(out4) == (10, 27, 1280, 4)               # a segment of the host path: sparse scan CTAs
    lib.acm_gpu_debug_geometry_walk(300, 148, 148, out4)
    assert tuple(out4)[0] == 3 and tuple(out4)[3] == 4    # a small resident batch
