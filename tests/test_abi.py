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
    assert api.lib().acm_gpu_abi_version() == 1
    assert C.sizeof(api.Opts) == 64
    assert api.STREAM_DTYPE.itemsize == 72


def test_sm100a_cubin_present():
    """the library carries sm_100a code (and nothing older)"""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "sm_90" not in out and "sm_80" not in out
