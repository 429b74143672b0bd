"""In-tree build of the native pieces (no JIT cache: the .so files travel with the repo).

    libacm_b200/_lib/libacm_b200.so   CUDA kernels + C ABI (nvcc, sm_100a only)
    libacm_b200/_lib/libacmgen.so     synthetic stream generator (gcc)
    libacm_b200/_lib/acmgen           generator CLI
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-Xptxas", "-v",
]
CUDA_SOURCES = ["acm_batch.cu", "acm_kernels.cu", "acm_gen2.cu", "acm_fast2.cu", "acm_split.cu", "acm_stream.cu", "acm_gen.cu"]
C_SOURCES = ["acm_tables.c", "acm_hostlogic.cpp"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log=None):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.append(res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return res.stdout


def build_generator(force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    src = os.path.join(CSRC, "acmgen.c")
    deps = [src, os.path.join(CSRC, "acmgen.h"), os.path.join(CSRC, "acmgen_core.h")]
    so = os.path.join(LIBDIR, "libacmgen.so")
    exe = os.path.join(LIBDIR, "acmgen")
    if force or _newer(so, deps):
        _run(["gcc", "-O2", "-Wall", "-Wextra", "-shared", "-fPIC", "-o", so, src])
    if force or _newer(exe, deps):
        _run(["gcc", "-O2", "-DACMGEN_MAIN", "-o", exe, src])
    return so


def build_cuda(force=False, verbose=False):
    """nvcc cross-compiles for sm_100a without a GPU."""
    os.makedirs(LIBDIR, exist_ok=True)
    so = os.path.join(LIBDIR, "libacm_b200.so")
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES + C_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    if not (force or _newer(so, srcs + hdrs)):
        return so
    objs, log = [], []
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        if force or _newer(o, [s] + hdrs):
            if s.endswith(".cu"):
                _run([_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-c", s, "-o", o], log)
            else:
                cc = "g++" if s.endswith(".cpp") else "gcc"
                _run([cc, "-O2", "-Wall", "-fPIC", "-I", INCLUDE, "-I", CSRC, "-c", s, "-o", o], log)
        objs.append(o)
    _run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so, *objs], log)
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return so


def build_oracle():
    """The checkers (test infrastructure): builds them, does not use them."""
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def build_all(force=False, verbose=False):
    build_generator(force)
    build_cuda(force, verbose)
    build_oracle()


def ensure_built():
    """Build only what is missing (no mtime checks: safe when several ranks start at once
    on a box that received prebuilt .so files)."""
    if not os.path.exists(os.path.join(LIBDIR, "libacmgen.so")):
        build_generator()
    if not os.path.exists(os.path.join(LIBDIR, "libacm_b200.so")):
        build_cuda()
    if not os.path.exists(os.path.join(ROOT, "oracle", "libacm_oracle.so")):
        build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("ok")
