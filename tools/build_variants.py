"""Tuning builds of libacm_b200.so with different -D flags (kernel geometry experiments).

    python tools/build_variants.py name1="-DF2_W=12 -DACM_UNI_KBITS=8" name2=...
    ACM_B200_LIB=libacm_b200/_lib/var/name1/libacm_b200.so python tools/profile_run.py ...

The variants live under libacm_b200/_lib/var/ (git-ignored, they travel to the GPU box)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libacm_b200 import build as B  # noqa: E402


def one(name, flags):
    out = os.path.join(B.LIBDIR, "var", name)
    os.makedirs(out, exist_ok=True)
    objs = []
    for s in B.CUDA_SOURCES + B.C_SOURCES:
        src = os.path.join(B.CSRC, s)
        o = os.path.join(out, s + ".o")
        if s.endswith(".cu"):
            cmd = [B._nvcc(), *B.NVCC_FLAGS, *flags, "-I", B.INCLUDE, "-I", B.CSRC, "-c", src, "-o", o]
        else:
            cc = "g++" if s.endswith(".cpp") else "gcc"
            cmd = [cc, "-O2", "-fPIC", *flags, "-I", B.INCLUDE, "-I", B.CSRC, "-c", src, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise SystemExit(f"variant {name}: {s} failed")
        if s == "acm_fast2.cu":
            for ln in (r.stdout + r.stderr).splitlines():
                if "fast2_kernelILb0" in ln or ("registers" in ln and "fast2" in prev):
                    pass
                prev = ln
        objs.append(o)
    so = os.path.join(out, "libacm_b200.so")
    subprocess.check_call([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so, *objs])
    return name, so


if __name__ == "__main__":
    jobs = []
    for a in sys.argv[1:]:
        name, flags = a.split("=", 1)
        jobs.append((name, flags.split()))
    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, so in ex.map(lambda j: one(*j), jobs):
            print(name, so)
