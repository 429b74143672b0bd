"""The libacm.h drop-in surface, timed on this library and on the reference (one host core):
  config 1: one 60 s stereo 22 050 Hz stream (level 7, 16 rows): open + acm_read_loop(8192 bytes per call,
            acmtool's pattern) + close
  config 5: one 5 min stereo 44 100 Hz stream: open, a first forward acm_seek_pcm to the middle, one read.
Prints one JSON line (bench.py's `streaming` block is made of it)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libacm_b200 import gen  # noqa: E402
from tests import api_driver as ad  # noqa: E402


def run(reps=4, seek=True):
    img = gen.make_stream(level=7, rows=16, channels=2, rate=22050, total_values=2_646_000, dist=gen.DIST_FALLOUT, seed=1)
    long_img = gen.make_stream(level=7, rows=16, channels=2, rate=44100, total_values=26_460_000,
                               dist=gen.DIST_FALLOUT, seed=2) if seek else None
    libs = [("gpu", ad.mine())] + ([("reference", ad.ref())] if ad.have_ref() else [])
    out = {}
    for name, lib in libs:
        best = None
        n = 0
        for _ in range(reps):
            t0 = time.perf_counter()
            h = ad.Handle(lib, img)
            n = 0
            while True:
                r, _ = h.read(8192, loop=True)
                if r <= 0:
                    break
                n += r
            h.close()
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        out[name] = {"read_loop_msamples_s": round(n / 2 / best / 1e6, 1), "read_loop_ms": round(best * 1e3, 2)}
        if seek:
            sbest, pos = None, 0
            for _ in range(3):
                h = ad.Handle(lib, long_img)
                t0 = time.perf_counter()
                pos = h.seek(26_460_000 // 4)  # pcm frames: the middle of the stream
                r, _ = h.read(4096)
                dt = time.perf_counter() - t0
                h.close()
                sbest = dt if sbest is None or dt < sbest else sbest
            out[name]["seek_middle_ms"] = round(sbest * 1e3, 2)
            out[name]["seek_pos"] = int(pos)
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
