"""Config 1 (one 60 s stereo 22 050 Hz stream, level 7, 16 rows) through the libacm.h drop-in surface:
open + acm_read_loop(8192 bytes per call, acmtool's pattern) + close, on this library and on the reference."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libacm_b200 import gen  # noqa: E402
from tests import api_driver as ad  # noqa: E402

img = gen.make_stream(level=7, rows=16, channels=2, rate=22050, total_values=2_646_000, dist=gen.DIST_FALLOUT, seed=1)
libs = [("libacm_b200 (GPU)", ad.mine())] + ([("reference (CPU, 1 core)", ad.ref())] if ad.have_ref() else [])
for name, lib in libs:
    best = None
    for rep in range(4):
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
    print(f"{name}: {n // 2} words in {best * 1e3:.1f} ms = {n / 2 / best / 1e6:.1f} Msamples/s")
