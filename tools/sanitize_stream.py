"""One stream through the libacm.h surface for compute-sanitizer: the one-stream walker (acm_walk1_kernel),
the split path's unpack and lift on a chunk, a forward seek (skip-ahead with two chunks in flight) and a
truncated stream (the walker's re-walk with the reference's verdicts)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libacm_b200 import gen  # noqa: E402
from tests import api_driver as ad  # noqa: E402

lib = ad.mine()
img = gen.make_stream(level=7, rows=16, channels=2, rate=22050, total_values=2048 * 700 + 77, dist=gen.DIST_FALLOUT, seed=1)
for image in (img, img[: len(img) // 2 + 13]):
    h = ad.Handle(lib, image)
    n = 0
    for _ in range(40):
        r, _d = h.read(8192, loop=True)
        if r <= 0:
            break
        n += r
    pos = h.seek(2048 * 300 // 2 + 5)
    r, _d = h.read(4096)
    while True:
        r, _d = h.read(65536, loop=True)
        if r <= 0:
            break
    print("read", n, "seek ->", pos, "last", r)
    h.close()
