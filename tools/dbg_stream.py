import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libacm_b200 import gen
from tests import api_driver as ad
img = gen.make_stream(level=7, rows=16, channels=2, rate=22050, total_values=2_646_000, dist=gen.DIST_FALLOUT, seed=1)
lib = ad.mine()
for rep in range(3):
    t0 = time.perf_counter()
    h = ad.Handle(lib, img)
    t1 = time.perf_counter()
    marks = []
    n = 0
    while True:
        ta = time.perf_counter()
        r, _ = h.read(8192, loop=True)
        tb = time.perf_counter()
        if tb - ta > 2e-4:
            marks.append((n // 2, round((tb - ta) * 1e3, 2)))
        if r <= 0:
            break
        n += r
    t2 = time.perf_counter()
    h.close()
    t3 = time.perf_counter()
    print("open %.2f ms, reads %.2f ms, close %.2f ms; slow reads (word, ms): %s" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, marks[:14]))
