"""Small mixed batch for compute-sanitizer: fast kernel (level 7 / 16 rows), generic kernel, error paths."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from libacm_b200 import gen  # noqa: E402
from tests import corpus, gpu_util as gu  # noqa: E402

plist = corpus.fallout_params(96, seed=3, hi=30_000) + corpus.stress_params(max_values=3000)[::13] + corpus.negative_params()
imgs = corpus.images(plist)
img = gen.make_stream(level=7, rows=16, channels=1, total_values=2048 * 3 + 9, dist=gen.DIST_STRESS, seed=5)
imgs += [img[:c] for c in (20, 300, 1500, len(img) - 2)]
kernel = int(os.environ.get("ACM_KERNEL", "0"))
s, out = gu.decode_device(imgs, want_checksums=1, kernel=kernel)
s2, out2 = gu.decode_host(imgs, align=1, lead=3, kernel=kernel)
print("statuses", sorted(set(s["status"].tolist())), "ok")
