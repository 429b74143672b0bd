"""Debug aid: prints the mismatch list of the unaligned back-to-back test."""
import sys
sys.path.insert(0, '.')
from oracle import bindings
from tests import corpus, gpu_util as gu

checker = bindings.best()
imgs = corpus.images(corpus.stress_params(max_values=8_000)[::2] + corpus.fallout_params(40, seed=9, hi=30_000))
for lead in (0, 1, 2, 3):
    s, out = gu.decode_host(imgs, align=1, lead=lead, kernel=0)
    bad = gu.compare(imgs, s, out, checker)
    print("lead", lead, "bad", bad[:8], len(bad))
