import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import corpus, gpu_util as gu
from oracle import bindings
checker = bindings.best()
plist = corpus.stress_params(max_values=8_000)[::2] + corpus.fallout_params(40, seed=9, hi=30_000)
imgs = corpus.images(plist)
for lead in (0,):
    s, out = gu.decode_host(imgs, align=1, lead=lead, kernel=2)
    for i in (180, 188):
        a = checker.decode(imgs[i])
        o = int(s["out_off"][i])
        got = out[o:o + a.info.total_values * 2].view(np.int16)
        exp = a.pcm.view(np.int16)
        d = np.flatnonzero(got != exp)
        print("stream", i, "n", len(exp), "mismatch", len(d), "first", d[:8], "last", d[-3:])
        print(" got", got[:12], "\n exp", exp[:12])
        blk = d // 2048
        print(" blocks with mismatches:", np.unique(blk)[:20])
        col = d % 128
        print(" cols with mismatches:", np.unique(col)[:40])
