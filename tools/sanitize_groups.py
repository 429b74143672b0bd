"""Small general-path batch cut into stream groups (one CUDA stream each) for compute-sanitizer:
healthy, truncated and corrupt streams of mixed shapes, the plan run twice."""
import os
import sys

os.environ.setdefault("ACM_B200_GEN_GROUP_MIN", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from libacm_b200 import api, gen  # noqa: E402
from tests import corpus, gpu_util as gu  # noqa: E402

rng = np.random.default_rng(5)
plist = []
for k in range(120):
    level, rows = int(rng.integers(0, 9)), int(rng.choice([1, 3, 4, 8, 16, 32]))
    blen = rows << level
    plist.append(gen.params(level=level, rows=rows, channels=1 + k % 2, total_values=int(rng.integers(blen + 1, max(2 * blen, 6000))),
                            dist=gen.DIST_STRESS, seed=900 + k))
imgs = corpus.images(plist + corpus.negative_params())
imgs += [img[:len(img) // 2] for img in imgs[:20]]
blob, offs, lens = gu.pack(imgs, align=1, lead=1)
opts = api.make_opts(want_checksums=1)
d_blob = torch.from_numpy(blob).cuda()
s = api.new_streams(offs, lens)
api.probe(d_blob, s, opts)
nbytes = api.layout(s, 2)
plan = api.Plan(s, opts)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    d_out = torch.zeros(nbytes + 16, dtype=torch.uint8, device="cuda")
    plan.run(d_blob, d_out, st)
    plan.fetch(s, st)
print("groups", plan.gen_groups(), "statuses", sorted(set(s["status"].tolist())), "ok")
plan.close()
