"""ncu target for the auxiliary kernels: generates N config-4-shaped streams on the GPU
(acmgen_size_kernel / acmgen_write_kernel), probes them from the device blob
(acm_gather_headers_kernel) and decodes them once with the generic kernel forced
(acm_decode_generic_kernel).

    ncu --set full --clock-control none --import-source on -k regex:acmgen_write -c 1 -o gpurun_out/prof_gen \
        python tools/prof_aux.py --streams 20000
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from libacm_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=20000)
ap.add_argument("--kernel", type=int, default=1)
args = ap.parse_args()
plist = bench.corpus_params(args.streams, 0, workload="config4")
_, _, used = api.generate_on_device(plist, None)
d_blob = torch.empty(used + 64, dtype=torch.uint8, device="cuda")
offs, lens, used = api.generate_on_device(plist, d_blob.data_ptr(), used + 64)
opts = api.make_opts(device=0, kernel=args.kernel, want_checksums=1)
s = api.new_streams(offs, lens)
api.probe(d_blob, s, opts)
d_out = torch.empty(api.layout(s, 2) + 64, dtype=torch.uint8, device="cuda")
plan = api.Plan(s, opts)
cs = torch.cuda.current_stream().cuda_stream
plan.run(d_blob, d_out, cs)
plan.fetch(s, cs)
print("streams", args.streams, "samples", int(s["total_values"].sum()), "ms", plan.last_ms(),
      "Msamples/s", s["total_values"].sum() / plan.last_ms() / 1e3, "split", plan.split())
assert np.all(s["status"] == 0)
