"""BASELINE configs[3] at its full stream count on ONE B200: 1 000 000 stereo streams (125 000 unique
config-4-shaped images, each referenced by 8 descriptors with their own output range: replication by
descriptor), decoded by one acm_gpu_plan_run.  Checks: every stream status 0 with
all its words, the 8 replicas of an image agree in checksum, a sample of images equals the reference."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from libacm_b200 import api  # noqa: E402
from oracle import bindings  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--unique", type=int, default=125_000)
ap.add_argument("--replicas", type=int, default=8)
ap.add_argument("--runs", type=int, default=2)
args = ap.parse_args()

t0 = time.perf_counter()
blob, offs, lens = bench.build_corpus(args.unique, 0, workload="config4")
offs_r = np.tile(offs, args.replicas)
lens_r = np.tile(lens, args.replicas)
n = offs_r.size
opts = api.make_opts(device=0, want_checksums=1)
s = api.new_streams(offs_r, lens_r)
api.probe(blob, s, opts)
nbytes = api.layout(s, 2)
print(f"{n} streams, {s['total_values'].sum() / 1e9:.2f} G samples, blob {blob.size / 1e9:.2f} GB, "
      f"PCM {nbytes / 1e9:.1f} GB, host prep {time.perf_counter() - t0:.1f} s", flush=True)
d_blob = torch.from_numpy(blob).cuda()
d_out = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
t0 = time.perf_counter()
plan = api.Plan(s, opts)
print(f"plan_create {time.perf_counter() - t0:.2f} s", flush=True)
cs = torch.cuda.current_stream().cuda_stream
for _ in range(args.runs):
    plan.run(d_blob, d_out, cs)
    torch.cuda.synchronize()
    ms = plan.last_ms()
    print(f"ms {ms:.2f}  Msamples/s {s['total_values'].sum() / ms / 1e3:.0f}", flush=True)
plan.fetch(s, cs)
plan.close()
assert np.all(s["status"] == 0) and np.array_equal(s["words"], s["total_values"])
ck = s["checksum"].reshape(args.replicas, args.unique)
assert np.all(ck == ck[0]), "replicas of an image disagree"
chk = bindings.best()
for i in range(0, args.unique, args.unique // 50):
    a = chk.decode(blob[int(offs[i]):int(offs[i]) + int(lens[i])])
    assert int(ck[0, i]) == api.checksum_ref(a.pcm, a.words), i
    r = args.replicas - 1
    o = int(s["out_off"][r * args.unique + i])
    assert np.array_equal(d_out[o:o + a.pcm.size].cpu().numpy(), a.pcm), i
print("ok: statuses, replica checksums and the reference sample agree")
