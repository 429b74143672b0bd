"""BASELINE configs[3] at its full stream count on ONE B200: 1 000 000 stereo streams (125 000 unique
config-4-shaped images, each referenced by 8 descriptors with their own output range: replication by
descriptor), decoded by one acm_gpu_plan_run.  Checks: every stream status 0 with
all its words, the 8 replicas of an image agree in checksum, a sample of images equals the reference.

With --device-gen the corpus never exists on the host: unique x replicas DIFFERENT images are generated
in HBM by acm_gpu_generate (the host generator compiled as device code), the
headers are probed on the device, and a sample of images is copied back and compared with the reference."""
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
ap.add_argument("--device-gen", action="store_true")
args = ap.parse_args()

if args.device_gen:
    n = args.unique * args.replicas
    t0 = time.perf_counter()
    plist = bench.corpus_params(n, 0, workload="config4")
    t1 = time.perf_counter()
    _, _, used = api.generate_on_device(plist, None)
    d_blob = torch.empty(used + 64, dtype=torch.uint8, device="cuda")
    offs, lens, used = api.generate_on_device(plist, d_blob.data_ptr(), used + 64)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    opts = api.make_opts(device=0, want_checksums=1)
    s = api.new_streams(offs, lens)
    api.probe(d_blob, s, opts)
    nbytes = api.layout(s, 2)
    print(f"{n} unique streams, {s['total_values'].sum() / 1e9:.2f} G samples, blob {used / 1e9:.2f} GB generated in HBM in "
          f"{t2 - t1:.2f} s (sizing pass + write pass; parameter records {t1 - t0:.1f} s on the host), PCM {nbytes / 1e9:.1f} GB",
          flush=True)
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
    chk = bindings.best()
    for i in range(0, n, n // 60):
        img = d_blob[int(offs[i]):int(offs[i]) + int(lens[i])].cpu().numpy()
        a = chk.decode(img)
        assert int(s["checksum"][i]) == api.checksum_ref(a.pcm, a.words), i
        o = int(s["out_off"][i])
        assert np.array_equal(d_out[o:o + a.pcm.size].cpu().numpy(), a.pcm), i
    print("ok: statuses and the reference sample agree")
    sys.exit(0)

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
