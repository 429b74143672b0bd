"""Small driver for ncu captures: decodes N config-2-shaped streams a few times.

    ncu --set full --clock-control none --import-source on -k regex:acm_decode_fast -s 1 -c 1 \
        -o gpurun_out/prof python tools/profile_run.py --streams 2000 --runs 2
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
ap.add_argument("--streams", type=int, default=2000)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--workload", default="config2")
args = ap.parse_args()

blob, offs, lens = bench.build_corpus(args.streams, 0, workload=args.workload)
opts = api.make_opts(device=0, kernel=args.kernel)
s = api.new_streams(offs, lens)
d_blob = torch.from_numpy(blob).cuda()
api.probe(blob, s, opts)
nbytes = api.layout(s, 2)
d_out = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
plan = api.Plan(s, opts)
cs = torch.cuda.current_stream().cuda_stream
for _ in range(args.runs):
    plan.run(d_blob, d_out, cs)
    torch.cuda.synchronize()
    print("ms", plan.last_ms(), "Msamples/s", s["total_values"].sum() / plan.last_ms() / 1e3)
if os.environ.get("WALK_PROF"):
    import ctypes as C
    buf = (C.c_uint64 * 64)()
    api.lib().acm_gpu_plan_debug_counters.argtypes = [C.c_void_p, C.c_void_p]
    api.lib().acm_gpu_plan_debug_counters(plan._h, buf)
    print("walk warp 0: cycles retire %d topup %d header %d steps %d looptail %d periods %d outer %d" % tuple(buf[40:47]))
if os.environ.get("F2_PROF"):
    import ctypes as C
    buf = (C.c_uint64 * 64)()
    api.lib().acm_gpu_plan_debug_counters.argtypes = [C.c_void_p, C.c_void_p]
    api.lib().acm_gpu_plan_debug_counters(plan._h, buf)
    v = [x / args.runs / 1e6 for x in buf]
    print("scan  Mcycles/run (all warps): publish %.1f flow %.1f hyst %.1f head %.1f steps %.1f topup %.1f" % tuple(v[0:6]))
    print("scan  busiest warp %.2f Mcycles, max rounds/warp %.0f, rounds %.0f, periods %.0f (x runs)" %
          (buf[16] / 1e6, buf[17], buf[18] / args.runs, buf[19] / args.runs))
    print("work  Mcycles/run (all warps): decode %.1f idle %.1f claim %.1f" % tuple(v[8:11]))
    print("work  phases Mcycles/run: record %.1f stage-wait %.1f unpack %.1f copy+dequant %.1f transform %.1f rest %.1f" % tuple(v[24:30]))
plan.fetch(s, cs)
assert np.all(s["status"] == 0)
# digest of the PCM (int64 wrap-around sum of the output's 8-byte words) and of the word counts: equal across variants
n8 = nbytes // 8
print("groups", plan.gen_groups(), "words", int(s["words"].sum()), "digest", int(d_out[:n8 * 8].view(torch.int64).sum().item()))
