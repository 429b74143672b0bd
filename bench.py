#!/usr/bin/env python
"""bench.py -- batched ACM decode throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU decoder

Workload (config.workload = BASELINE configs[1]): 10 000 synthetic Fallout-style mono
22 050 Hz clips, level 7, 16 rows, 1-10 s each (about 1.2 G PCM words) PER GPU -- weak
scaling: every rank decodes its own 10k-stream shard, no data-path collective; the only
collective is an all_gather of the per-stream checksums (NCCL) after the timed region.

A "step" decodes the whole shard once.  `value` is PCM Msamples/s with blob and PCM
resident in HBM (acm_gpu_plan_run, CUDA events, max over ranks); `e2e` is the same
metric through acm_gpu_decode_batch with pinned HOST buffers (H2D of the compressed
images and D2H of all PCM inside the timed region).  The working set per step
(~0.5 GB in, ~2.4 GB out) is far larger than the 126 MB L2, so no flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STREAMS = 10_000
WORKLOAD = ("batch of 10k synthetic Fallout-style mono 22050 Hz ACM clips "
            "(level 7, 16 rows, 1-10 s) per GPU via acm_gpu_decode_batch")
HBM_FALLBACK_GBS = 6650.0


# --------------------------------------------------------------------------- helpers

def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def corpus_params(n_streams, rank, workload="config2"):
    """Generator parameters of the synthetic workload (one record per stream)."""
    from libacm_b200 import gen
    rng = np.random.default_rng(1234 + rank)
    if workload == "config4":
        # BASELINE configs[3] shape: mixed-length stereo streams, log-uniform 0.05-5 s at 22050 Hz
        dur = np.exp(rng.uniform(np.log(0.05), np.log(5.0), size=n_streams))
        tv = (2 * np.round(22050 * dur)).astype(np.int64)
        ch = 2
    else:
        tv = rng.integers(22050, 220500 + 1, size=n_streams)
        ch = 1
    plist = [gen.params(level=7, rows=16, channels=ch, rate=22050, total_values=int(t),
                        dist=gen.DIST_FALLOUT, seed=(rank << 32) + 17 * i + 1)
             for i, t in enumerate(tv)]
    return plist


def build_corpus(n_streams, rank, threads=None, workload="config2"):
    from libacm_b200 import gen
    blob, offs, lens = gen.make_batch(corpus_params(n_streams, rank, workload), threads=threads)
    return blob, offs, lens


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed
    region runs (the nvidia-smi clocks line of the profiling recipe, without a fork)."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_decode_rate(blob, offs, lens, idx, threads):
    """Decode the streams `idx` with the CPU checker on `threads` host threads.
    Returns (Msamples/s, kind, words, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import bindings
    chk = bindings.best()
    chunks = np.array_split(np.asarray(idx), threads * 4)

    def work(ch):
        w = 0
        for i in ch:
            o, l = int(offs[i]), int(lens[i])
            _, words = chk.time_decode(blob[o:o + l], 1)
            w += words
        return w

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        words = sum(ex.map(work, chunks))
    dt = time.perf_counter() - t0
    return words / dt / 1e6, chk.kind, words, dt


# --------------------------------------------------------------------------- reference arm

def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_sample = min(N_STREAMS, args.ref_streams)
    blob, offs, lens = build_corpus(n_sample, 0)
    idx = np.arange(n_sample)
    for _ in range(args.warmup):
        cpu_decode_rate(blob, offs, lens, idx[: max(threads * 4, 64)], threads)
    rates, words, secs = [], 0, 0.0
    kind = "port"
    for _ in range(args.steps):
        r, kind, w, dt = cpu_decode_rate(blob, offs, lens, idx, threads)
        rates.append(r)
        words += w
        secs += dt
    value = words / secs / 1e6
    sample = (f"{n_sample} of the {N_STREAMS} streams of the workload per step "
              f"({words // max(1, args.steps)} words), in-memory images, {threads} threads, one decode each")
    line = {
        "impl": "reference", "metric": "batched decode PCM Msamples/s", "value": round(value, 2),
        "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * secs / max(1, args.steps), 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "streams_per_gpu": N_STREAMS, "cpu_sample_streams": n_sample},
        "cpu_baseline": {"value": round(value, 2), "unit": "Msamples/s", "cores": threads,
                         "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "Msamples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    from libacm_b200 import api

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    t_gen = time.perf_counter()
    threads = max(1, (os.cpu_count() or 1) // max(1, world))
    blob, offs, lens = build_corpus(args.streams, rank, threads=threads, workload=args.workload)
    t_gen = time.perf_counter() - t_gen

    opts = api.make_opts(device=local, want_checksums=0)
    streams = api.new_streams(offs, lens)
    # pinned host copies (e2e path) and device-resident copies (kernel path)
    h_blob = torch.from_numpy(blob).pin_memory()
    d_blob = h_blob.to(dev, non_blocking=True)
    api.probe(h_blob.numpy(), streams, opts)
    out_bytes = api.layout(streams, 2)
    d_out = torch.empty(out_bytes + 64, dtype=torch.uint8, device=dev)
    total_words = int(streams["total_values"].sum())
    in_bytes = int(lens.astype(np.int64).sum())
    algo_bytes = in_bytes + 2 * total_words

    plan = api.Plan(streams, opts)
    n_fast, n_generic = plan.split()
    cs = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        plan.run(d_blob, d_out, cs)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record()
    for _ in range(args.steps):
        plan.run(d_blob, d_out, cs)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms.append(plan.last_ms())
    # correctness gate on the timed output: statuses, word counts, checksum of checksums
    chk_opts = api.make_opts(device=local, want_checksums=1)
    plan_c = api.Plan(streams, chk_opts)
    plan_c.run(d_blob, d_out, cs)
    plan_c.fetch(streams, cs)
    plan_c.close()
    ok = bool(np.all(streams["status"] == 0) and np.array_equal(streams["words"], streams["total_values"]))
    checksums = streams["checksum"].copy()

    # ---- e2e: host buffers through the one-shot C ABI
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if args.no_e2e:
        e2e_s = 0.0
    else:
        del d_out
        torch.cuda.empty_cache()
        h_out = torch.empty(out_bytes + 64, dtype=torch.uint8).pin_memory()
        api.decode_batch(h_blob.numpy(), streams, h_out.numpy(), opts)  # warm-up (allocations, page-in)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            api.decode_batch(h_blob.numpy(), streams, h_out.numpy(), opts)
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps

    # ---- reductions over ranks: max time, sum of work
    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    w = torch.tensor([float(total_words), float(algo_bytes), float(in_bytes), float(ok)],
                     dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        # optional result gather (the only collective; NCCL, outside the timed region)
        from libacm_b200 import shard
        base = rank * args.streams
        _, g_words, g_cks = shard.gather_results(np.arange(base, base + args.streams), streams["status"],
                                                 streams["words"], checksums, world * args.streams, device=dev)
        ck_of_ck = int(np.sum(g_cks, dtype=np.uint64))
    else:
        ck_of_ck = int(np.sum(checksums.astype(np.uint64), dtype=np.uint64))
    ms_total, e2e_ms = float(t[0]), float(t[1])
    words_all, bytes_all, in_all, ok_all = float(w[0]), float(w[1]), float(w[2]), float(w[3])
    ms_per_step = ms_total / args.steps
    value = words_all / (ms_per_step * 1e-3) / 1e6
    e2e_value = words_all / (e2e_ms * 1e-3) / 1e6 if e2e_ms > 0 else 0.0

    line = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        # one launch per step on this rank: per-launch figures are rank 0's own shard
        launch_ms = ms_per_step
        achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.workload == "config2" and args.streams == N_STREAMS:
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu:
            n_sample = min(args.streams, args.ref_streams)
            threads_cpu = os.cpu_count() or 1
            r, kind, wds, dt = cpu_decode_rate(blob, offs, lens, np.arange(n_sample), threads_cpu)
            cpu = {"value": round(r, 2), "unit": "Msamples/s", "cores": threads_cpu, "kind": kind,
                   "sample": f"first {n_sample} streams of the workload ({wds} words, {dt:.2f} s wall), "
                             f"in-memory images, one decode each"}
        line = {
            "metric": "batched decode PCM Msamples/s", "value": round(value, 1), "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.workload == "config2" else
                       "mixed-length synthetic stereo 22050 Hz streams (level 7, 16 rows, log-uniform 0.05-5 s), "
                       "sharded by stream (BASELINE configs[3] shape)",
                       "streams_per_gpu": args.streams,
                       "words_per_gpu": total_words, "compressed_bytes_per_gpu": in_bytes,
                       "bits_per_sample": round(8.0 * in_bytes / total_words, 3),
                       "format": "s16le", "l2_policy": "working set 3 GB >> 126 MB L2, no flush",
                       "kernel_split": {"fast": n_fast, "generic": n_generic},
                       "parallelism": f"shard-by-stream x{world}, no data-path collective",
                       "corpus_gen_s": round(t_gen, 2)},
            "parity_gate": {"all_status_ok": bool(ok_all == world), "checksum_of_checksums": ck_of_ck},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "launch_ms": round(launch_ms, 4), "plan_last_ms": round(kernel_ms[-1], 4)},
            "cpu_baseline": cpu,
            "e2e": None if args.no_e2e else
                   {"value": round(e2e_value, 1), "unit": "Msamples/s",
                    "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": int(out_bytes),
                    "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps},
            "gpu_launches": plan.launches * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    plan.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=N_STREAMS)
    ap.add_argument("--ref-streams", type=int, default=N_STREAMS,
                    help="streams of the workload the CPU legs decode per step")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="config2", choices=["config2", "config4"],
                    help="config2 = BASELINE configs[1] (the headline); config4 = configs[3] shape per GPU")
    args = ap.parse_args()
    from libacm_b200 import build
    rank, _, _ = dist_env()
    if rank == 0:
        build.ensure_built()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
