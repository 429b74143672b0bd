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


def config4_lengths(n_streams, rank=0):
    """total_values of the BASELINE configs[3] streams: stereo, log-uniform 0.05-5 s at 22050 Hz."""
    rng = np.random.default_rng(1234 + rank)
    dur = np.exp(rng.uniform(np.log(0.05), np.log(5.0), size=n_streams))
    return (2 * np.round(22050 * dur)).astype(np.int64)


def corpus_params(n_streams, rank, workload="config2"):
    """Generator parameters of the synthetic workload (one record per stream)."""
    from libacm_b200 import gen
    rng = np.random.default_rng(1234 + rank)
    if workload == "config3":
        # BASELINE configs[2]: filler-stress streams (random per-column fill types: every f_k / f_t / f_linear
        # path), levels 4-10, rows 4-32, plain and WAVC headers, mono and stereo, 1-10 s at 22050 Hz
        tv = rng.integers(22050, 220500 + 1, size=n_streams)
        lv = rng.integers(4, 11, size=n_streams)
        rw = rng.choice(np.array([4, 8, 16, 32]), size=n_streams)
        chs = rng.integers(1, 3, size=n_streams)
        wv = rng.integers(0, 2, size=n_streams)
        return [gen.params(level=int(lv[i]), rows=int(rw[i]), channels=int(chs[i]), rate=22050,
                           total_values=int(tv[i]) * int(chs[i]), wavc=int(wv[i]), dist=gen.DIST_STRESS,
                           seed=(rank << 32) + 17 * i + 1)
                for i in range(n_streams)]
    if workload == "config4":
        # BASELINE configs[3] shape: mixed-length stereo streams, log-uniform 0.05-5 s at 22050 Hz
        tv = config4_lengths(n_streams, rank)
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


def acmtool_xargs_rate(blob, offs, lens, idx):
    """BASELINE.md section 3 as written: the reference's own CLI in its benchmark mode, one process per
    host core over the streams `idx` written as individual files on tmpfs:
        ls corpus/*.acm | xargs -P $(nproc) -n 64 oracle/_ref/acmtool -d -n -q
    Returns a dict (Msamples/s total and per core), or None when the reference CLI is not built."""
    import shutil
    import subprocess
    import tempfile
    tool = os.path.join(ROOT, "oracle", "_ref", "acmtool")
    if not os.path.exists(tool):
        return None
    nproc = os.cpu_count() or 1
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="acm_bench_", dir=base)
    try:
        words = 0
        names = []
        for i in idx:
            o, l = int(offs[i]), int(lens[i])
            fn = os.path.join(d, f"{int(i):07d}.acm")
            with open(fn, "wb") as f:
                f.write(blob[o:o + l].tobytes())
            names.append(fn)
            words += int(blob[o + 4]) | int(blob[o + 5]) << 8 | int(blob[o + 6]) << 16 | int(blob[o + 7]) << 24
        inp = "\n".join(names).encode()
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            r = subprocess.run(["xargs", "-P", str(nproc), "-n", "64", tool, "-d", "-n", "-q"], input=inp,
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return None
            best = dt if best is None else min(best, dt)
        return {"value": round(words / best / 1e6, 2), "unit": "Msamples/s", "per_core": round(words / best / 1e6 / nproc, 2),
                "nproc": nproc, "streams": len(names), "wall_s": round(best, 3),
                "command": "ls corpus/*.acm | xargs -P $(nproc) -n 64 oracle/_ref/acmtool -d -n -q  (files on tmpfs)"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def oracle_check(chk, api, d_blob, d_out, streams, offs, lens, picks):
    """Decode the streams `picks` (local indices) with the CPU checker and compare status, words, checksum and
    PCM bytes with what the GPU left in d_out / streams.  Returns (checked, failures)."""
    bad = 0
    for i in picks:
        o, l = int(offs[i]), int(lens[i])
        img = d_blob[o:o + l].cpu().numpy()
        a = chk.decode(img)
        p0 = int(streams["out_off"][i])
        same = (int(streams["status"][i]), int(streams["words"][i])) == (a.status, a.words)
        same = same and int(streams["checksum"][i]) == api.checksum_ref(a.pcm, a.words)
        same = same and bool(np.array_equal(d_out[p0:p0 + a.pcm.size].cpu().numpy(), a.pcm))
        bad += 0 if same else 1
    return len(picks), bad


def run_config4(args, rank, world, local, dev, use_dist, peak):
    """BASELINE configs[3] as written: ONE corpus of args.config4_streams mixed-length stereo streams (all
    different: stream i has seed 17 i + 1), cut into `world` shards by libacm_b200.shard.partition (LPT on
    total_values), every rank generates ITS streams in HBM (acm_gpu_generate) and decodes them; no
    data-path collective.  Afterwards the per-stream results are all_gathered (NCCL) and a sample of every
    rank's shard is decoded by the reference on the host and compared (status, words, checksum, bytes)."""
    import torch
    from libacm_b200 import api, gen, shard
    from oracle import bindings
    n_total = args.config4_streams
    t0 = time.perf_counter()
    tv = config4_lengths(n_total)
    mine = shard.partition(tv, world)[rank]
    plist = [gen.params(level=7, rows=16, channels=2, rate=22050, total_values=int(tv[i]),
                        dist=gen.DIST_FALLOUT, seed=17 * int(i) + 1) for i in mine]
    _, _, used = api.generate_on_device(plist, None, device=local)
    d_blob = torch.empty(used + 64, dtype=torch.uint8, device=dev)
    offs, lens, used = api.generate_on_device(plist, d_blob.data_ptr(), used + 64, device=local)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    opts = api.make_opts(device=local, want_checksums=0)
    s = api.new_streams(offs, lens)
    api.probe(d_blob, s, opts)
    out_bytes = api.layout(s, 2)
    d_out = torch.empty(out_bytes + 64, dtype=torch.uint8, device=dev)
    words = int(s["total_values"].sum())
    in_bytes = int(lens.astype(np.int64).sum())
    plan = api.Plan(s, opts)
    cs = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        plan.run(d_blob, d_out, cs)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, args.config4_steps)
    ev0.record()
    for _ in range(steps):
        plan.run(d_blob, d_out, cs)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / steps
    plan.close()
    # results with checksums (the <true> kernel instance), then the reference on a sample of THIS shard
    copts = api.make_opts(device=local, want_checksums=1)
    plan_c = api.Plan(s, copts)
    plan_c.run(d_blob, d_out, cs)
    plan_c.fetch(s, cs)
    plan_c.close()
    ok = bool(np.all(s["status"] == 0) and np.array_equal(s["words"], s["total_values"])
              and np.array_equal(s["total_values"], tv[mine].astype(np.uint32)))
    chk = bindings.best()
    k = min(len(mine), args.config4_check)
    picks = np.linspace(0, len(mine) - 1, k).astype(np.int64) if k else []
    checked, failures = oracle_check(chk, api, d_blob, d_out, s, offs, lens, picks)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    w = torch.tensor([float(words), float(in_bytes + 2 * words), float(ok), float(checked), float(failures),
                      float(len(mine))], dtype=torch.float64, device=dev)
    ck_of_ck = int(np.sum(s["checksum"].astype(np.uint64), dtype=np.uint64))
    if use_dist:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        g_status, g_words, g_cks = shard.gather_results(mine, s["status"], s["words"], s["checksum"], n_total, device=dev)
        ok_gather = bool(np.all(g_status == 0) and np.array_equal(g_words, tv.astype(np.uint32)))
        ck_of_ck = int(np.sum(g_cks, dtype=np.uint64))
    else:
        ok_gather = ok
    del d_blob, d_out
    torch.cuda.empty_cache()
    ms = float(t[0])
    words_all, bytes_all = float(w[0]), float(w[1])
    return {
        "streams": n_total, "unique_streams": n_total, "streams_on_rank0": int(len(mine)),
        "partition": "libacm_b200.shard.partition (LPT on total_values), one shard per rank",
        "corpus": "generated in HBM by acm_gpu_generate, stream i: seed 17 i + 1, stereo, log-uniform 0.05-5 s",
        "words": int(words_all), "ms_per_step": round(ms, 3), "steps": steps,
        "gsamples_s": round(words_all / (ms * 1e-3) / 1e9, 2),
        "frac": round(bytes_all / (ms * 1e-3) / 1e9 / (peak * world), 4),
        "all_status_ok": bool(w[2] == world) and ok_gather,
        "oracle_checked": int(w[3]), "oracle_failures": int(w[4]), "oracle_kind": chk.kind,
        "gathered": "all_gather of (index, status, words, checksum) per stream over NCCL" if use_dist else "single rank",
        "checksum_of_checksums": ck_of_ck, "gen_s_rank0": round(t_gen, 2),
    }


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
    xargs = None if args.no_acmtool else acmtool_xargs_rate(blob, offs, lens, idx[: min(n_sample, args.acmtool_streams)])
    sample = (f"{n_sample} of the {N_STREAMS} streams of the workload per step "
              f"({words // max(1, args.steps)} words), in-memory images, {threads} threads, one decode each")
    line = {
        "impl": "reference", "metric": "batched decode PCM Msamples/s", "value": round(value, 2),
        "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * secs / max(1, args.steps), 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "streams_per_gpu": N_STREAMS, "cpu_sample_streams": n_sample},
        "cpu_baseline": {"value": round(value, 2), "unit": "Msamples/s", "cores": threads,
                         "kind": kind, "sample": sample,
                         "how": "in-process: one host thread per core, each running acm_open_decoder + acm_read_loop "
                                "(acmtool's 8 KiB requests) over in-memory images; acmtool_xargs is BASELINE.md "
                                "section 3 as written (one acmtool -d -n -q process per core over files on tmpfs)",
                         "acmtool_xargs": xargs},
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
    n_launches = plan.launches
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
    # correctness gate.  The timed launches run the kernel instance without checksums; its output is kept,
    # the instance with checksums decodes the batch again, and the two outputs must be the same bytes.  The
    # checksummed run is what the reference is compared with: statuses, word counts, and -- on EVERY rank --
    # status / words / checksum / PCM bytes of a sample of that rank's own streams decoded on the host.
    d_timed = d_out.clone()
    chk_opts = api.make_opts(device=local, want_checksums=1)
    plan_c = api.Plan(streams, chk_opts)
    plan_c.run(d_blob, d_out, cs)
    plan_c.fetch(streams, cs)
    plan_c.close()
    timed_equal = bool(torch.equal(d_timed, d_out))
    del d_timed
    ok = bool(np.all(streams["status"] == 0) and np.array_equal(streams["words"], streams["total_values"])) and timed_equal
    checksums = streams["checksum"].copy()
    from oracle import bindings
    chk = bindings.best()
    picks = np.linspace(0, args.streams - 1, min(args.streams, args.oracle_check)).astype(np.int64)
    n_checked, n_failed = oracle_check(chk, api, d_blob, d_out, streams, offs, lens, picks)
    ok = ok and n_failed == 0

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
        # what the copies alone cost on this box with all ranks copying at the same time: the floor of the
        # e2e step (PCIe / host memory, not kernels).  Same pinned buffers, same byte counts.
        d_tmp = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
        d_in = torch.empty(h_blob.numel(), dtype=torch.uint8, device=dev)
        copy_ms = []
        for direction in ("d2h", "h2d"):
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                if direction == "d2h":
                    h_out[:out_bytes].copy_(d_tmp, non_blocking=True)
                else:
                    d_in.copy_(h_blob, non_blocking=True)
            barrier()
            copy_ms.append((time.perf_counter() - t0) / 3 * 1e3)
        del d_tmp, d_in

    if args.no_e2e:
        copy_ms = [0.0, 0.0]
    # ---- reductions over ranks: max time, sum of work
    t = torch.tensor([ms_total, e2e_s * 1e3, copy_ms[0], copy_ms[1]], dtype=torch.float64, device=dev)
    w = torch.tensor([float(total_words), float(algo_bytes), float(in_bytes), float(ok), float(n_checked),
                      float(n_failed), float(timed_equal)], dtype=torch.float64, device=dev)
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
    checked_all, failed_all, timed_equal_all = int(w[4]), int(w[5]), int(w[6])
    peak, peak_src = hbm_peak()
    config4 = None
    if not args.no_config4:
        plan.close()
        del d_blob
        if args.no_e2e:
            del d_out
        torch.cuda.empty_cache()
        config4 = run_config4(args, rank, world, local, dev, use_dist, peak)
    ms_per_step = ms_total / args.steps
    value = words_all / (ms_per_step * 1e-3) / 1e6
    e2e_value = words_all / (e2e_ms * 1e-3) / 1e6 if e2e_ms > 0 else 0.0

    line = None
    if rank == 0:
        # one launch per step on this rank: per-launch figures are rank 0's own shard
        launch_ms = ms_per_step
        achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.workload == "config2" and args.streams == N_STREAMS:
            try:
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = ("profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full "
                               "capture of this kernel on this workload (" + str(tj.get("kernel_version", "?")) + ", " +
                               str(tj.get("source", "?")) + "); not measured in this run")
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
        streaming = None
        if world == 1 and not args.no_streaming:
            # BASELINE configs[0] / configs[4]: ONE stream through the libacm.h drop-in surface, beside the
            # reference on one host core (same C harness on both sides; tools/stream_time.py)
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import stream_time
                streaming = stream_time.run(reps=3, seek=True)
                streaming["what"] = ("config 1: open + acm_read_loop(8 KiB requests) + close of one 60 s stereo 22 050 Hz "
                                     "stream; config 5: first forward acm_seek_pcm to the middle of a 5 min stereo 44 100 Hz "
                                     "stream + one read; best of 3; reference = unmodified libacm on one host core")
            except Exception as e:  # the bench line must not depend on it
                streaming = {"error": repr(e)}
        line = {
            "metric": "batched decode PCM Msamples/s", "value": round(value, 1), "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.workload == "config2" else
                       "filler-stress corpus (BASELINE configs[2]): random per-column fill types, levels 4-10, rows 4-32, "
                       "plain and WAVC headers, mono and stereo, 1-10 s at 22050 Hz, via acm_gpu_decode_batch"
                       if args.workload == "config3" else
                       "mixed-length synthetic stereo 22050 Hz streams (level 7, 16 rows, log-uniform 0.05-5 s), "
                       "sharded by stream (BASELINE configs[3] shape)",
                       "streams_per_gpu": args.streams,
                       "words_per_gpu": total_words, "compressed_bytes_per_gpu": in_bytes,
                       "bits_per_sample": round(8.0 * in_bytes / total_words, 3),
                       "format": "s16le", "l2_policy": "working set 3 GB >> 126 MB L2, no flush",
                       "kernel_split": {"fast": n_fast, "generic": n_generic},
                       "parallelism": f"shard-by-stream x{world}, no data-path collective",
                       "corpus_gen_s": round(t_gen, 2)},
            "parity_gate": {"all_status_ok": bool(ok_all == world), "checksum_of_checksums": ck_of_ck,
                            "timed_output_equals_checked_output": bool(timed_equal_all == world),
                            "oracle_checked_streams": checked_all, "oracle_failures": failed_all,
                            "oracle_kind": chk.kind, "oracle_on_every_rank": True},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "launch_ms": round(launch_ms, 4), "plan_last_ms": round(kernel_ms[-1], 4)},
            "cpu_baseline": cpu,
            "e2e": None if args.no_e2e else
                   {"value": round(e2e_value, 1), "unit": "Msamples/s",
                    "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": int(out_bytes),
                    "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps,
                    "copy_floor": {"d2h_only_ms": round(float(t[2]), 3), "h2d_only_ms": round(float(t[3]), 3),
                                   "d2h_gbs_aggregate": round(world * out_bytes / max(float(t[2]), 1e-9) / 1e6, 1),
                                   "what": "plain cudaMemcpyAsync of the same pinned buffers, all ranks at once, max over "
                                           "ranks: the step cannot be faster than the slower of the two"}},
            "gpu_launches": n_launches * args.steps,
            "clocks": clocks,
            "config4": config4,
            "streaming": streaming,
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
    ap.add_argument("--no-config4", action="store_true", help="skip the 1M-stream sharded run (BASELINE configs[3])")
    ap.add_argument("--config4-streams", type=int, default=1_000_000)
    ap.add_argument("--config4-steps", type=int, default=3)
    ap.add_argument("--config4-check", type=int, default=48, help="streams per rank decoded by the reference")
    ap.add_argument("--oracle-check", type=int, default=64, help="config-2 streams per rank decoded by the reference")
    ap.add_argument("--no-acmtool", action="store_true", help="reference arm: skip the acmtool / xargs variant")
    ap.add_argument("--acmtool-streams", type=int, default=N_STREAMS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-streaming", action="store_true", help="skip the single-stream libacm.h timing block")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4"],
                    help="config2 = BASELINE configs[1] (the headline); config3 = configs[2] (filler-stress corpus, levels 4-10: the "
                         "general path); config4 = configs[3] shape per GPU")
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
