"""Sharding a batch of independent streams over the GPUs of one box.

Streams share no state, so the data path needs no collective: every rank decodes its own
streams.  The only exchange is an optional all_gather of the per-stream results
(status / words / checksum), 16 bytes per stream, over NCCL on GPUs (gloo in the CPU tests).
"""
from __future__ import annotations

import heapq

import numpy as np


def partition(work, world: int):
    """Longest-processing-time-first split of stream indices into `world` shards.

    `work` is a per-stream cost (total_values: decode cost is proportional to the words
    produced).  Returns a list of `world` int64 index arrays, each sorted ascending, whose
    cost sums differ by at most one stream's cost.  Deterministic.
    """
    work = np.asarray(work, dtype=np.int64)
    order = np.argsort(-work, kind="stable")
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(int(i))
        heapq.heappush(heap, (load + int(work[i]), r))
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def gather_results(local_idx, status, words, checksum, n_total: int, device=None):
    """all_gather the per-stream results of every rank into full-length arrays (indexed by the
    caller's global stream index).  Works on any initialised torch.distributed backend."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    n_local = len(local_idx)
    counts = torch.tensor([n_local], dtype=torch.int64, device=device)
    all_counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    cap = int(max(int(c.item()) for c in all_counts))
    pack = torch.zeros((cap, 4), dtype=torch.int64, device=device)
    if n_local:
        pack[:n_local, 0] = torch.as_tensor(np.asarray(local_idx, dtype=np.int64), device=device)
        pack[:n_local, 1] = torch.as_tensor(np.asarray(status, dtype=np.int64), device=device)
        pack[:n_local, 2] = torch.as_tensor(np.asarray(words, dtype=np.int64), device=device)
        pack[:n_local, 3] = torch.as_tensor(np.asarray(checksum, dtype=np.uint64).view(np.int64), device=device)
    out = [torch.zeros_like(pack) for _ in range(world)]
    dist.all_gather(out, pack)
    g_status = np.zeros(n_total, dtype=np.int32)
    g_words = np.zeros(n_total, dtype=np.uint32)
    g_cks = np.zeros(n_total, dtype=np.uint64)
    for r in range(world):
        k = int(all_counts[r].item())
        t = out[r][:k].cpu().numpy()
        idx = t[:, 0]
        g_status[idx] = t[:, 1]
        g_words[idx] = t[:, 2]
        g_cks[idx] = t[:, 3].view(np.uint64)
    return g_status, g_words, g_cks
