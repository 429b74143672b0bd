"""CPU tier: the N>1 host logic (stream partition + result gather) with world_size 2 on gloo.
The per-rank "decode" here is the checker, standing in for the GPU: what is under test is
the sharding and the collective, not the decoder."""
import os
import socket

import numpy as np
import pytest

from libacm_b200 import api, shard
from tests import corpus


def test_partition_balanced_and_complete():
    rng = np.random.default_rng(0)
    work = rng.integers(22050, 220500, size=10_000)
    for world in (1, 2, 4, 8):
        parts = shard.partition(work, world)
        allidx = np.concatenate(parts)
        assert np.array_equal(np.sort(allidx), np.arange(work.size))
        loads = np.array([work[p].sum() for p in parts])
        assert loads.max() - loads.min() <= work.max()
    assert [p.tolist() for p in shard.partition([5, 1, 1, 1, 1, 1], 2)] == [[0], [1, 2, 3, 4, 5]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import bindings
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plist = corpus.stress_params(max_values=3000)[::7] + corpus.negative_params()
    imgs = corpus.images(plist)
    chk = bindings.Oracle()
    totals = [chk.parse(i)[1].total_values for i in imgs]
    mine = shard.partition(totals, world)[rank]
    st, wd, ck = [], [], []
    for i in mine:
        r = chk.decode(imgs[i])
        st.append(r.status), wd.append(r.words), ck.append(api.checksum_ref(r.pcm, r.words))
    g = shard.gather_results(mine, st, wd, ck, len(imgs))
    if rank == 0:
        q.put(tuple(x.tolist() for x in g))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    import torch.multiprocessing as mp
    from oracle import bindings
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    plist = corpus.stress_params(max_values=3000)[::7] + corpus.negative_params()
    chk = bindings.Oracle()
    want = [chk.decode(i) for i in corpus.images(plist)]
    assert got[0] == [r.status for r in want]
    assert got[1] == [r.words for r in want]
    assert got[2] == [api.checksum_ref(r.pcm, r.words) for r in want]
    assert -6 in got[0]
