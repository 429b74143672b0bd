"""GPU tier: several callers and several GPUs through the C ABI.

* SURVEY.md section 8b "threading": the reference has no globals and is re-entrant across streams;
  the drop-in must be too -- two host threads, each with its own ACMStream / its own batch, on one GPU.
* a batch shards over the GPUs of a box by stream with no collective;
  acm_gpu_opts.device_mask does it inside ONE acm_gpu_decode_batch call (needs >= 2 GPUs).
"""
import threading

import numpy as np
import pytest

from libacm_b200 import api, gen
from tests import api_driver as ad
from tests import corpus, gpu_util as gu

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_two_threads_two_batches_one_gpu(checker):
    """Two host threads call acm_gpu_decode_batch at the same time on the same device (they take
    turns on its workspace): each gets exactly what it gets alone."""
    sets = [corpus.images(corpus.fallout_params(60, seed=70 + k, hi=60_000) +
                          corpus.stress_params(max_values=10_000)[k::7]) for k in range(2)]
    alone = [gu.decode_host(imgs, want_checksums=1) for imgs in sets]
    got = [None, None]

    def work(k):
        for _ in range(3):
            got[k] = gu.decode_host(sets[k], want_checksums=1)

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for k in range(2):
        (s0, o0), (s1, o1) = alone[k], got[k]
        for f in ("status", "words", "checksum"):
            assert np.array_equal(s0[f], s1[f]), (k, f)
        assert np.array_equal(o0, o1), k
        assert gu.compare(sets[k][::9], s1[::9], o1, checker, checksums=True) == []


@pytest.mark.skipif(not ad.have_ref(), reason="oracle/_ref not built")
def test_two_threads_two_streams_libacm_surface():
    """The libacm.h surface from two host threads, one ACMStream each (the reference's contract:
    re-entrant across streams, not thread-safe per stream): every acm_read of either thread returns
    what the reference returns for that stream."""
    libs = (ad.mine(), ad.ref())
    imgs = [gen.make_stream(level=7, rows=16, channels=2, rate=44100, total_values=2048 * 120 + 77 * k,
                            seed=300 + k) for k in range(2)]
    want = []
    for img in imgs:
        h = ad.Handle(libs[1], img)
        chunks = []
        while True:
            r, dta = h.read(6000, loop=True)
            chunks.append((r, dta))
            if r <= 0:
                break
        h.close()
        want.append(chunks)
    errs = []

    def work(k):
        try:
            h = ad.Handle(libs[0], imgs[k])
            for n, (r0, d0) in enumerate(want[k]):
                r, dta = h.read(6000, loop=True)
                if (r, dta) != (r0, d0):
                    errs.append((k, n, r, r0))
                    break
            h.close()
        except Exception as ex:  # noqa: BLE001
            errs.append((k, repr(ex)))

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert errs == []


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_device_mask_shards_one_call_over_gpus(checker):
    """One acm_gpu_decode_batch call with device_mask = all GPUs of the box: same bytes, statuses,
    words and checksums as the single-GPU call; defective streams keep their verdicts wherever the
    cut puts them."""
    n = min(_n_gpus(), 8)
    imgs = corpus.images(corpus.fallout_params(600, seed=91, hi=90_000) + corpus.negative_params() +
                         corpus.stress_params(max_values=12_000)[::3])
    s1, o1 = gu.decode_host(imgs, want_checksums=1, device=0)
    sN, oN = gu.decode_host(imgs, want_checksums=1, device_mask=(1 << n) - 1)
    for f in ("status", "words", "checksum", "total_values"):
        assert np.array_equal(s1[f], sN[f]), f
    assert np.array_equal(o1, oN)
    assert gu.compare(imgs[::23], sN[::23], oN, checker, checksums=True) == []
    # device buffers belong to one GPU: refused, not guessed
    import torch
    blob, offs, lens = gu.pack(imgs[:4])
    s = api.new_streams(offs, lens)
    opts = api.make_opts(device_mask=3)
    api.probe(blob, s, opts)
    d_out = torch.empty(api.layout(s, 2), dtype=torch.uint8, device="cuda:0")
    with pytest.raises(Exception):
        api.decode_batch(torch.from_numpy(blob).cuda(), s, d_out, opts)
