"""GPU tier: the on-GPU corpus generator (acm_gpu_generate) writes the
same bytes as the host generator, and what it writes decodes like the reference says."""
import numpy as np
import pytest
import torch

from libacm_b200 import api, gen
from tests import corpus, gpu_util as gu

pytestmark = pytest.mark.gpu


def _plist():
    pl = corpus.stress_params(max_values=12_000)[::3]
    pl = [p for p in pl if p.level <= 10]
    pl += corpus.fallout_params(24, seed=11, hi=60_000)
    pl += corpus.single_filler_params(level=7, rows=16)
    pl += corpus.negative_params()
    pl += [gen.params(level=7, rows=16, channels=2, wavc=1, total_values=2048 * 3 + 17, seed=5)]
    return pl


def test_device_images_equal_host_images(checker):
    pl = _plist()
    blob, offs, lens = gen.make_batch(pl)
    d_offs, d_lens, used = api.generate_on_device(pl, None)
    assert np.array_equal(d_lens, lens)
    assert np.array_equal(d_offs, offs)
    assert used == int(offs[-1]) + int(lens[-1])
    d_blob = torch.full((used + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    o2, l2, u2 = api.generate_on_device(pl, d_blob.data_ptr(), used + 64)
    assert (u2, o2.tolist(), l2.tolist()) == (used, offs.tolist(), lens.tolist())
    got = d_blob.cpu().numpy()
    for o, n in zip(offs.tolist(), lens.tolist()):
        assert np.array_equal(got[o:o + n], blob[o:o + n])
        assert not got[o + n:min((o + n + 15) & ~15, used)].any()      # the gap up to the next image reads as zero
    assert np.all(got[used:] == 0xEE)

    # decoded in place (device-resident blob, no host round trip of the corpus)
    s = api.new_streams(offs, lens)
    opts = api.make_opts(device=0, want_checksums=1)
    api.probe(blob, s, opts)
    nbytes = api.layout(s, 2)
    d_out = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    plan = api.Plan(s, opts)
    plan.run(d_blob, d_out, torch.cuda.current_stream().cuda_stream)
    plan.fetch(s, torch.cuda.current_stream().cuda_stream)
    imgs = [bytes(blob[int(o):int(o) + int(n)]) for o, n in zip(offs, lens)]
    assert gu.compare(imgs, s, d_out.cpu().numpy(), checker, checksums=True) == []


def test_too_small_blob_and_bad_level_are_errors():
    pl = corpus.fallout_params(4, seed=2)
    _, _, used = api.generate_on_device(pl, None)
    d_blob = torch.empty(used, dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError):
        api.generate_on_device(pl, d_blob.data_ptr(), used - 16)
    with pytest.raises(RuntimeError):
        api.generate_on_device([gen.params(level=11, rows=1, total_values=4096)], None)
