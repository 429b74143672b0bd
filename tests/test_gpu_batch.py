"""GPU tier: acm_gpu_decode_batch / acm_gpu_plan_* through the C ABI against the checker
(the reference itself when oracle/_ref is present).  Bit-exact, zero tolerance."""
import numpy as np
import pytest

from libacm_b200 import api, gen
from tests import corpus, gpu_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_stress_corpus_host_buffers(checker, kernel):
    imgs = corpus.images(corpus.stress_params(max_values=60_000))
    s, out = gu.decode_host(imgs, want_checksums=1, kernel=kernel)
    assert gu.compare(imgs, s, out, checker, checksums=True) == []


@pytest.mark.parametrize("be,sgned", [(0, 1), (1, 1), (0, 0), (1, 0)])
def test_formats_device_resident(checker, be, sgned):
    imgs = corpus.images(corpus.stress_params(max_values=20_000)[::3] + corpus.fallout_params(24, seed=3, hi=40_000))
    s, out = gu.decode_device(imgs, bigendianp=be, sgned=sgned, want_checksums=1)
    assert gu.compare(imgs, s, out, checker, be=be, sgned=sgned, checksums=True) == []


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_unaligned_back_to_back_images(checker, kernel):
    imgs = corpus.images(corpus.stress_params(max_values=8_000)[::2] + corpus.fallout_params(40, seed=9, hi=30_000))
    for lead in (0, 1, 2, 3):
        s, out = gu.decode_host(imgs, align=1, lead=lead, kernel=kernel)
        assert gu.compare(imgs, s, out, checker) == []


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_single_fillers_and_negatives(checker, kernel):
    imgs = corpus.images(corpus.single_filler_params() + corpus.single_filler_params(level=7, rows=16) +
                         corpus.negative_params())
    s, out = gu.decode_host(imgs, kernel=kernel)
    assert gu.compare(imgs, s, out, checker) == []
    assert set(s["status"].tolist()) == {0, -6}


@pytest.mark.parametrize("kernel", [0, 1, 2])
@pytest.mark.parametrize("shape", [(5, 7, 2), (7, 16, 1)])
def test_truncations(checker, kernel, shape):
    level, rows, ch = shape
    img = gen.make_stream(level=level, rows=rows, channels=ch, total_values=(rows << level) * 3 + 9,
                          dist=gen.DIST_STRESS, seed=5)
    cuts = list(range(0, 60)) + list(range(60, len(img), 7)) + [len(img) - 1, len(img)]
    imgs = [img[:c] for c in cuts]
    s, out = gu.decode_host(imgs, align=1, kernel=kernel)
    assert gu.compare(imgs, s, out, checker) == []
    assert {0, -3, -7} <= set(s["status"].tolist())


@pytest.mark.parametrize("fc", [-1, 0, 1, 2, 3])
def test_force_chans(checker, fc):
    imgs = []
    for wavc in (0, 1):
        for ch in (1, 2):
            for level, rows in ((0, 5), (3, 3), (7, 16)):
                imgs.append(gen.make_stream(level=level, rows=rows, channels=ch, wavc=wavc,
                                            total_values=(rows << level) * 3 + 1, dist=gen.DIST_STRESS,
                                            seed=fc + 10 * wavc + 100 * ch + level))
    s, out = gu.decode_host(imgs, force_chans=fc)
    assert gu.compare(imgs, s, out, checker, force_chans=fc) == []


@pytest.mark.parametrize("wordlen", [3, 4])
def test_wordlen_3_4_unpinned_invariant(checker, wordlen):
    """No reference answer exists for wordlen 3/4 (decode.c:832-835): low 16 bits must
    equal the s16 output, and the oracle port's definition must match."""
    from oracle import bindings
    port = bindings.Oracle()
    imgs = corpus.images(corpus.stress_params(max_values=8_000)[::5] + corpus.fallout_params(8, seed=2, hi=30_000))
    for be in (0, 1):
        for sg in (0, 1):
            s, out = gu.decode_host(imgs, wordlen=wordlen, bigendianp=be, sgned=sg)
            for i, img in enumerate(imgs):
                want = port.decode(img, be=be, wordlen=wordlen, sgned=sg)
                o = int(s["out_off"][i])
                assert np.array_equal(out[o:o + want.pcm.size], want.pcm)
                s16 = checker.decode(img).pcm.view("<u2")
                b = want.pcm.reshape(-1, wordlen)
                lo = (b[:, -1].astype(np.uint16) | (b[:, -2].astype(np.uint16) << 8)) if be else \
                     (b[:, 0].astype(np.uint16) | (b[:, 1].astype(np.uint16) << 8))
                assert np.array_equal(lo, s16)


def test_golden_fixtures_on_gpu():
    import hashlib
    import json
    import os
    g = os.path.join(os.path.dirname(__file__), "golden")
    meta = json.load(open(os.path.join(g, "golden.json")))
    names = sorted(meta)
    for fc in sorted({m["force_chans"] for m in meta.values()}):
        sel = [n for n in names if meta[n]["force_chans"] == fc]
        imgs = [open(os.path.join(g, n), "rb").read() for n in sel]
        for be in (0, 1):
            for sg in (0, 1):
                s, out = gu.decode_host(imgs, bigendianp=be, sgned=sg, force_chans=fc)
                for i, n in enumerate(sel):
                    m = meta[n]
                    if m["open_err"] < 0:
                        assert s["status"][i] == m["open_err"]
                        continue
                    assert (int(s["status"][i]), int(s["words"][i])) == (m["status"], m["words"]), n
                    o = int(s["out_off"][i])
                    pcm = out[o:o + m["info"]["total_values"] * 2]
                    assert hashlib.sha256(pcm.tobytes()).hexdigest() == m["sha256"][f"b{be}s{sg}"], n


def test_fallout_batch_checksums_and_roundtrip_properties(checker):
    """Larger batch (config 2 shape, scaled): per-stream checksums equal the checker's and
    the checksum of checksums is invariant under a permutation of the batch order."""
    imgs = corpus.images(corpus.fallout_params(300, seed=7))
    s, out = gu.decode_device(imgs, want_checksums=1)
    assert gu.compare(imgs[::15], s[::15], out, checker, checksums=True) == []
    perm = np.random.default_rng(0).permutation(len(imgs))
    s2, _ = gu.decode_device([imgs[i] for i in perm], want_checksums=1)
    assert np.array_equal(s2["checksum"], s["checksum"][perm])
    assert np.all(s["status"] == 0) and np.array_equal(s["words"], s["total_values"])


def test_host_path_segmented_pipeline_matches_resident(checker):
    """A batch large enough (>= 32 MB of PCM) for acm_gpu_decode_batch to cut it into segments
    whose copy-in / decode / copy-out overlap: same bytes as the single-launch resident path."""
    imgs = corpus.images(corpus.fallout_params(420, seed=21) + corpus.stress_params(max_values=20_000)[::11])
    s1, out1 = gu.decode_host(imgs, want_checksums=1)
    s2, out2 = gu.decode_device(imgs, want_checksums=1)
    assert out1.size >= (32 << 20)
    assert np.array_equal(s1["status"], s2["status"]) and np.array_equal(s1["words"], s2["words"])
    assert np.array_equal(s1["checksum"], s2["checksum"])
    for i in range(len(imgs)):
        o, nb = int(s1["out_off"][i]), int(s1["total_values"][i]) * 2
        assert np.array_equal(out1[o:o + nb], out2[o:o + nb]), i
        gap = out1[o + nb:o + ((nb + 15) & ~15)]
        assert not gap.any()            # alignment gaps are zeroed by the kernels, never stale
    assert gu.compare(imgs[::37], s1[::37], out1, checker, checksums=True) == []
    # second call reuses the cached workspace
    s3, out3 = gu.decode_host(imgs[:200], want_checksums=1)
    assert np.array_equal(s3["checksum"], s1["checksum"][:200])


def test_fast_shape_negatives_among_good_streams(checker):
    """Level-7 / 16-row streams (the scan-CTA / decode-CTA kernel) with one defect each -- bad
    selectors (the scan's table walk ends on its BAD page and the block is re-walked with the
    reference's verdicts) and out-of-range t-codes (found by the decode side, which must stop the
    slot's scan lane) -- in the middle of healthy streams sharing the same slots."""
    plist = corpus.fallout_params(40, seed=31, hi=50_000)
    for k, bad in enumerate(gen.BAD_INDS):
        plist.insert(3 * k + 1, gen.params(level=7, rows=16, total_values=2048 * 6 + 11 * k, dist=gen.DIST_STRESS,
                                           seed=5000 + k, inject=gen.INJECT_BAD_IND, inject_block=k % 5,
                                           inject_col=(37 * k + 3) % 128, inject_value=bad))
    for k, ind in enumerate((19, 22, 29, 19, 22, 29)):
        plist.insert(5 * k + 2, gen.params(level=7, rows=16, total_values=2048 * 7, dist=gen.DIST_SINGLE,
                                           single_ind=ind, seed=5100 + k, inject=gen.INJECT_BAD_TCODE,
                                           inject_block=k, inject_col=(29 * k + 1) % 128))
    imgs = corpus.images(plist)
    for dec in (gu.decode_host, gu.decode_device):
        s, out = dec(imgs, want_checksums=1)
        assert gu.compare(imgs, s, out, checker, checksums=True) == []
        assert set(s["status"].tolist()) == {0, -6}


def test_more_streams_than_slots(checker):
    """Thousands of one-to-three-block streams: every stream slot is reused many times, every decode
    CTA owns many slots, records of consecutive streams share a slot's ring."""
    rng = np.random.default_rng(11)
    tv = rng.integers(1, 3 * 2048 + 1, size=3000)
    plist = [gen.params(level=7, rows=16, channels=1 + (i & 1), total_values=int(t), wavc=(i >> 1) & 1,
                        dist=gen.DIST_FALLOUT if i % 3 else gen.DIST_STRESS, seed=9000 + i)
             for i, t in enumerate(tv)]
    imgs = corpus.images(plist)
    s, out = gu.decode_device(imgs, want_checksums=1)
    assert np.all(s["status"] == 0)
    assert gu.compare(imgs[::7], s[::7], out, checker, checksums=True) == []
    want = s["total_values"] - s["total_values"] % s["channels"]
    assert np.array_equal(s["words"], want)


def test_config2_full_size(checker):
    """BASELINE configs[1] at full size (the bench workload: 10 000 clips, 1.2 G samples): every
    stream ends with status 0 and all its words, every 12th stream is compared byte for byte (and
    by checksum) with the checker, and the checksum of checksums equals the value the bench's parity
    gate has reported since the first generic-kernel run of this corpus."""
    import torch
    import bench
    blob, offs, lens = bench.build_corpus(bench.N_STREAMS, 0)
    opts = api.make_opts(want_checksums=1)
    s = api.new_streams(offs, lens)
    d_blob = torch.from_numpy(blob).cuda()
    api.probe(blob, s, opts)
    nbytes = api.layout(s, 2)
    d_out = torch.full((nbytes + 16,), 0xAA, dtype=torch.uint8, device="cuda")
    plan = api.Plan(s, opts)
    cs = torch.cuda.current_stream().cuda_stream
    plan.run(d_blob, d_out, cs)
    plan.fetch(s, cs)
    plan.close()
    assert np.all(s["status"] == 0) and np.array_equal(s["words"], s["total_values"])
    assert int(np.sum(s["checksum"].astype(np.uint64), dtype=np.uint64)) == 2964506047360995033
    out = d_out.cpu().numpy()
    for i in range(0, bench.N_STREAMS, 12):
        o, l = int(offs[i]), int(lens[i])
        a = checker.decode(blob[o:o + l])
        p = int(s["out_off"][i])
        assert (a.status, a.words) == (0, int(s["words"][i])), i
        assert np.array_equal(out[p:p + a.pcm.size], a.pcm), i
        assert int(s["checksum"][i]) == api.checksum_ref(a.pcm, a.words), i


def test_replication_by_descriptor():
    """BASELINE configs[3] mechanics at small scale (tools/scale_1m.py runs it with 1 000 000 streams):
    several descriptors reference the same image and write their own output range; all replicas of an
    image produce the same bytes and checksum."""
    import torch
    imgs = corpus.images(corpus.fallout_params(300, seed=41, lo=500, hi=30_000, channels=2))
    blob, offs, lens = gu.pack(imgs)
    reps = 6
    opts = api.make_opts(want_checksums=1)
    s = api.new_streams(np.tile(offs, reps), np.tile(lens, reps))
    d_blob = torch.from_numpy(blob).cuda()
    api.probe(blob, s, opts)
    nbytes = api.layout(s, 2)
    d_out = torch.full((nbytes + 16,), 0xAA, dtype=torch.uint8, device="cuda")
    plan = api.Plan(s, opts)
    plan.run(d_blob, d_out, torch.cuda.current_stream().cuda_stream)
    plan.fetch(s, torch.cuda.current_stream().cuda_stream)
    plan.close()
    out = d_out.cpu().numpy()
    assert np.all(s["status"] == 0)
    ck = s["checksum"].reshape(reps, len(imgs))
    assert np.all(ck == ck[0])
    for i in range(0, len(imgs), 17):
        nb = int(s["total_values"][i]) * 2
        first = out[int(s["out_off"][i]):int(s["out_off"][i]) + nb]
        for r in range(1, reps):
            o = int(s["out_off"][r * len(imgs) + i])
            assert np.array_equal(out[o:o + nb], first)


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_garbage_payloads_status_and_words(checker, kernel):
    """Arbitrary bytes behind a valid level-7 / 16-row header.  The PCM of such streams is not defined by
    the reference (a dequantisation index outside the block's table reads stale memory), but
    what its read loop returns -- status and words delivered -- is: both kernels must agree with it."""
    rng = np.random.default_rng(77)
    hdr = bytes(corpus.images([gen.params(level=7, rows=16, total_values=2048 * 6, dist=gen.DIST_STRESS, seed=1)])[0][:14])
    imgs = []
    for k in range(240):
        body = rng.integers(0, 256, size=int(rng.integers(40, 7000)), dtype=np.uint8)
        if k % 3 == 0:
            body[rng.random(body.size) < 0.985] = 0
        if k % 3 == 1:
            body &= 0x9C
        imgs.append(hdr + body.tobytes())
    s, _ = gu.decode_host(imgs, align=1, kernel=kernel)
    seen = set()
    for i, img in enumerate(imgs):
        a = checker.decode(img)
        assert (int(s["status"][i]), int(s["words"][i])) == (a.status, a.words), (i, k)
        seen.add(a.status)
    assert {-6, 0} <= seen


def test_healthy_streams_never_take_the_re_walk_path(checker):
    """Every byte alignment of the stream start (position 0 of a 16-byte chunk included): the table
    walk alone must carry a healthy stream.  The re-walk by the generic block scan is for bad selectors
    and truncated streams; a walk that silently leans on it is slow, not wrong, so parity cannot see it."""
    import torch
    imgs = corpus.images(corpus.fallout_params(48, seed=21, lo=8_000, hi=40_000))
    for lead in range(16):
        blob, offs, lens = gu.pack(imgs, 1, lead)
        opts = api.make_opts(device=0, want_checksums=1)
        d_blob = torch.from_numpy(blob).cuda()
        s = api.new_streams(offs, lens)
        api.probe(d_blob, s, opts)
        d_out = torch.empty(api.layout(s, 2) + 16, dtype=torch.uint8, device="cuda")
        plan = api.Plan(s, opts)
        plan.run(d_blob, d_out, torch.cuda.current_stream().cuda_stream)
        plan.fetch(s, torch.cuda.current_stream().cuda_stream)
        rewalked = int(plan.counters()[32])
        plan.close()
        assert set(s["status"].tolist()) == {0}
        assert rewalked == 0, (lead, rewalked)
    assert gu.compare(imgs, s, d_out.cpu().numpy(), checker, checksums=True) == []
    # and it does count: a truncated stream ends in a re-walk
    cut = [bytes(imgs[0])[: len(imgs[0]) * 2 // 3]]
    blob, offs, lens = gu.pack(cut, 16, 0)
    d_blob = torch.from_numpy(blob).cuda()
    s = api.new_streams(offs, lens)
    api.probe(d_blob, s, opts)
    d_out = torch.empty(api.layout(s, 2) + 16, dtype=torch.uint8, device="cuda")
    plan = api.Plan(s, opts)
    plan.run(d_blob, d_out, torch.cuda.current_stream().cuda_stream)
    plan.fetch(s, torch.cuda.current_stream().cuda_stream)
    assert int(plan.counters()[32]) == 1
    plan.close()


@pytest.mark.parametrize("kernel", [0, 1, 2])
def test_decode_twice_same_stream_array(checker, kernel):
    """acm_gpu_stream.status is an output: a stream array that has been through a decode -- with
    streams that ended ACM_ERR_CORRUPT (-6), ACM_ERR_UNEXPECTED_EOF (-7) and a rejected header (-3)
    -- decodes to the very same bytes, statuses and word counts when handed in again."""
    good = corpus.images(corpus.fallout_params(6, seed=5, hi=30_000))
    bad = corpus.images(corpus.negative_params())
    # a cut that lands inside a filler payload ends ACM_ERR_UNEXPECTED_EOF (most cuts do; the
    # reference decides which)
    cut7 = next(bytes(good[0])[:n] for n in range(len(good[0]) * 3 // 5, len(good[0]))
                if checker.decode(bytes(good[0])[:n]).status == -7)
    cut = [cut7, bytes(good[1])[:9]]
    imgs = good + bad + cut
    blob, offs, lens = gu.pack(imgs)
    opts = api.make_opts(want_checksums=1, kernel=kernel)
    s = api.new_streams(offs, lens)
    api.probe(blob, s, opts)
    nbytes = api.layout(s, 2)
    out1 = np.full(nbytes, 0xAA, np.uint8)
    api.decode_batch(blob, s, out1, opts)
    first = s.copy()
    assert {0, -3, -6, -7} <= set(first["status"].tolist())
    assert gu.compare(imgs, s, out1, checker, checksums=True) == []
    out2 = np.full(nbytes, 0x55, np.uint8)
    api.decode_batch(blob, s, out2, opts)             # same array, statuses now hold the first verdicts
    for f in ("status", "words", "checksum"):
        assert np.array_equal(s[f], first[f]), f
    assert gu.compare(imgs, s, out2, checker, checksums=True) == []


def test_out_buffer_of_exactly_the_layout_size():
    """`out` has to hold acm_gpu_layout() bytes -- every slot ends on a 16-byte boundary that the
    kernels zero-fill up to -- and a buffer one byte short of the last slot is refused, not overrun."""
    imgs = corpus.images(corpus.fallout_params(5, seed=8, lo=3_001, hi=9_000))
    blob, offs, lens = gu.pack(imgs)
    opts = api.make_opts()
    s = api.new_streams(offs, lens)
    api.probe(blob, s, opts)
    nbytes = api.layout(s, 2)
    assert int(s["total_values"][-1]) * 2 % 16 != 0
    guard = np.full(nbytes + 64, 0x77, np.uint8)
    api.decode_batch(blob, s, guard[:nbytes], opts)
    assert np.all(s["status"] == 0) and np.all(guard[nbytes:] == 0x77)
    with pytest.raises(Exception):
        api.decode_batch(blob, s, guard[:nbytes - 1], opts)


def test_routes_stress_corpus_takes_the_throughput_paths(checker):
    """BASELINE configs[2]: every stream of the filler-stress corpus (levels 0-10, any rows) is decoded by a
    throughput path -- the fused level-7 / 16-row kernel or the general scan -> unpack -> tile-lift path --
    and none by the block-at-a-time backstop, which is only for levels 11-15."""
    import torch
    imgs = corpus.images(corpus.stress_params(max_values=20_000))
    blob, offs, lens = gu.pack(imgs)
    opts = api.make_opts()
    d_blob = torch.from_numpy(blob).cuda()
    s = api.new_streams(offs, lens)
    api.probe(d_blob, s, opts)
    plan = api.Plan(s, opts)
    fused, split, general, backstop = plan.routes()
    plan.close()
    assert fused + split + general + backstop == len(imgs)
    assert backstop == 0 and split == 0
    # the handful of level-7 / 16-row streams of this corpus go with the majority (a fused-kernel launch
    # for four streams would last as long as the walk of the longest of them)
    assert fused == 0 and general == len(imgs)
    # ... while a batch of that shape alone is the fused kernel's
    imgs = corpus.images(corpus.fallout_params(40, seed=2, hi=30_000))
    blob, offs, lens = gu.pack(imgs)
    s = api.new_streams(offs, lens)
    api.probe(blob, s, opts)
    plan = api.Plan(s, opts)
    assert plan.routes() == (40, 0, 0, 0)
    plan.close()


def test_levels_11_and_12_take_the_backstop(checker):
    """acm_level up to 15 is accepted by the reference (decode.c:747); levels above 10 (cols > 1024) are decoded
    block by block through global scratch.  Mixed with tile-path streams in one batch."""
    from libacm_b200 import gen
    plist = []
    for k, (level, rows) in enumerate([(11, 1), (11, 3), (12, 1), (12, 2), (10, 3), (9, 5), (6, 16)]):
        blen = rows << level
        plist.append(gen.params(level=level, rows=rows, channels=1 + k % 2, total_values=blen * 3 + blen // 3 + 1,
                                dist=gen.DIST_STRESS, seed=7000 + k))
    imgs = corpus.images(plist)
    s, out = gu.decode_host(imgs, want_checksums=1)
    assert gu.compare(imgs, s, out, checker, checksums=True) == []
    s, out = gu.decode_device(imgs, bigendianp=1, sgned=0)
    assert gu.compare(imgs, s, out, checker, be=1, sgned=0) == []


def test_general_path_stream_groups(checker, monkeypatch):
    """A resident plan over many general-path streams walks them in groups on CUDA streams of their own
    (acm_gpu_plan_gen_groups): same PCM, statuses, word counts and checksums as the reference, with healthy,
    truncated and corrupt streams of every shape spread over the groups, run twice on the same plan."""
    import torch
    monkeypatch.setenv("ACM_B200_GEN_GROUP_MIN", "64")
    rng = np.random.default_rng(77)
    plist = []
    for k in range(420):
        level, rows = int(rng.integers(0, 10)), int(rng.choice([1, 2, 3, 4, 7, 8, 16, 32]))
        blen = rows << level
        tv = int(rng.integers(blen + 1, max(blen * 2, 24_000)))
        plist.append(gen.params(level=level, rows=rows, channels=1 + k % 2, total_values=tv, wavc=k % 3 == 0,
                                dist=gen.DIST_STRESS, seed=31_000 + k))
    imgs = corpus.images(plist + corpus.negative_params())
    imgs += [img[:len(img) * (3 + k % 5) // 9] for k, img in enumerate(imgs[:60])]   # truncated copies
    order = rng.permutation(len(imgs))
    imgs = [imgs[i] for i in order]
    blob, offs, lens = gu.pack(imgs, align=1, lead=3)
    for groups in ("8", "3", "1"):
        monkeypatch.setenv("ACM_B200_GEN_GROUPS", groups)
        opts = api.make_opts(want_checksums=1)
        d_blob = torch.from_numpy(blob).cuda()
        s = api.new_streams(offs, lens)
        api.probe(d_blob, s, opts)
        nbytes = api.layout(s, opts.wordlen)
        plan = api.Plan(s, opts)
        fused, split, general, backstop = plan.routes()
        assert fused == 0 and split == 0 and backstop == 0 and general > 400
        assert (plan.gen_groups() == 1) if groups == "1" else (2 <= plan.gen_groups() <= int(groups))
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            d_out = torch.full((nbytes + 16,), 0xAA, dtype=torch.uint8, device="cuda")
            plan.run(d_blob, d_out, st)
            plan.fetch(s, st)
            assert gu.compare(imgs, s, d_out.cpu().numpy(), checker, checksums=True) == []
        plan.close()
        # the one-shot entry point with device-resident buffers builds the same kind of plan
        s2 = api.new_streams(offs, lens)
        api.probe(d_blob, s2, opts)
        assert api.layout(s2, opts.wordlen) == nbytes
        d_out = torch.full((nbytes + 16,), 0x55, dtype=torch.uint8, device="cuda")
        api.decode_batch(d_blob, s2, d_out, opts)
        assert gu.compare(imgs, s2, d_out.cpu().numpy(), checker, checksums=True) == []
    assert {0, -6} <= set(s["status"].tolist())
