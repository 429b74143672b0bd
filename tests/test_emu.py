"""CPU tier: the product's __host__ __device__ decode logic (code tables, block scan,
column decode, transform, output, descriptor construction), driven by a CPU mirror of
the generic kernel's control flow, against the checker."""
import numpy as np

from libacm_b200 import api, gen
from tests import corpus, emu_bindings as emu


def _check(img, checker, **kw):
    a = checker.decode(img, force_chans=kw.get("force_chans", 0), be=kw.get("be", 0), sgned=kw.get("sgned", 1))
    err, st, words, pcm, cks = emu.decode(img, **kw)
    assert err == a.open_err
    if err == 0:
        assert (st, words) == (a.status, a.words)
        assert np.array_equal(pcm, a.pcm)
        assert cks == api.checksum_ref(a.pcm, a.words, 2, kw.get("be", 0))
    return a


def test_emu_stress_corpus(checker):
    for k, img in enumerate(corpus.images(corpus.stress_params(max_values=40_000))):
        _check(img, checker, be=k & 1, sgned=(k >> 1) & 1, lead=k % 7, nthreads=(1, 32, 64, 256)[k % 4])


def test_emu_single_fillers_and_negatives(checker):
    seen = set()
    for img in corpus.images(corpus.single_filler_params() + corpus.negative_params()):
        seen.add(_check(img, checker).status)
    assert seen == {0, -6}


def test_emu_truncations(checker):
    img = gen.make_stream(level=5, rows=7, channels=2, total_values=5000, dist=gen.DIST_STRESS, seed=5)
    seen = set()
    for cut in range(len(img)):
        a = _check(img[:cut], checker, lead=cut % 5)
        if a.open_err == 0:
            seen.add(a.status)
    assert seen == {0, -6, -7}


def test_emu_force_chans(checker):
    for fc in (-1, 0, 1, 2, 3):
        for wavc in (0, 1):
            for ch in (1, 2):
                for level, rows in ((0, 5), (3, 3), (7, 16)):
                    img = gen.make_stream(level=level, rows=rows, channels=ch, wavc=wavc,
                                          total_values=(rows << level) * 3 + 1, dist=gen.DIST_STRESS,
                                          seed=fc + 10 * wavc + 100 * ch + level)
                    _check(img, checker, force_chans=fc)


def test_fast2_core_walk_and_unpack():
    """The fast kernel's table walk (uni16) and column unpackers, run on the CPU, agree with the
    generic scan / column decode on every column: Fallout mix, stress mix (all fillers, extreme
    val), every single filler, a bad selector, a bad t-code and every truncation of a stream."""
    blocks = 0
    plist = corpus.fallout_params(6, seed=3, lo=20_000, hi=60_000)
    plist += [gen.params(level=7, rows=16, channels=1 + (k & 1), total_values=2048 * 5 + 77 * k, wavc=k & 1,
                         dist=gen.DIST_STRESS, seed=700 + k) for k in range(6)]
    plist += [gen.params(level=7, rows=16, total_values=2048 * 2 + 5, dist=gen.DIST_SINGLE, single_ind=ind,
                         seed=800 + ind) for ind in gen.VALID_INDS]
    plist += [gen.params(level=7, rows=16, total_values=2048 * 3, dist=gen.DIST_STRESS, seed=900,
                         inject=gen.INJECT_BAD_IND, inject_block=1, inject_col=77, inject_value=25),
              gen.params(level=7, rows=16, total_values=2048 * 3, dist=gen.DIST_SINGLE, single_ind=22, seed=901,
                         inject=gen.INJECT_BAD_TCODE, inject_block=1, inject_col=5)]
    for img in corpus.images(plist):
        n = emu.fast2_check(img)
        assert n >= 0, n
        blocks += n
    img = corpus.images([gen.params(level=7, rows=16, total_values=2048 * 2, dist=gen.DIST_STRESS, seed=77)])[0]
    for cut in range(14, len(img), 7):
        assert emu.fast2_check(img[:cut]) >= 0, cut
    assert blocks > 100


def test_fast2_core_on_garbage_payloads():
    """Arbitrary bytes behind a valid level-7 / 16-row header: the table walk must reach the same
    verdict as the generic scan (a clean block, or "re-walk me": bad selector / ran past the end),
    and where a block is clean, the same offsets and values.  Random data runs into bad selectors
    (6 of 32 codes), 16-bit linear columns (the SKIP6 page) and the end of the stream all the time."""
    rng = np.random.default_rng(2026)
    hdr = bytes(corpus.images([gen.params(level=7, rows=16, total_values=2048 * 40, dist=gen.DIST_STRESS, seed=1)])[0][:14])
    clean = 0
    for k in range(400):
        body = rng.integers(0, 256, size=int(rng.integers(60, 9000)), dtype=np.uint8)
        if k % 3 == 0:
            body[rng.random(body.size) < 0.985] = 0   # long runs of zero columns / zero symbols
        if k % 3 == 1:
            body &= 0x9C                              # selectors biased towards valid small codes
        n = emu.fast2_check(hdr + body.tobytes())
        assert n >= 0, (k, n)
        clean += n
    assert clean > 20


def test_fast2_walk_statistics_of_the_bench_corpus():
    """The numbers DESIGN.md / profiles quote for the walk on Fallout-style streams: about 397 table
    steps and 7.1 kbit per block, 128 columns of which about 32 linear, 30 radix-coded, 60 prefix-coded;
    one step in six advances more than 31 bits (whole linear / radix columns)."""
    st = emu.fast2_stats(corpus.images(corpus.fallout_params(12, seed=3, lo=20_000, hi=120_000)))
    b = st["blocks"]
    assert b > 300 and st["columns"] == 128 * b
    assert st["zero"] + st["linear"] + st["k"] + st["t"] == st["columns"]
    assert 380 < st["steps"] / b < 415
    assert 6800 < st["bits"] / b < 7500
    assert st["linear"] == 32 * b and 25 < st["t"] / b < 35
    assert st["big_steps"] == st["linear"] + st["t"]
    assert st["capped_steps"] > 1.2 * st["steps"]
