"""CPU tier: the oracle restatement is pinned against the reference itself
(oracle/_ref, built from /root/reference) and against the committed golden fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest

from libacm_b200 import gen
from oracle import bindings
from tests import corpus

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not bindings.have_ref(), reason="oracle/_ref not built (no /root/reference)")


def _same(a, b):
    assert a.open_err == b.open_err
    assert a.status == b.status
    assert a.words == b.words
    assert np.array_equal(a.pcm, b.pcm)
    assert a.info.as_dict() == b.info.as_dict()


@needs_ref
def test_port_equals_reference_on_stress_corpus(oracle_port):
    ref = bindings.Ref()
    for k, img in enumerate(corpus.images(corpus.stress_params(max_values=60_000))):
        for be, sg in ((0, 1), (1, 1), (0, 0), (1, 0)):
            _same(oracle_port.decode(img, be=be, sgned=sg), ref.decode(img, be=be, sgned=sg))


@needs_ref
def test_port_equals_reference_single_fillers_and_negatives(oracle_port):
    ref = bindings.Ref()
    seen = set()
    for img in corpus.images(corpus.single_filler_params() + corpus.negative_params()):
        a, b = oracle_port.decode(img), ref.decode(img)
        _same(a, b)
        seen.add(b.status)
    assert -6 in seen and 0 in seen


@needs_ref
def test_port_equals_reference_on_every_truncation(oracle_port):
    ref = bindings.Ref()
    img = gen.make_stream(level=5, rows=7, channels=2, total_values=5000, dist=gen.DIST_STRESS, seed=5)
    seen = set()
    for cut in range(len(img)):
        a, b = oracle_port.decode(img[:cut]), ref.decode(img[:cut])
        assert a.open_err == b.open_err
        if b.open_err == 0:
            _same(a, b)
            seen.add(b.status)
    assert seen == {0, -6, -7}


@needs_ref
@pytest.mark.parametrize("fc", [-1, 0, 1, 2, 3])
def test_port_equals_reference_force_chans(oracle_port, fc):
    ref = bindings.Ref()
    for wavc in (0, 1):
        for ch in (1, 2):
            for level, rows in ((0, 5), (3, 3), (7, 16)):
                img = gen.make_stream(level=level, rows=rows, channels=ch, wavc=wavc,
                                      total_values=(rows << level) * 3 + 1, dist=gen.DIST_STRESS,
                                      seed=fc + 10 * wavc + 100 * ch + level)
                _same(oracle_port.decode(img, force_chans=fc), ref.decode(img, force_chans=fc))


def test_golden_fixtures(oracle_port):
    """tests/golden/*.acm were decoded ONCE by the compiled reference
    (tests/golden/make_golden.py); their PCM hashes are committed."""
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        meta = json.load(f)
    assert len(meta) >= 10
    checkers = [oracle_port] + ([bindings.Ref()] if bindings.have_ref() else [])
    for name, want in meta.items():
        img = open(os.path.join(GOLDEN, name), "rb").read()
        for chk in checkers:
            for fmt, digest in want["sha256"].items():
                be, sg = int(fmt[1]), int(fmt[3])
                r = chk.decode(img, force_chans=want["force_chans"], be=be, sgned=sg)
                assert r.open_err == want["open_err"]
                if r.open_err == 0:
                    assert (r.status, r.words) == (want["status"], want["words"])
                    assert hashlib.sha256(r.pcm.tobytes()).hexdigest() == digest, (name, fmt, chk.kind)


def test_wordlen_3_4_low16_invariant(oracle_port):
    """Parity UNPINNED for wordlen 3/4 (the reference returns BADFMT): pinned only by
    'low 16 bits equal the s16 output'."""
    img = gen.make_stream(level=6, rows=5, channels=2, total_values=2000, dist=gen.DIST_STRESS, seed=3)
    s16 = oracle_port.decode(img).pcm.view("<u2")
    for wl in (3, 4):
        for be in (0, 1):
            for sg in (0, 1):
                r = oracle_port.decode(img, be=be, wordlen=wl, sgned=sg)
                b = r.pcm.reshape(-1, wl)
                lo = (b[:, wl - 1].astype(np.uint16) | (b[:, wl - 2].astype(np.uint16) << 8)) if be else \
                     (b[:, 0].astype(np.uint16) | (b[:, 1].astype(np.uint16) << 8))
                assert np.array_equal(lo, s16)
