"""Seeded synthetic corpora shared by the CPU and GPU tests."""
from __future__ import annotations

import numpy as np

from libacm_b200 import gen

STRESS_LEVELS = range(0, 11)
STRESS_ROWS = (1, 2, 3, 5, 16, 33, 100)


def stress_params(max_values=120_000, seed=4000):
    """Config 3: every level x rows x {plain, WAVC} x {mono, stereo}, random fillers,
    total_values NOT a multiple of block_len, >= 4 blocks where affordable."""
    out, k = [], 0
    for level in STRESS_LEVELS:
        for rows in STRESS_ROWS:
            for wavc in (0, 1):
                for ch in (1, 2):
                    blen = rows << level
                    tv = min(blen * 4 + max(1, blen // 3), max(max_values, blen + 1))
                    out.append(gen.params(level=level, rows=rows, channels=ch, total_values=tv,
                                          wavc=wavc, dist=gen.DIST_STRESS, seed=seed + k))
                    k += 1
    return out


def single_filler_params(level=5, rows=16, seed=9000):
    """One stream per valid selector (unit tier: each filler in isolation)."""
    return [gen.params(level=level, rows=rows, total_values=(rows << level) * 3 + 5,
                       dist=gen.DIST_SINGLE, single_ind=ind, seed=seed + ind)
            for ind in gen.VALID_INDS]


def fallout_params(n, seed=1, lo=22050, hi=220500, level=7, rows=16, channels=1):
    """Config 2 shape: mono 22050 Hz clips, uniform 1-10 s."""
    rng = np.random.default_rng(seed)
    tv = rng.integers(lo, hi + 1, size=n)
    return [gen.params(level=level, rows=rows, channels=channels, total_values=int(t),
                       dist=gen.DIST_FALLOUT, seed=seed * 1_000_003 + i) for i, t in enumerate(tv)]


def negative_params(seed=12000):
    """Streams with one deliberate defect each: bad selectors and out-of-range t-codes."""
    out = []
    for k, bad in enumerate(gen.BAD_INDS):
        out.append(gen.params(level=5, rows=7, total_values=224 * 5, dist=gen.DIST_STRESS,
                              seed=seed + k, inject=gen.INJECT_BAD_IND, inject_block=k % 4,
                              inject_col=(5 * k + 3) % 32, inject_value=bad))
    for k, ind in enumerate((19, 22, 29)):
        out.append(gen.params(level=4, rows=9, total_values=144 * 4, dist=gen.DIST_SINGLE,
                              single_ind=ind, seed=seed + 100 + k, inject=gen.INJECT_BAD_TCODE,
                              inject_block=1 + k % 2, inject_col=(3 * k + 2) % 16))
    return out


def images(plist):
    return [gen.make_stream(**{f: getattr(p, f) for f, _ in p._fields_ if f != "reserved"}) for p in plist]
