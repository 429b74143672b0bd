"""Generates tests/golden/: small ACM images + the REFERENCE's answers for them.

Run in the build container (needs oracle/_ref, i.e. /root/reference):
    python tests/golden/make_golden.py
The images come from the C generator; the expected status / word count / PCM sha256
come from the unmodified reference decoder (oracle/_ref/libacm_ref.so).  Only this
script and its outputs are committed -- no reference source.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from libacm_b200 import gen  # noqa: E402
from oracle import bindings  # noqa: E402

CASES = {
    "l7r16_mono_fallout.acm": dict(level=7, rows=16, channels=1, total_values=2048 * 6 + 777, seed=11),
    "l7r16_stereo_fallout.acm": dict(level=7, rows=16, channels=2, total_values=2048 * 5 + 2, seed=12),
    "l7r16_stress.acm": dict(level=7, rows=16, channels=2, total_values=2048 * 4 + 100, seed=13, dist=gen.DIST_STRESS),
    "l4r33_wavc_stress.acm": dict(level=4, rows=33, channels=1, total_values=528 * 7 + 5, seed=14, dist=gen.DIST_STRESS, wavc=1),
    "l10r2_stress.acm": dict(level=10, rows=2, channels=2, total_values=2048 * 3 + 10, seed=15, dist=gen.DIST_STRESS),
    "l0r5_stereo_stall.acm": dict(level=0, rows=5, channels=2, total_values=23, seed=16, dist=gen.DIST_STRESS),
    "l1r1_mono.acm": dict(level=1, rows=1, channels=1, total_values=41, seed=17, dist=gen.DIST_STRESS),
    "l5r100_wavc.acm": dict(level=5, rows=100, channels=2, total_values=3200 * 2 + 1, seed=18, dist=gen.DIST_STRESS, wavc=1),
    "l6r3_odd_total_stereo.acm": dict(level=6, rows=3, channels=2, total_values=192 * 3 + 1, seed=19, dist=gen.DIST_STRESS),
    "bad_selector.acm": dict(level=5, rows=7, channels=1, total_values=224 * 4, seed=20, dist=gen.DIST_STRESS,
                             inject=gen.INJECT_BAD_IND, inject_block=2, inject_col=9, inject_value=25),
    "bad_tcode.acm": dict(level=4, rows=9, channels=1, total_values=144 * 4, seed=21, dist=gen.DIST_SINGLE,
                          single_ind=22, inject=gen.INJECT_BAD_TCODE, inject_block=1, inject_col=5),
}
TRUNCATED = {"truncated_mid_block.acm": ("l7r16_stress.acm", 3001), "truncated_header.acm": ("l7r16_stress.acm", 13)}
FORCE = {"l7r16_mono_fallout.acm": -1}


def main():
    ref = bindings.Ref()
    meta = {}
    images = {}
    for name, kw in CASES.items():
        images[name] = gen.make_stream(**kw)
    for name, (src, cut) in TRUNCATED.items():
        images[name] = images[src][:cut]
    for name, img in images.items():
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(img)
        fc = FORCE.get(name, 0)
        entry = {"force_chans": fc, "sha256": {}}
        for be in (0, 1):
            for sg in (0, 1):
                r = ref.decode(img, force_chans=fc, be=be, sgned=sg)
                entry["open_err"], entry["status"], entry["words"] = r.open_err, r.status, r.words
                entry["info"] = r.info.as_dict()
                entry["sha256"][f"b{be}s{sg}"] = hashlib.sha256(r.pcm.tobytes()).hexdigest()
        meta[name] = entry
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(meta), "fixtures,", sum(len(i) for i in images.values()), "bytes")


if __name__ == "__main__":
    main()
