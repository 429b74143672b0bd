"""GPU tier: the reference's own CLI (acmtool.c, unmodified) linked against libacm_b200.so must write
the same files as the reference build of acmtool (BASELINE config 1)."""
import os
import subprocess

import pytest

from libacm_b200 import gen
from oracle import bindings

REF = bindings.REF_ACMTOOL
OURS = os.path.join(os.path.dirname(REF), "acmtool_b200")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(OURS)), reason="oracle/_ref not built")]

CASES = {
    "cfg1_stereo_60s": dict(level=7, rows=16, channels=2, rate=22050, total_values=2_646_000, seed=1),
    "mono_wavc": dict(level=6, rows=9, channels=1, rate=22050, total_values=50_001, seed=2, wavc=1, dist=gen.DIST_STRESS),
    "odd_total_stereo": dict(level=5, rows=3, channels=2, rate=11025, total_values=9_999, seed=3, dist=gen.DIST_STRESS),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("flags", [["-d", "-r", "-q"], ["-d", "-q"], ["-d", "-q", "-m"], ["-d", "-q", "-s", "-r"]])
def test_same_output_files(tmp_path, name, flags):
    src = tmp_path / (name + ".acm")
    src.write_bytes(gen.make_stream(**CASES[name]))
    outs = []
    for tool, tag in ((REF, "ref"), (OURS, "b200")):
        out = tmp_path / f"{name}.{tag}.out"
        r = subprocess.run([tool, *flags, "-o", str(out), str(src)], capture_output=True, text=True, timeout=300)
        outs.append((r.returncode, out.read_bytes(), r.stderr))
    assert outs[0][0] == outs[1][0] == 0
    assert outs[0][1] == outs[1][1]
    assert len(outs[0][1]) > 0


def test_truncated_file_same_padding_and_message(tmp_path):
    img = gen.make_stream(**CASES["mono_wavc"])
    src = tmp_path / "cut.acm"
    src.write_bytes(img[: len(img) // 2])
    res = []
    for tool, tag in ((REF, "ref"), (OURS, "b200")):
        out = tmp_path / f"cut.{tag}.raw"
        r = subprocess.run([tool, "-d", "-r", "-o", str(out), str(src)], capture_output=True, text=True, timeout=300)
        res.append((r.returncode, out.read_bytes(), r.stderr.replace(tag, "")))
    assert res[0][:2] == res[1][:2]          # same bytes incl. acmtool's zero "filler_samples" (acmtool.c:293-310)


def test_info_line(tmp_path):
    src = tmp_path / "i.acm"
    src.write_bytes(gen.make_stream(**CASES["cfg1_stereo_60s"]))
    a = subprocess.run([REF, "-i", str(src)], capture_output=True, text=True)
    b = subprocess.run([OURS, "-i", str(src)], capture_output=True, text=True)
    assert a.stdout == b.stdout and a.returncode == b.returncode == 0
