"""Helpers for the GPU parity tests: pack images into a blob, decode through the C ABI,
compare stream by stream with the checker."""
from __future__ import annotations

import numpy as np

from libacm_b200 import api


def pack(images, align=16, lead=0, fill=0xFF):
    """Concatenate file images; `align`=1 packs them back to back (worst case for the
    device bit reader), gaps and tail are filled with garbage, not zeros."""
    offs, lens, at = [], [], lead
    for img in images:
        at = (at + align - 1) // align * align
        offs.append(at)
        lens.append(len(img))
        at += len(img)
    blob = np.full(at + 64, fill, np.uint8)
    for o, img in zip(offs, images):
        blob[o:o + len(img)] = np.frombuffer(bytes(img), np.uint8)
    return blob, np.array(offs, np.uint64), np.array(lens, np.uint32)


def decode_host(images, align=16, lead=0, **optkw):
    """One-shot acm_gpu_decode_batch with HOST buffers.  Returns (streams, out)."""
    blob, offs, lens = pack(images, align, lead)
    opts = api.make_opts(**optkw)
    s = api.new_streams(offs, lens)
    api.probe(blob, s, opts)
    nbytes = api.layout(s, opts.wordlen)
    out = np.full(nbytes + 16, 0xAA, np.uint8)
    api.decode_batch(blob, s, out, opts)
    return s, out


def decode_device(images, align=16, lead=0, **optkw):
    """Resident path: torch CUDA tensors + acm_gpu_plan_*.  Returns (streams, out ndarray)."""
    import torch
    blob, offs, lens = pack(images, align, lead)
    opts = api.make_opts(**optkw)
    d_blob = torch.from_numpy(blob).cuda()
    s = api.new_streams(offs, lens)
    api.probe(d_blob, s, opts)
    nbytes = api.layout(s, opts.wordlen)
    d_out = torch.full((nbytes + 16,), 0xAA, dtype=torch.uint8, device="cuda")
    plan = api.Plan(s, opts)
    plan.run(d_blob, d_out, torch.cuda.current_stream().cuda_stream)
    plan.fetch(s, torch.cuda.current_stream().cuda_stream)
    plan.close()
    return s, d_out.cpu().numpy()


def compare(images, s, out, checker, wordlen=2, be=0, sgned=1, force_chans=0, checksums=False):
    bad = []
    for i, img in enumerate(images):
        a = checker.decode(img, force_chans=force_chans, be=be, sgned=sgned)
        if a.open_err < 0:
            if s["status"][i] != a.open_err:
                bad.append((i, "open", int(s["status"][i]), a.open_err))
            continue
        if (int(s["status"][i]), int(s["words"][i])) != (a.status, a.words):
            bad.append((i, "status/words", int(s["status"][i]), int(s["words"][i]), a.status, a.words))
            continue
        o = int(s["out_off"][i])
        got = out[o:o + a.info.total_values * wordlen]
        if not np.array_equal(got, a.pcm):
            first = int(np.flatnonzero(got != a.pcm)[0])
            bad.append((i, "pcm", first // wordlen, a.info.acm_level, a.info.acm_rows))
        if checksums and int(s["checksum"][i]) != api.checksum_ref(a.pcm, a.words, wordlen, be):
            bad.append((i, "checksum"))
    return bad
