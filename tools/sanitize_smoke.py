"""Small end-to-end exercise of every kernel for compute-sanitizer runs (memcheck / racecheck / synccheck):
compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from qoipp_b200 import api, synth
from oracle.pyoracle import Oracle
from tests import helpers as H

ctx = api.Context(0)
n = 0
for kind in synth.CLASSES:
    for ch in (3, 4):
        for (w, h) in ((1, 1), (1, 63), (97, 41), (300, 100)):
            raw = synth.generate(kind, w, h, ch)
            ref = Oracle.encode(raw, w, h, ch)
            got = ctx.encode(raw, w, h, ch)
            assert np.array_equal(got, ref), (kind, ch, w, h)
            for tgt in (0, 7 - ch):
                e, px, _ = ctx.decode_into(ref, tgt, flip=bool(n & 1))
                assert e == 0 and np.array_equal(px, Oracle.decode(ref, tgt, bool(n & 1))), (kind, ch, w, h, tgt)
            e, out, wr, ok = ctx.encode_into(raw, w, h, ch, cap=max(14, ref.size // 2))
            assert e == 0 and wr == Oracle.encode_into(raw, w, h, ch, cap=max(14, ref.size // 2))[2]
            n += 1
fx = H.fixtures()
enc, dec = api.StreamEncoder(ctx), api.StreamDecoder(ctx)
for ch in (3, 4):
    f = fx[ch]
    for size in (5, 17, 100, 1024):
        assert np.array_equal(H.stream_encode(enc, f["desc"], size, f["raw"]), f["qoi"])
        px, _ = H.stream_decode(dec, size, f["qoi"])
        assert np.array_equal(px, f["raw"])
    px, _ = H.stream_decode(dec, 64, f["qoi_incomplete"])
# the parallel resumable decode (inputs of a few KB and more), capacity cuts and incomplete ops included
rng = np.random.default_rng(3)
for kind, ch in (("photo", 4), ("noise", 3), ("long_runs", 4), ("hash_collide", 4)):
    w, h = 320, 180
    raw = synth.generate(kind, w, h, ch)
    q = Oracle.encode(raw, w, h, ch)
    a, b = api.StreamDecoder(ctx), Oracle.StreamDecoder()
    assert a.initialize(q[:14])[0] == 0 and b.initialize(q[:14])[0] == 0
    off = 14
    while off < q.size - 8:
        cap, take = int(rng.integers(3000, 90000)), int(rng.integers(4000, 40000))
        oa, ob = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        ra, rb = a.decode(oa, q[off: min(off + take, q.size - 8)]), b.decode(ob, q[off: min(off + take, q.size - 8)])
        assert ra == rb and np.array_equal(oa[: ra[2]], ob[: rb[2]]) and bytes(a.s.seen) == bytes(b.s.seen), (kind, ch, off)
        off += ra[1]
# a small batch through the batch entry points (alpha blobs: in-place repairs and retry rounds)
import torch
from qoipp_b200 import synth_torch
B, w, h = 24, 256, 256
d_raw = synth_torch.generate("photo", w, h, 4, seeds=[0x51F0 + 1765 + k for k in range(B)], device="cuda").reshape(-1)
stride = (5 * w * h + 22 + 255) // 256 * 256
d_q = torch.empty(stride * B, dtype=torch.uint8, device="cuda"); d_wr = torch.zeros(B, dtype=torch.int64, device="cuda")
ctx.encode_batch_dev(d_raw, w * h * 4, B, w, h, 4, 0, d_q, stride, stride, d_wr, 0); torch.cuda.synchronize()
sizes = d_wr.cpu().numpy().astype(np.uint64); offs = np.zeros(B + 1, np.uint64); offs[1:] = np.cumsum(sizes)
packed = torch.empty(int(offs[-1]) + 64, dtype=torch.uint8, device="cuda")
for k in range(B): packed[int(offs[k]): int(offs[k + 1])] = d_q[k * stride: k * stride + int(sizes[k])]
d_out = torch.zeros(w * h * 4 * B, dtype=torch.uint8, device="cuda")
ctx.decode_batch_dev(packed, offs, w, h, 4, 0, 0, d_out, w * h * 4, 0); torch.cuda.synchronize()
assert torch.equal(d_out, d_raw)
print("sanitize smoke ok:", n, "image cases + resumable + batch")
