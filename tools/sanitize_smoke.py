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
print("sanitize smoke ok:", n, "image cases")
