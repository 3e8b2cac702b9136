"""GPU (-m gpu): the opt-in thread-serial decode fast path (QOIPP_B200_DECODE_TS=1, qoipp_b200/csrc/decode_ts.cuh) through the
C ABI against the oracle: verified content stays on it (path < 1000), refuted content is decoded by the general machinery."""
import os

import numpy as np
import pytest

from oracle.pyoracle import Oracle
from qoipp_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from qoipp_b200 import api

    os.environ["QOIPP_B200_DECODE_TS"] = "1"  # read when the context is created
    try:
        c = api.Context(0)
    finally:
        del os.environ["QOIPP_B200_DECODE_TS"]
    yield c
    c.close()


def decode_dev(ctx, q, w, h, ch, target=0):
    import torch

    tgt = target or ch
    d_q = torch.from_numpy(np.ascontiguousarray(q)).cuda()
    d_out = torch.full((w * h * tgt + 64,), 0xAA, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.decode_dev(d_q, q.size, w, h, ch, 0, target, False, d_out, w * h * tgt, st)
    path = ctx.decode_status(st)
    out = d_out.cpu().numpy()
    assert np.all(out[w * h * tgt:] == 0xAA), "decode wrote past the image"
    return out[: w * h * tgt], path


@pytest.mark.parametrize("kind", synth.CLASSES)
def test_classes(ctx, kind):
    for ch in (3, 4):
        for (w, h), target in (((333, 77), 0), ((512, 512), 3), ((640, 481), 4)):
            raw = synth.generate(kind, w, h, ch)
            q = Oracle.encode(raw, w, h, ch)
            px, path = decode_dev(ctx, q, w, h, ch, target)
            assert np.array_equal(px, Oracle.decode(q, target or ch, False)), (kind, ch, w, h, target, path)


def test_opaque_photo_is_verified_by_the_fast_path(ctx):
    for ch in (3, 4):
        w, h = 1920, 1080
        raw = synth.generate("photo", w, h, 3)
        if ch == 4:
            raw = np.concatenate([raw.reshape(-1, 3), np.full((w * h, 1), 255, np.uint8)], axis=1).reshape(-1)
        q = Oracle.encode(raw, w, h, ch)
        px, path = decode_dev(ctx, q, w, h, ch)
        assert np.array_equal(px, raw) and path == 0, path


def test_refuted_and_truncated_streams_fall_back(ctx):
    w, h = 300, 200
    raw = synth.generate("alpha_toggle", w, h, 4)
    q = Oracle.encode(raw, w, h, 4)
    px, path = decode_dev(ctx, q, w, h, 4)
    assert np.array_equal(px, raw)
    raw = synth.generate("photo", w, h, 3)
    q = Oracle.encode(raw, w, h, 3)
    cut = q[: q.size // 2]
    px, path = decode_dev(ctx, cut, w, h, 3)
    assert np.array_equal(px, Oracle.decode(cut, 3, False)) and path >= 1000, path


def test_batch(ctx):
    import torch

    w, h, ch, n = 256, 192, 4, 7
    raws = [synth.generate(["photo", "noise", "dither"][k % 3], w, h, ch, seed=50 + k) for k in range(n)]
    for r in raws:
        r[3::4] = 255
    qs = [Oracle.encode(r, w, h, ch) for r in raws]
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([q.size for q in qs])
    d_q = torch.from_numpy(np.concatenate(qs)).cuda()
    stride = w * h * ch
    d_out = torch.zeros(stride * n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.decode_batch_dev(d_q, offs, w, h, ch, 0, 0, d_out, stride, st)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    for k in range(n):
        assert np.array_equal(out[k * stride: (k + 1) * stride], raws[k]), k
