"""GPU (-m gpu): the page-locked (zero-copy) host path -- the path bench.py's `e2e` runs on -- against the oracle.

Page-locked buffers are handed to the kernels in place (csrc/qoipp_b200.cu mapped_host): the encode then takes the
general single-pass kernel reading / writing host memory over PCIe, the decode reads the stream and writes the pixels the
same way.  Everything the pageable tests check is checked again here on pinned buffers, incl. partial capacities and the
guard bytes beyond `written`."""
import ctypes as C

import numpy as np
import pytest

from oracle.pyoracle import Oracle
from qoipp_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from qoipp_b200 import api

    c = api.Context(0)
    yield c
    c.close()


def _pinned(n, fill=None):
    import torch

    t = torch.empty(max(int(n), 1), dtype=torch.uint8).pin_memory()
    if fill is not None:
        t.fill_(fill)
    return t


def _encode_pinned(ctx, raw, w, h, ch, cs=0, cap=None, pin_in=True, pin_out=True):
    from qoipp_b200._lib import Desc, lib

    if cap is None:
        cap = (ch + 1) * w * h + 22
    guard = 64
    t_in = _pinned(raw.size) if pin_in else None
    a_in = t_in.numpy() if pin_in else raw.copy()
    a_in[: raw.size] = raw
    t_out = _pinned(cap + guard, 0xAA) if pin_out else None
    a_out = t_out.numpy() if pin_out else np.full(cap + guard, 0xAA, np.uint8)
    written, complete = C.c_uint64(0), C.c_int32(0)
    e = lib.qoipp_b200_encode_host(ctx._h, C.c_void_p(a_in.ctypes.data), raw.size, C.byref(Desc(w, h, ch, cs)), C.c_void_p(a_out.ctypes.data), cap,
                                   C.byref(written), C.byref(complete))
    return e, a_out.copy(), written.value, bool(complete.value)


def _decode_pinned(ctx, qoi, n_out, target=0, flip=False, pin_in=True, pin_out=True):
    from qoipp_b200._lib import Desc, lib

    guard = 64
    t_in = _pinned(qoi.size) if pin_in else None
    a_in = t_in.numpy() if pin_in else qoi.copy()
    a_in[: qoi.size] = qoi
    t_out = _pinned(n_out + guard, 0xAA) if pin_out else None
    a_out = t_out.numpy() if pin_out else np.full(n_out + guard, 0xAA, np.uint8)
    d = Desc()
    e = lib.qoipp_b200_decode_host(ctx._h, C.c_void_p(a_in.ctypes.data), qoi.size, target, int(flip), C.c_void_p(a_out.ctypes.data), n_out, C.byref(d))
    return e, a_out.copy(), (d.width, d.height, d.channels, d.colorspace)


@pytest.mark.parametrize("kind", synth.CLASSES)
@pytest.mark.parametrize("ch", [3, 4])
def test_pinned_encode_and_decode_match_the_oracle(ctx, kind, ch):
    for (w, h) in ((1, 1), (29, 17), (640, 353)):
        raw = synth.generate(kind, w, h, ch)
        ref = Oracle.encode(raw, w, h, ch)
        e, out, n, ok = _encode_pinned(ctx, raw, w, h, ch)
        assert e == 0 and ok and n == ref.size
        assert np.array_equal(out[:n], ref)
        assert (out[n:] == 0xAA).all(), "bytes beyond `written` were touched"
        e, px, d = _decode_pinned(ctx, ref, raw.size)
        assert e == 0 and d[:3] == (w, h, ch)
        assert np.array_equal(px[: raw.size], raw)
        assert (px[raw.size:] == 0xAA).all()


@pytest.mark.parametrize("pin_in,pin_out", [(True, False), (False, True)])
def test_mixed_pinned_and_pageable_buffers(ctx, pin_in, pin_out):
    w, h, ch = 1000, 700, 4
    raw = synth.generate("photo", w, h, ch)
    ref = Oracle.encode(raw, w, h, ch)
    e, out, n, ok = _encode_pinned(ctx, raw, w, h, ch, pin_in=pin_in, pin_out=pin_out)
    assert e == 0 and ok and np.array_equal(out[:n], ref) and (out[n:] == 0xAA).all()
    e, px, _ = _decode_pinned(ctx, ref, raw.size, pin_in=pin_in, pin_out=pin_out)
    assert e == 0 and np.array_equal(px[: raw.size], raw) and (px[raw.size:] == 0xAA).all()


@pytest.mark.parametrize("ch", [3, 4])
def test_pinned_partial_capacity(ctx, ch):
    """encode_into with a short page-locked buffer: `written` = largest chunk boundary <= capacity (util.hpp:240-246),
    nothing beyond it is touched; incl. the reference's own 1007-byte boundary case (simple_test.cpp:24-25)."""
    f = H.fixtures()[ch]
    w, h, c, cs = f["desc"]
    for cap in (0, 13, 14, 15, 100, H.CHUNK_BOUNDARY, f["qoi"].size - 1, f["qoi"].size):
        e, out, n, ok = _encode_pinned(ctx, f["raw"], w, h, c, cs, cap=cap)
        eo, oo, no, oko = Oracle.encode_into(f["raw"], w, h, c, cs, cap=cap)
        assert (e, n, ok) == (eo, no, oko), cap
        assert np.array_equal(out[:n], oo[:n])
        assert (out[n:] == 0xAA).all(), cap
    w, h = 700, 300
    raw = synth.generate("photo", w, h, ch)
    ref = Oracle.encode(raw, w, h, ch)
    for cap in (ref.size // 3, ref.size - 9, ref.size - 1):
        e, out, n, ok = _encode_pinned(ctx, raw, w, h, ch, cap=cap)
        eo, oo, no, oko = Oracle.encode_into(raw, w, h, ch, cap=cap)
        assert (e, n, ok) == (eo, no, oko)
        assert np.array_equal(out[:n], oo[:n]) and (out[n:] == 0xAA).all()


@pytest.mark.parametrize("target,flip", [(3, False), (4, True), (0, True)])
def test_pinned_decode_targets_and_flip(ctx, target, flip):
    w, h, ch = 333, 211, 4
    raw = synth.generate("photo", w, h, ch)
    q = Oracle.encode(raw, w, h, ch)
    want = Oracle.decode(q, target, flip)
    e, px, d = _decode_pinned(ctx, q, w * h * 4, target, flip)
    assert e == 0 and np.array_equal(px[: want.size], want)


def test_pinned_decode_of_content_that_needs_retry_rounds(ctx):
    for kind in ("alpha_toggle", "hash_collide", "wrap", "photo"):
        w, h, ch = 1920, 1080, 4
        raw = synth.generate(kind, w, h, ch)
        q = Oracle.encode(raw, w, h, ch)
        e, px, _ = _decode_pinned(ctx, q, raw.size)
        assert e == 0 and np.array_equal(px[: raw.size], raw), kind


def test_batch_host_entry_points(ctx):
    """qoipp_b200_encode_batch_host / _decode_batch_host on pinned and on pageable buffers against the oracle."""
    from qoipp_b200._lib import Desc, lib

    w, h, ch, B = 200, 120, 4, 9
    raws = [synth.generate(["photo", "palette", "noise", "flat", "hash_collide"][k % 5], w, h, ch, seed=50 + k) for k in range(B)]
    refs = [Oracle.encode(r, w, h, ch) for r in raws]
    raw_one = w * h * ch
    stride = ((ch + 1) * w * h + 22 + 255) // 256 * 256
    for pinned in (True, False):
        t = [_pinned(raw_one * B), _pinned(stride * B, 0xAA), _pinned(raw_one * B, 0xAA)] if pinned else None
        a_raw = t[0].numpy() if pinned else np.empty(raw_one * B, np.uint8)
        a_q = t[1].numpy() if pinned else np.full(stride * B, 0xAA, np.uint8)
        a_out = t[2].numpy() if pinned else np.full(raw_one * B, 0xAA, np.uint8)
        a_raw[:] = np.concatenate(raws)
        written = np.zeros(B, dtype=np.uint64)
        e = lib.qoipp_b200_encode_batch_host(ctx._h, C.c_void_p(a_raw.ctypes.data), raw_one, B, C.byref(Desc(w, h, ch, 0)), C.c_void_p(a_q.ctypes.data), stride,
                                             stride, written.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert e == 0
        for k in range(B):
            assert written[k] == refs[k].size
            assert np.array_equal(a_q[k * stride: k * stride + refs[k].size], refs[k])
            assert (a_q[k * stride + refs[k].size: (k + 1) * stride] == 0xAA).all()
        e = lib.qoipp_b200_decode_batch_host(ctx._h, C.c_void_p(a_q.ctypes.data), stride, written.ctypes.data_as(C.POINTER(C.c_uint64)), B,
                                             C.byref(Desc(w, h, ch, 0)), 0, C.c_void_p(a_out.ctypes.data), raw_one)
        assert e == 0
        assert np.array_equal(a_out, np.concatenate(raws))
        paths = ctx.decode_status_batch(B, 0)
        assert paths.shape == (B,) and (paths < 100).all()


def test_staged_entry_points(ctx):
    """qoipp_b200_encode_staged / _decode_staged / _fetch_staged (what qoipp::encode / qoipp::decode of the C++ layer call):
    the result stays on the device until its size is known / while the caller allocates; error order of the reference."""
    from qoipp_b200._lib import Desc, lib

    for kind, w, h, ch in (("photo", 800, 450, 4), ("noise", 333, 111, 3), ("flat", 1, 1, 4)):
        raw = synth.generate(kind, w, h, ch)
        ref = Oracle.encode(raw, w, h, ch)
        written = C.c_uint64(0)
        e = lib.qoipp_b200_encode_staged(ctx._h, C.c_void_p(raw.ctypes.data), raw.size, C.byref(Desc(w, h, ch, 0)), C.byref(written))
        assert e == 0 and written.value == ref.size
        out = np.full(ref.size + 32, 0xAA, np.uint8)
        assert lib.qoipp_b200_fetch_staged(ctx._h, C.c_void_p(out.ctypes.data), ref.size + 1) == 7  # more than is staged: NotEnoughSpace
        assert lib.qoipp_b200_fetch_staged(ctx._h, C.c_void_p(out.ctypes.data), ref.size) == 0
        assert np.array_equal(out[: ref.size], ref) and (out[ref.size:] == 0xAA).all()
        for target, flip in ((0, False), (7 - ch, True)):
            d, need = Desc(), C.c_uint64(0)
            e = lib.qoipp_b200_decode_staged(ctx._h, C.c_void_p(ref.ctypes.data), ref.size, target, int(flip), C.byref(d), C.byref(need))
            want = Oracle.decode(ref, target, flip)
            assert e == 0 and need.value == want.size and d.channels == (target or ch)
            px = np.full(want.size + 32, 0xAA, np.uint8)
            assert lib.qoipp_b200_fetch_staged(ctx._h, C.c_void_p(px.ctypes.data), want.size) == 0
            assert np.array_equal(px[: want.size], want) and (px[want.size:] == 0xAA).all()
    raw = synth.generate("photo", 16, 16, 3)
    wr = C.c_uint64(0)
    assert lib.qoipp_b200_encode_staged(ctx._h, C.c_void_p(raw.ctypes.data), 0, C.byref(Desc(16, 16, 3, 0)), C.byref(wr)) == 1          # Empty
    assert lib.qoipp_b200_encode_staged(ctx._h, C.c_void_p(raw.ctypes.data), raw.size, C.byref(Desc(16, 16, 5, 0)), C.byref(wr)) == 5  # InvalidDesc
    assert lib.qoipp_b200_encode_staged(ctx._h, C.c_void_p(raw.ctypes.data), raw.size - 3, C.byref(Desc(16, 16, 3, 0)), C.byref(wr)) == 6  # MismatchedDesc
    q = Oracle.encode(raw, 16, 16, 3)
    d, need = Desc(), C.c_uint64(0)
    assert lib.qoipp_b200_decode_staged(ctx._h, C.c_void_p(q.ctypes.data), 0, 0, 0, C.byref(d), C.byref(need)) == 1    # Empty
    assert lib.qoipp_b200_decode_staged(ctx._h, C.c_void_p(q.ctypes.data), 22, 0, 0, C.byref(d), C.byref(need)) == 2   # TooShort
    bad = q.copy()
    bad[0] = ord("x")
    assert lib.qoipp_b200_decode_staged(ctx._h, C.c_void_p(bad.ctypes.data), bad.size, 0, 0, C.byref(d), C.byref(need)) == 4  # NotQoi
