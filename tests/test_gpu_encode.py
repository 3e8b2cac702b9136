"""GPU (-m gpu): the encode path through the C ABI (libqoipp_b200.so) against the oracle and, when it travelled
with the snapshot, the unmodified reference (oracle/_ref).  Replays the reference's own encode assertions
(test/source/simple_test.cpp:77-177, stream_test.cpp:192-201) and the adversarial sweep of SURVEY 8(d)."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle, Ref
from qoipp_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

FX = H.fixtures()
SMALL = [(1, 1), (1, 2), (1, 61), (1, 62), (1, 63), (1, 123), (1, 124), (1, 125), (29, 17), (24, 14), (2047, 1), (2048, 1),
         (2049, 1), (333, 77)]


@pytest.fixture(scope="module")
def ctx():
    from qoipp_b200 import api

    c = api.Context(0)
    yield c
    c.close()


def check(ctx, raw, w, h, ch, cs=0, cap=None):
    e, ref, rn, rok = Oracle.encode_into(raw, w, h, ch, cs, cap=cap)
    assert e == 0
    e, out, n, ok = ctx.encode_into(raw, w, h, ch, cs, cap=cap)
    assert e == 0
    assert (n, ok) == (rn, rok), (w, h, ch, cap, n, rn, ok, rok)
    if not np.array_equal(out[:n], ref[:n]):
        bad = int(np.nonzero(out[:n] != ref[:n])[0][0])
        raise AssertionError(f"{w}x{h}x{ch} cap={cap}: first differing byte {bad}: {out[bad-4:bad+8]} vs {ref[bad-4:bad+8]}")
    assert np.all(out[n:] == 0xAA), "bytes beyond `written` must stay untouched"


@pytest.mark.parametrize("ch", [3, 4])
def test_fixture_encode(ctx, ch):
    f = FX[ch]
    w, h, _, _ = f["desc"]
    assert np.array_equal(ctx.encode(f["raw"], w, h, ch), f["qoi"])
    e, out, n, ok = ctx.encode_into(f["raw"], w, h, ch, cap=H.CHUNK_BOUNDARY)
    assert e == 0 and not ok and n == H.CHUNK_BOUNDARY and np.array_equal(out[:n], f["qoi"][:n])


def test_committed_reference_vectors(ctx):
    v = H.ref_vectors()
    W, Hh = 37, 23
    for k in v.keys():
        parts = k.split("/")
        if parts[0] == "enc":
            kind, ch, cs = parts[1], int(parts[2]), int(parts[3])
            assert np.array_equal(ctx.encode(synth.generate(kind, W, Hh, ch), W, Hh, ch, cs), v[k]), k
        elif parts[0] == "partial":
            kind, ch, cap = parts[1], int(parts[2]), int(parts[3])
            blob = v[k]
            written, complete = (int(x) for x in blob[:16].view(np.uint64))
            e, out, n, ok = ctx.encode_into(synth.generate(kind, W, Hh, ch), W, Hh, ch, 0, cap=cap)
            assert (e, n, ok) == (0, written, bool(complete)), k
            assert np.array_equal(out[:n], blob[16:]), k


@pytest.mark.parametrize("kind", synth.CLASSES)
def test_classes_small_sizes(ctx, kind):
    for ch in (3, 4):
        for i, (w, h) in enumerate(SMALL):
            check(ctx, synth.generate(kind, w, h, ch), w, h, ch, cs=i & 1)


@pytest.mark.parametrize("kind", synth.CLASSES)
def test_classes_512(ctx, kind):
    for ch in (3, 4):
        check(ctx, synth.generate(kind, 512, 512, ch), 512, 512, ch)


@pytest.mark.parametrize("kind,w,h,ch", [("photo", 3840, 2160, 3), ("photo", 1920, 1080, 4), ("noise", 1920, 1080, 4),
                                         ("gradient", 1920, 1080, 3), ("flat", 1920, 1080, 4), ("long_runs", 3840, 2160, 4),
                                         ("dither", 1920, 1080, 3), ("palette", 1920, 1080, 4), ("resync", 1920, 1080, 3)])
def test_full_sizes(ctx, kind, w, h, ch):
    check(ctx, synth.generate(kind, w, h, ch), w, h, ch)


def test_reference_agrees_when_present(ctx):
    if not Ref.available():
        pytest.skip("oracle/_ref did not travel")
    for kind in ("photo", "hash_collide", "alpha_toggle"):
        for ch in (3, 4):
            raw = synth.generate(kind, 640, 360, ch)
            assert np.array_equal(ctx.encode(raw, 640, 360, ch), Ref.encode(raw, 640, 360, ch))


def test_partial_capacity(ctx):
    rng = np.random.default_rng(3)
    for it, kind in enumerate(synth.CLASSES):
        ch = 3 + (it & 1)
        w, h = 130, 67
        raw = synth.generate(kind, w, h, ch)
        full = Oracle.encode(raw, w, h, ch)
        caps = {0, 13, 14, 15, full.size - 9, full.size - 8, full.size - 1, full.size, full.size + 3}
        caps |= set(int(c) for c in rng.integers(14, full.size, size=8))
        for cap in sorted(c for c in caps if c >= 0):
            check(ctx, raw, w, h, ch, cap=cap)


def test_error_codes(ctx):  # order of source/simple.cpp:235-244
    raw = synth.generate("noise", 4, 4, 4)
    for args, want in [((raw[:0], 4, 4, 4), 1), ((raw, 0, 4, 4), 5), ((raw, 4, 4, 5), 5), ((raw[:-1], 4, 4, 4), 6), ((raw, 4, 4, 3), 6)]:
        assert ctx.encode_into(*args)[0] == want == Oracle.encode_into(*args)[0]


def test_batch_encode(ctx):
    import torch

    w, h, ch, B = 64, 48, 4, 37
    imgs = [synth.generate("photo", w, h, ch, seed=0x51F0 + k) for k in range(B)]
    d_raw = torch.from_numpy(np.concatenate(imgs)).cuda()
    stride = (ch + 1) * w * h + 22
    stride = (stride + 15) // 16 * 16
    d_out = torch.full((B * stride,), 0xAA, dtype=torch.uint8, device="cuda")
    d_written = torch.zeros(B, dtype=torch.int64, device="cuda")
    ctx.encode_batch_dev(d_raw, w * h * ch, B, w, h, ch, 0, d_out, stride, stride, d_written, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out, written = d_out.cpu().numpy(), d_written.cpu().numpy()
    for k in range(B):
        ref = Oracle.encode(imgs[k], w, h, ch)
        assert written[k] == ref.size and np.array_equal(out[k * stride: k * stride + ref.size], ref), k


@pytest.mark.parametrize("ch", [3, 4])
def test_stream_encoder_sweep(ctx, ch):  # stream_test.cpp:192-201, every buffer size 5..1024
    from qoipp_b200 import api

    f = FX[ch]
    enc = api.StreamEncoder(ctx)
    for size in range(5, 1025):
        got = H.stream_encode(enc, f["desc"], size, f["raw"])
        assert np.array_equal(got, f["qoi"]), size


def test_stream_encoder_state_by_state(ctx):
    from qoipp_b200 import api

    rng = np.random.default_rng(5)
    for it in range(24):
        kind = synth.CLASSES[it % len(synth.CLASSES)]
        ch = 3 + (it & 1)
        w, h = int(rng.integers(1, 200)), int(rng.integers(1, 60))
        raw = synth.generate(kind, w, h, ch, seed=77 + it)
        a, b = api.StreamEncoder(ctx), Oracle.StreamEncoder()
        hd = np.zeros(14, np.uint8)
        a.initialize(hd, w, h, ch)
        b.initialize(hd.copy(), w, h, ch)
        off = 0
        while off < raw.size:
            cap, take = int(rng.integers(5, 3000)), int(rng.integers(1, 9000))
            oa, ob = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
            ra, rb = a.encode(oa, raw[off: off + take]), b.encode(ob, raw[off: off + take])
            assert ra == rb, (kind, ch, off, cap, take, ra, rb)
            assert np.array_equal(oa[: ra[2]], ob[: rb[2]])
            assert a.s.run == b.s.run and bytes(a.s.prev) == bytes(b.s.prev) and bytes(a.s.seen) == bytes(b.s.seen)
            off += ra[1]
            if ra[1] == 0 and ra[2] == 0:
                break
