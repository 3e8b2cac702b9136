"""GPU (-m gpu): the decode path through the C ABI against the oracle, the committed outputs of the unmodified
reference and (when it travelled) oracle/_ref.  Replays test/source/simple_test.cpp:179-242,316-322 and
stream_test.cpp:204-252, plus the adversarial sweep of SURVEY 8(d)."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle, Ref
from qoipp_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

FX = H.fixtures()
SMALL = [(1, 1), (1, 2), (1, 61), (1, 62), (1, 63), (1, 124), (29, 17), (24, 14), (2049, 3), (333, 77)]


@pytest.fixture(scope="module")
def ctx():
    from qoipp_b200 import api

    c = api.Context(0)
    yield c
    c.close()


def check(ctx, qoi, src_ch, target=0, flip=False):
    tgt = target or src_ch
    ref = Oracle.decode(qoi, tgt, flip)
    e, px, desc = ctx.decode_into(qoi, target, flip)
    assert e == 0, e
    assert desc[2] == tgt
    if not np.array_equal(px, ref):
        bad = int(np.nonzero(px != ref)[0][0]) // tgt
        raise AssertionError(f"src {src_ch} -> {tgt} flip={flip}: first wrong pixel {bad} of {ref.size // tgt}")


@pytest.mark.parametrize("ch", [3, 4])
def test_fixture_decode(ctx, ch):
    f = FX[ch]
    for target in (0, 3, 4):
        e, px, desc = ctx.decode_into(f["qoi"], target)
        assert e == 0 and desc == (f["desc"][0], f["desc"][1], target or ch, 0)
        assert np.array_equal(px, H.retarget(f["raw"], ch, target))
    check(ctx, f["qoi_incomplete"], ch)


def test_committed_reference_vectors(ctx):
    v = H.ref_vectors()
    n = 0
    for k in v.keys():
        parts = k.split("/")
        if parts[0] == "adv" and parts[2] == "in":
            name = parts[1]
            target = int(v[f"adv/{name}/target"][0])
            e, px, _ = ctx.decode_into(v[k], target)
            assert e == 0 and np.array_equal(px, v[f"adv/{name}/out"]), name
            n += 1
        elif parts[0] == "enc":
            kind, ch = parts[1], int(parts[2])
            raw = synth.generate(kind, 37, 23, ch)
            for target in (0, 3, 4):
                e, px, _ = ctx.decode_into(v[k], target)
                assert e == 0 and np.array_equal(px, H.retarget(raw, ch, target)), k
    assert n >= 39


@pytest.mark.parametrize("kind", synth.CLASSES)
def test_classes_small_sizes(ctx, kind):
    for ch in (3, 4):
        for i, (w, h) in enumerate(SMALL):
            q = Oracle.encode(synth.generate(kind, w, h, ch), w, h, ch)
            check(ctx, q, ch, target=[0, 3, 4][i % 3], flip=bool(i & 1))


@pytest.mark.parametrize("kind", synth.CLASSES)
def test_classes_512(ctx, kind):
    for ch in (3, 4):
        raw = synth.generate(kind, 512, 512, ch)
        q = Oracle.encode(raw, 512, 512, ch)
        e, px, _ = ctx.decode_into(q)
        assert e == 0 and np.array_equal(px, raw)


@pytest.mark.parametrize("kind,w,h,ch", [("photo", 3840, 2160, 3), ("photo", 1920, 1080, 4), ("noise", 1920, 1080, 4),
                                         ("gradient", 1920, 1080, 3), ("flat", 1920, 1080, 4), ("long_runs", 3840, 2160, 4),
                                         ("dither", 1920, 1080, 3), ("palette", 1920, 1080, 3), ("resync", 1920, 1080, 3),
                                         ("resync", 1920, 1080, 4)])
def test_full_sizes_roundtrip(ctx, kind, w, h, ch):
    raw = synth.generate(kind, w, h, ch)
    q = ctx.encode(raw, w, h, ch)  # GPU encode -> GPU decode -> original, and the oracle agrees on the stream
    assert np.array_equal(q, Oracle.encode(raw, w, h, ch))
    e, px, _ = ctx.decode_into(q)
    assert e == 0 and np.array_equal(px, raw)


def test_opaque_content_uses_the_parallel_path(ctx):
    import torch

    for kind, ch in (("photo", 3), ("dither", 3), ("palette", 3), ("noise", 4), ("resync", 4)):
        w, h = 640, 360
        raw = synth.generate(kind, w, h, ch)
        q = Oracle.encode(raw, w, h, ch)
        d_q = torch.from_numpy(q).cuda()
        d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        ctx.decode_dev(d_q, q.size, w, h, ch, 0, 0, False, d_out, d_out.numel(), st)
        assert ctx.decode_status(st) == 0, kind
        assert np.array_equal(d_out.cpu().numpy(), raw)


def test_reference_agrees_when_present(ctx):
    if not Ref.available():
        pytest.skip("oracle/_ref did not travel")
    for kind in ("photo", "hash_collide", "alpha_toggle", "wrap"):
        for ch in (3, 4):
            raw = synth.generate(kind, 640, 360, ch)
            q = Ref.encode(raw, 640, 360, ch)
            for target in (0, 3, 4):
                e, px, _ = ctx.decode_into(q, target, True)
                assert e == 0 and np.array_equal(px, Ref.decode(q, target, True))


def test_truncated_streams(ctx):
    for kind in ("photo", "palette", "long_runs"):
        for ch in (3, 4):
            w, h = 200, 111
            q = Oracle.encode(synth.generate(kind, w, h, ch), w, h, ch)
            for cut in (q.size - 8, q.size - 9, q.size - 11, q.size // 2, 40, 23):
                check(ctx, q[:cut], ch)


def test_random_op_soup(ctx):
    rng = np.random.default_rng(99)
    tags = np.array([0, 1, 5, 53, 0x6A, 0x55, 0x7F, 0xA0, 0x88, 0x11, 0xC1, 0xC5, 38, 17, 0xFE, 0xFF, 0x80, 0x3F], dtype=np.uint8)
    for it in range(40):
        nb = int(rng.integers(1, 30000))
        body = rng.choice(tags, size=nb) if it % 2 else rng.integers(0, 256, size=nb, dtype=np.uint8)
        ch = 3 + (it & 1)
        w, h = 257, int(rng.integers(1, 300))
        hdr = np.frombuffer(b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([ch, 0]), dtype=np.uint8)
        q = np.concatenate([hdr, body.astype(np.uint8), np.array([0, 0, 0, 0, 0, 0, 0, 1], np.uint8)])
        check(ctx, q, ch, target=[0, 3, 4][it % 3], flip=bool(it & 2))


def test_error_codes(ctx):  # order of source/simple.cpp:451-474
    raw = synth.generate("noise", 4, 4, 4)
    q = Oracle.encode(raw, 4, 4, 4)
    for bad in (q[:0], q[:10], q[:22]):
        assert ctx.decode_into(bad)[0] == Oracle.decode_into(bad)[0] != 0
    nq = q.copy()
    nq[0] = 0
    assert ctx.decode_into(nq)[0] == 4
    nq = q.copy()
    nq[12] = 7
    assert ctx.decode_into(nq)[0] == 5
    assert ctx.decode_into(q, cap=4 * 4 * 4 - 1)[0] == 7


def test_batch_decode(ctx):
    import torch

    w, h, ch, B = 64, 48, 4, 41
    kinds = ["photo", "palette", "hash_collide", "noise", "flat"]
    raws = [synth.generate(kinds[k % 5], w, h, ch, seed=100 + k) for k in range(B)]
    qs = [Oracle.encode(r, w, h, ch) for r in raws]
    offs = np.zeros(B + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([q.size for q in qs])
    d_q = torch.from_numpy(np.concatenate(qs)).cuda()
    for target in (4, 3):
        stride = w * h * target
        d_out = torch.zeros(B * stride, dtype=torch.uint8, device="cuda")
        ctx.decode_batch_dev(d_q, offs, w, h, ch, 0, target, d_out, stride, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out = d_out.cpu().numpy()
        for k in range(B):
            assert np.array_equal(out[k * stride: (k + 1) * stride], H.retarget(raws[k], ch, target)), k


@pytest.mark.parametrize("ch", [3, 4])
def test_stream_decoder_sweep(ctx, ch):  # stream_test.cpp:204-252, every buffer size 5..1024
    from qoipp_b200 import api

    f = FX[ch]
    dec = api.StreamDecoder(ctx)
    for size in range(5, 1025):
        targets = (0, 3, 4) if size % 8 == 5 else (0,)
        for target in targets:
            px, desc = H.stream_decode(dec, size, f["qoi"], target)
            assert np.array_equal(px, H.retarget(f["raw"], ch, target)), (size, target)
        if size % 4 == 1:
            px, _ = H.stream_decode(dec, size, f["qoi_incomplete"])
            assert px.size != f["raw"].size and np.array_equal(px, f["raw"][: px.size]), size


def test_alpha_changing_index_ops_use_retry_rounds_not_the_sequential_kernel(ctx):
    """decode_status path: 0 = first pass verified (possibly after a tile repaired itself in place), 1..4 = retry rounds
    used, >= 100 = sequential kernel."""
    import torch

    for kind in ("hash_collide", "wrap", "alpha_toggle"):
        w, h, ch = 1920, 1080, 4
        raw = synth.generate(kind, w, h, ch)
        q = Oracle.encode(raw, w, h, ch)
        d_q = torch.from_numpy(q).cuda()
        d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        ctx.decode_dev(d_q, q.size, w, h, ch, 0, 0, False, d_out, d_out.numel(), st)
        path = ctx.decode_status(st)
        assert np.array_equal(d_out.cpu().numpy(), raw), kind
        assert path < 100, (kind, path)
