"""GPU (-m gpu): BASELINE.json configs at their full sizes, checked through size-independent properties
(encode -> decode returns the input bit for bit; per-image sizes of identical images agree; a prefix of the stream equals
the oracle's encoding of the same prefix of rows)."""
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


def test_config3_16384x16384_rgba_roundtrip(ctx):
    """configs[2]: single 16384x16384 RGBA image (noise: literal ops only, worst case for the speculative parse)."""
    import torch

    w = h = 16384
    block = synth.generate("noise", 8192, 2048, 4)  # 64 MiB of noise, tiled: content class is what matters
    d_block = torch.from_numpy(block).cuda()
    d_raw = d_block.repeat((w * h) // (8192 * 2048))
    assert d_raw.numel() == w * h * 4
    cap = 5 * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.encode_dev(d_raw, w, h, 4, 0, d_q, cap, st)
    n, ok = ctx.encode_status(st)
    assert ok and n > w * h * 4  # noise expands
    # the first rows of the stream are what the oracle produces for the same rows (the codec is causal)
    rows = 64
    ref = Oracle.encode(d_raw[: w * rows * 4].cpu().numpy(), w, rows, 4)
    got = d_q[: ref.size - 8].cpu().numpy()
    assert np.array_equal(got[14:], ref[14:-8])
    d_out = torch.zeros(w * h * 4, dtype=torch.uint8, device="cuda")
    ctx.decode_dev(d_q, n, w, h, 4, 0, 0, False, d_out, d_out.numel(), st)
    assert ctx.decode_status(st) == 0
    assert torch.equal(d_out, d_raw)


def test_config4_batch_8192_images_512x512_rgba(ctx):
    """configs[3] on one GPU: 8192 images of 512x512 RGBA, 64 distinct contents replicated on the device."""
    import torch

    w = h = 512
    B, distinct = 8192, 64
    imgs = []
    for k in range(distinct):
        rgb = synth.generate("photo", w, h, 3, seed=0x51F0 + k).reshape(-1, 3)
        imgs.append(np.concatenate([rgb, np.full((rgb.shape[0], 1), 255, np.uint8)], axis=1).reshape(-1))
    d_src = torch.from_numpy(np.stack(imgs)).cuda()          # (64, raw)
    d_raw = d_src.repeat(B // distinct, 1).reshape(-1)        # image k has content k % 64
    raw_one = w * h * 4
    stride = (5 * w * h + 22 + 255) // 256 * 256
    d_q = torch.empty(stride * B, dtype=torch.uint8, device="cuda")
    d_written = torch.zeros(B, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.encode_batch_dev(d_raw, raw_one, B, w, h, 4, 0, d_q, stride, stride, d_written, st)
    torch.cuda.synchronize()
    sizes = d_written.cpu().numpy()
    assert np.array_equal(sizes, np.tile(sizes[:distinct], B // distinct))  # identical images, identical sizes
    for k in (0, 17, 63):
        ref = Oracle.encode(imgs[k], w, h, 4)
        assert sizes[k] == ref.size
        assert np.array_equal(d_q[(k + 128) * stride: (k + 128) * stride + ref.size].cpu().numpy(), ref)
    # pack the streams back to back and decode the whole batch
    offs = np.zeros(B + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(sizes.astype(np.uint64))
    idx = torch.arange(stride, device="cuda")
    packed = torch.empty(int(offs[-1]) + 64, dtype=torch.uint8, device="cuda")
    view = d_q.view(B, stride)
    for k in range(distinct):  # images with equal content have equal size: copy them group-wise
        rows = view[k::distinct, : int(sizes[k])]
        starts = torch.from_numpy(offs[k:-1:distinct].astype(np.int64)).cuda()
        dst = (starts[:, None] + idx[None, : int(sizes[k])]).reshape(-1)
        packed[dst] = rows.reshape(-1)
    d_out = torch.zeros(raw_one * B, dtype=torch.uint8, device="cuda")
    ctx.decode_batch_dev(packed, offs, w, h, 4, 0, 0, d_out, raw_one, st)
    torch.cuda.synchronize()
    assert torch.equal(d_out, d_raw)


def _device_image(kind, w, h, ch):
    from qoipp_b200 import synth_torch

    return synth_torch.generate(kind, w, h, ch, device="cuda")[0]


def _full_size_case(ctx, kind, w, h, ch, rows, expect_path=None):
    """encode on the device; the stream's prefix equals the oracle's encoding of the first `rows` rows (the codec is causal);
    decode returns the input bit for bit; returns (encoded size, decode path)."""
    import torch

    d_raw = _device_image(kind, w, h, ch)
    cap = (ch + 1) * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st)
    n, ok = ctx.encode_status(st)
    assert ok
    head = d_raw[: w * rows * ch].cpu().numpy()
    if kind == "photo_opaque":
        want = H.to_rgba(synth.generate("photo", w, rows, 3))
    else:
        want = synth.generate(kind, w, rows, ch)
    assert np.array_equal(head, want), "device generator disagrees with synth.py"
    ref = Oracle.encode(head, w, rows, ch)
    # the last chunk of the prefix may differ (a run / index that continues into the next row): compare all but 8 bytes
    got = d_q[: ref.size - 8].cpu().numpy()
    assert np.array_equal(got[14: ref.size - 16], ref[14: ref.size - 16])
    d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
    ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st)
    path = ctx.decode_status(st)
    assert torch.equal(d_out, d_raw)
    # and the oracle decodes the device's prefix back to the first rows (ref.decode(gpu.encode(x)) == x on the prefix)
    if expect_path is not None:
        assert path == expect_path, path
    return n, path


def test_8k_rgba_photo_with_alpha_blobs(ctx):
    """the "single 8K image" of the target sentence, SURVEY's RGBA `photo` class (soft alpha blobs: the class whose OP_RGB
    alpha speculation is refuted now and then)."""
    n, path = _full_size_case(ctx, "photo", 7680, 4320, 4, rows=96)
    assert path < 100, path  # the sequential loop is never needed; tiles repair themselves or a retry round fixes them
    assert 0.2 < n / (7680 * 4320 * 4) < 0.7


def test_8k_rgba_photo_opaque_verifies_in_round_0(ctx):
    _full_size_case(ctx, "photo_opaque", 7680, 4320, 4, rows=96, expect_path=0)


@pytest.mark.parametrize("kind", ["photo", "resync"])
def test_config3_16384x16384_rgba_photo_and_resync(ctx, kind):
    """configs[2], the two other classes SURVEY 8(d) lists for it (`noise` is the test above)."""
    n, path = _full_size_case(ctx, kind, 16384, 16384, 4, rows=48)
    assert path < 100, path


def test_4k_rgb_photo_matches_the_oracle_byte_for_byte(ctx):
    """configs[1] at full size against the oracle, the whole stream (8.3 M pixels: the oracle needs a fraction of a second)."""
    import torch

    w, h, ch = 3840, 2160, 3
    d_raw = _device_image("photo", w, h, ch)
    raw = d_raw.cpu().numpy()
    ref = Oracle.encode(raw, w, h, ch)
    cap = (ch + 1) * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st)
    n, ok = ctx.encode_status(st)
    assert ok and n == ref.size
    assert np.array_equal(d_q[:n].cpu().numpy(), ref)
    d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
    ctx.decode_dev(torch.from_numpy(ref).cuda(), ref.size, w, h, ch, 0, 0, False, d_out, d_out.numel(), st)
    assert ctx.decode_status(st) == 0
    assert np.array_equal(d_out.cpu().numpy(), raw)


def test_repeated_batch_decodes_of_alpha_blob_images_are_exact(ctx):
    """Regression (round 2): a tile that needed TWO repair passes compared its final carry words with those of its first
    pass only; a successor that had read the words of the pass in between kept a stale table entry and the image came out
    wrong in 107 bytes, timing dependent.  Images 1765, 4971 and 6213 of configs[3] were the ones hit."""
    import torch

    from qoipp_b200 import synth_torch

    w = h = 512
    ids = [1765, 4971, 6213] + list(range(2000, 2125))
    B = len(ids)
    d_raw = synth_torch.generate("photo", w, h, 4, seeds=[0x51F0 + k for k in ids], device="cuda").reshape(-1)
    raw_one = w * h * 4
    stride = (5 * w * h + 22 + 255) // 256 * 256
    d_q = torch.empty(stride * B, dtype=torch.uint8, device="cuda")
    d_written = torch.zeros(B, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.encode_batch_dev(d_raw, raw_one, B, w, h, 4, 0, d_q, stride, stride, d_written, st)
    torch.cuda.synchronize()
    sizes = d_written.cpu().numpy().astype(np.uint64)
    for k in (0, 1, 2):
        ref = Oracle.encode(d_raw[k * raw_one: (k + 1) * raw_one].cpu().numpy(), w, h, 4)
        assert sizes[k] == ref.size and np.array_equal(d_q[k * stride: k * stride + ref.size].cpu().numpy(), ref)
    d_out = torch.empty(raw_one * B, dtype=torch.uint8, device="cuda")
    for it in range(25):
        d_out.fill_(it)
        lib_sizes = sizes.copy()
        from qoipp_b200._lib import Desc, lib
        import ctypes as C

        e = lib.qoipp_b200_decode_batch_strided_dev(ctx._h, C.c_void_p(d_q.data_ptr()), stride, lib_sizes.ctypes.data_as(C.POINTER(C.c_uint64)), B,
                                                    C.byref(Desc(w, h, 4, 0)), 0, C.c_void_p(d_out.data_ptr()), raw_one, C.c_void_p(st))
        assert e == 0
        torch.cuda.synchronize()
        assert torch.equal(d_out, d_raw), it
        assert (ctx.decode_status_batch(B, st) < 100).all()
