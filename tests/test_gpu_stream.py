"""GPU (-m gpu): the resumable codec at sizes where the parallel tile kernels run (StreamDecoder::decode on
decode_wt_stream_kernel, StreamEncoder::encode on encode_kernel with carried state), state by state against the oracle,
and the device-pointer entry points qoipp_b200_stream_{encode,decode}_dev."""
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


def _same_state(a, b):
    return a.s.run == b.s.run and bytes(a.s.prev) == bytes(b.s.prev) and bytes(a.s.seen) == bytes(b.s.seen)


@pytest.mark.parametrize("kind", synth.CLASSES)
@pytest.mark.parametrize("ch", [3, 4])
def test_stream_decoder_large_chunks_state_by_state(ctx, kind, ch):
    """Chunks of tens of KB (many tiles per call), random output capacities: capacity cuts inside tiles and runs, incomplete
    ops at the end of a chunk, pending runs; (processed, written), pixels and the carried state equal the oracle's."""
    from qoipp_b200 import api

    import zlib

    rng = np.random.default_rng(zlib.crc32(f"{kind}/{ch}".encode()))  # the same chunking in every run
    w, h = 700, 400
    raw = synth.generate(kind, w, h, ch)
    q = Oracle.encode(raw, w, h, ch)
    for tgt in (0, 7 - ch):
        a, b = api.StreamDecoder(ctx), Oracle.StreamDecoder()
        assert a.initialize(q[:14], tgt)[0] == 0 and b.initialize(q[:14], tgt)[0] == 0
        off, got = 14, []
        end = q.size - 8
        for _ in range(100000):
            if off >= end:
                break
            cap = int(rng.integers(3000, 400000)) if rng.integers(0, 4) else int(rng.integers(4, 200))
            take = int(rng.integers(2000, 120000))
            oa, ob = np.full(cap + 32, 0xAA, np.uint8), np.full(cap + 32, 0xAA, np.uint8)
            chunk = q[off: min(off + take, end)]
            ra, rb = a.decode(oa[:cap], chunk), b.decode(ob[:cap], chunk)
            assert ra == rb, (kind, ch, off, cap, take, ra, rb)
            assert np.array_equal(oa[: ra[2]], ob[: rb[2]]) and (oa[cap:] == 0xAA).all()
            assert _same_state(a, b), (kind, ch, off, cap, take)
            off += ra[1]
            got.append(oa[: ra[2]].copy())
        out = np.zeros(4096, np.uint8)
        while a.has_run_count():
            e, n = a.drain_run(out)
            assert e == 0
            got.append(out[:n].copy())
        want = H.retarget(raw, ch, tgt)
        assert np.array_equal(np.concatenate(got), want), (kind, ch, tgt)


def test_stream_decoder_one_call_for_a_whole_image(ctx):
    """a 4K image in a single decode call: what the kernel writes equals the one-shot decode"""
    from qoipp_b200 import api

    w, h, ch = 3840, 2160, 3
    raw = synth.generate("photo", w, h, ch)
    q = Oracle.encode(raw, w, h, ch)
    d = api.StreamDecoder(ctx)
    assert d.initialize(q[:14])[0] == 0
    out = np.full(raw.size + 64, 0xAA, np.uint8)
    e, p, n = d.decode(out[: raw.size], q[14:-8])
    assert (e, p, n) == (0, q.size - 22, raw.size)
    assert np.array_equal(out[: raw.size], raw) and (out[raw.size:] == 0xAA).all()


def test_stream_dev_entry_points(ctx):
    """qoipp_b200_stream_encode_dev / _decode_dev: state, input, output and result all in device memory, one enqueue per call."""
    import torch

    from qoipp_b200._lib import lib

    w, h, ch = 900, 600, 4
    raw = synth.generate("photo", w, h, ch)
    ref = Oracle.encode(raw, w, h, ch)
    st = torch.cuda.current_stream().cuda_stream
    d_raw = torch.from_numpy(raw).cuda()

    def dev_state(prev, run, seen):
        a = np.zeros(66, dtype=np.uint32)
        a[0], a[1] = prev, run
        a[2:] = seen
        return torch.from_numpy(a.view(np.uint8)).cuda()

    # ---- encode in three calls of uneven size
    s_enc = dev_state(0xFF000000, 0, np.zeros(64, np.uint32))
    d_res = torch.zeros(16, dtype=torch.uint8, device="cuda")
    d_out = torch.full((raw.size * 2,), 0xAA, dtype=torch.uint8, device="cuda")
    body, off = [], 0
    for take in (w * 100 * ch, w * 333 * ch + 8, raw.size):
        take = min(take, raw.size - off) // ch * ch
        e = lib.qoipp_b200_stream_encode_dev(ctx._h, ch, C.c_void_p(s_enc.data_ptr()), C.c_void_p(d_raw[off:].data_ptr()), take, C.c_void_p(d_out.data_ptr()),
                                             d_out.numel(), C.c_void_p(d_res.data_ptr()), C.c_void_p(st))
        assert e == 0
        torch.cuda.synchronize()
        processed, written = (int(x) for x in d_res.cpu().numpy().view(np.uint64))
        assert processed == take
        body.append(d_out[:written].cpu().numpy())
        off += processed
    state = s_enc.cpu().numpy().view(np.uint32)
    tail = [0xC0 | (int(state[1]) - 1)] if state[1] else []
    stream = np.concatenate([ref[:14]] + body + [np.array(tail + [0, 0, 0, 0, 0, 0, 0, 1], np.uint8)])
    assert np.array_equal(stream, ref)

    # ---- decode in two calls; the second one finds an incomplete op at the end of the first
    seen = np.zeros(64, np.uint32)
    seen[53] = 0xFF000000
    s_dec = dev_state(0xFF000000, 0, seen)
    d_q = torch.from_numpy(ref[14:-8].copy()).cuda()
    d_px = torch.full((raw.size + 64,), 0xAA, dtype=torch.uint8, device="cuda")
    off = wr = 0
    for take in (d_q.numel() // 3 + 1, d_q.numel()):
        take = min(take, d_q.numel() - off)
        e = lib.qoipp_b200_stream_decode_dev(ctx._h, ch, C.c_void_p(s_dec.data_ptr()), C.c_void_p(d_q[off:].data_ptr()), take,
                                             C.c_void_p(d_px[wr:].data_ptr()), raw.size - wr, C.c_void_p(d_res.data_ptr()), C.c_void_p(st))
        assert e == 0
        torch.cuda.synchronize()
        processed, written = (int(x) for x in d_res.cpu().numpy().view(np.uint64))
        off += processed
        wr += written
    assert off == d_q.numel() and wr == raw.size
    assert np.array_equal(d_px[: raw.size].cpu().numpy(), raw) and bool((d_px[raw.size:] == 0xAA).all())
