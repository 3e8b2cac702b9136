"""CPU: the product DECODE kernels (qoipp_b200/csrc/decode_kernel.cuh) stepped by the SIMT emulator and compared
with the oracle / the committed outputs of the unmodified reference."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle
from qoipp_b200 import synth
from tests import emu_lib as E
from tests import helpers as H

FX = H.fixtures()
SIZES = [(1, 1), (1, 2), (1, 61), (1, 62), (1, 63), (1, 124), (29, 17), (24, 14), (255, 3), (100, 41)]


def check(qoi, w, h, src_ch, target=0, flip=False, expect_path=None, **kw):
    tgt = target or src_ch
    ref = Oracle.decode(qoi, tgt, flip)
    px, path = E.decode(qoi, w, h, tgt, flip=flip, **kw)
    if not np.array_equal(px[0], ref):
        bad = int(np.nonzero(px[0] != ref)[0][0]) // tgt
        raise AssertionError(f"{w}x{h} src {src_ch} -> {tgt} flip={flip}: first wrong pixel {bad} (path {path})")
    if expect_path == 'serial':
        assert path[0] >= 100
    elif expect_path is not None:
        assert path[0] == expect_path
    return path[0]


@pytest.mark.parametrize("ch", [3, 4])
def test_fixtures(ch):  # simple_test.cpp:179-223, 316-322
    f = FX[ch]
    w, h, _, _ = f["desc"]
    for target in (0, 3, 4):
        px, path = E.decode(f["qoi"], w, h, target or ch)
        assert np.array_equal(px[0], H.retarget(f["raw"], ch, target))
    check(f["qoi_incomplete"], w, h, ch)  # truncated: decodes the zero padding exactly like the reference


@pytest.mark.parametrize("kind", synth.CLASSES)
def test_classes_and_sizes(kind):
    for ch in (3, 4):
        for i, (w, h) in enumerate(SIZES):
            raw = synth.generate(kind, w, h, ch)
            q = Oracle.encode(raw, w, h, ch)
            check(q, w, h, ch, target=[0, 3, 4][i % 3], flip=bool(i & 1), seed=i)


@pytest.mark.parametrize("kind", ["photo", "dither", "palette", "noise", "resync", "gradient", "long_runs"])
def test_opaque_content_stays_on_the_parallel_path(kind):
    for ch in (3, 4):
        w, h = 160, 100
        raw = synth.generate(kind, w, h, ch)
        if ch == 4:
            raw = raw.copy()
            raw[3::4] = 255
        q = Oracle.encode(raw, w, h, ch)
        for seed, resident in ((0, 1), (2, 3), (5, 8)):
            check(q, w, h, ch, seed=seed, resident=resident, expect_path=0)


def test_truncated_and_padded_streams():
    for kind in ("photo", "palette", "long_runs"):
        for ch in (3, 4):
            w, h = 64, 37
            raw = synth.generate(kind, w, h, ch)
            q = Oracle.encode(raw, w, h, ch)
            for cut in (q.size - 8, q.size - 9, q.size - 11, q.size // 2, 40, 23):
                if cut > 22:
                    check(q[:cut], w, h, ch)


def test_committed_adversarial_streams():
    v = H.ref_vectors()
    n = 0
    for k in v.keys():
        parts = k.split("/")
        if parts[0] == "adv" and parts[2] == "in":
            name = parts[1]
            q = v[k]
            target = int(v[f"adv/{name}/target"][0])
            e, (w, h, ch, cs) = Oracle.read_header(q)
            assert e == 0
            tgt = target or ch
            px, path = E.decode(q, w, h, tgt)
            assert np.array_equal(px[0], v[f"adv/{name}/out"]), (name, path)
            n += 1
    assert n >= 39


def test_random_op_soup():
    rng = np.random.default_rng(99)
    tags = np.array([0, 1, 5, 53, 0x6A, 0x55, 0x7F, 0xA0, 0x88, 0x11, 0xC1, 0xC5, 38, 17, 0xFE, 0xFF, 0x80, 0x3F], dtype=np.uint8)
    for it in range(60):
        nb = int(rng.integers(1, 5000))
        body = rng.choice(tags, size=nb) if it % 2 else rng.integers(0, 256, size=nb, dtype=np.uint8)
        ch = 3 + (it & 1)
        w, h = 97, int(rng.integers(1, 80))
        hdr = np.frombuffer(b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([ch, 0]), dtype=np.uint8)
        q = np.concatenate([hdr, body.astype(np.uint8), np.array([0, 0, 0, 0, 0, 0, 0, 1], np.uint8)])
        check(q, w, h, ch, target=[0, 3, 4][it % 3], seed=it)


def test_forced_sequential_kernel():
    for kind in ("photo", "hash_collide", "alpha_toggle"):
        for ch in (3, 4):
            w, h = 80, 45
            raw = synth.generate(kind, w, h, ch)
            q = Oracle.encode(raw, w, h, ch)
            for target in (3, 4):
                check(q, w, h, ch, target=target, flip=True, force_serial=True, expect_path='serial')


def test_batch_decode():
    w, h, ch = 40, 30, 4
    raws = [synth.generate(["photo", "palette", "hash_collide", "noise", "flat"][k % 5], w, h, ch, seed=100 + k) for k in range(9)]
    qs = [Oracle.encode(r, w, h, ch) for r in raws]
    px, path = E.decode(qs, w, h, 4, seed=4)
    for k in range(9):
        assert np.array_equal(px[k], raws[k]), (k, path)
    px, path = E.decode(qs, w, h, 3, seed=5)
    for k in range(9):
        assert np.array_equal(px[k], H.to_rgb(raws[k])), (k, path)


@pytest.mark.parametrize("ch", [3, 4])
def test_stream_decoder_sweep(ch):  # stream_test.cpp:204-252 (subset of sizes; the GPU test runs them all)
    f = FX[ch]
    dec = E.StreamDecoder()
    for size in list(range(5, 40)) + list(range(40, 1025, 29)):
        for target in (0, 3, 4):
            px, desc = H.stream_decode(dec, size, f["qoi"], target)
            assert np.array_equal(px, H.retarget(f["raw"], ch, target)), (size, target)
        px, _ = H.stream_decode(dec, size, f["qoi_incomplete"])
        assert px.size != f["raw"].size and np.array_equal(px, f["raw"][: px.size]), size


def test_stream_decoder_state_by_state():
    rng = np.random.default_rng(8)
    for it in range(30):
        kind = synth.CLASSES[it % len(synth.CLASSES)]
        ch = 3 + (it & 1)
        w, h = int(rng.integers(1, 60)), int(rng.integers(1, 40))
        raw = synth.generate(kind, w, h, ch, seed=300 + it)
        q = Oracle.encode(raw, w, h, ch)
        a, b = E.StreamDecoder(), Oracle.StreamDecoder()
        tgt = [0, 3, 4][it % 3]
        assert a.initialize(q[:14], tgt)[0] == 0 and b.initialize(q[:14], tgt)[0] == 0
        off = 14
        for _ in range(10000):
            if off >= q.size:
                break
            cap, take = int(rng.integers(4, 300)), int(rng.integers(1, 200))
            oa, ob = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
            ra, rb = a.decode(oa, q[off: off + take]), b.decode(ob, q[off: off + take])
            assert ra == rb, (kind, ch, off, cap, take, ra, rb)
            assert np.array_equal(oa[: ra[2]], ob[: rb[2]])
            assert a.s.run == b.s.run and bytes(a.s.prev) == bytes(b.s.prev) and bytes(a.s.seen) == bytes(b.s.seen)
            off += ra[1]


def test_alpha_changing_index_ops_converge_in_retry_rounds():
    """INDEX ops that change alpha refute the round-0 speculation; the retry rounds learn the alphas and converge
    without the sequential kernel (path = number of rounds used, < 100)."""
    for kind in ("hash_collide", "wrap", "alpha_toggle", "palette", "photo"):
        for (w, h) in ((96, 64), (200, 150)):
            raw = synth.generate(kind, w, h, 4)
            q = Oracle.encode(raw, w, h, 4)
            for target in (4, 3):
                ref = Oracle.decode(q, target)
                px, path = E.decode(q, w, h, target, seed=w)
                assert np.array_equal(px[0], ref), (kind, w, h, target, path)
                assert path[0] < 100, (kind, w, h, path)


def test_restart_of_the_sequential_kernel_mid_stream():
    """A stream that no retry round can verify (INDEX of a never-written slot late in the stream): the sequential
    kernel resumes behind the last verified tile instead of at the image start."""
    w, h = 120, 90
    raw = synth.generate("photo", w, h, 3)
    q = Oracle.encode(raw, w, h, 3)
    body = q[14:-8].copy()
    # overwrite an op near the end with INDEX 9 followed by the same bytes: slot 9 may or may not have been written,
    # so check against the oracle rather than against `raw`
    cut = body.size - 300
    bad = np.concatenate([q[:14], body[:cut], np.array([9, 0x3F, 0x11], np.uint8), body[cut:], q[-8:]])
    ref = Oracle.decode(bad, 3)
    px, path = E.decode(bad, w, h, 3, seed=2)
    assert np.array_equal(px[0], ref), path


@pytest.mark.parametrize("ch", [3, 4])
def test_parallel_stream_decoder_sweep(ch):
    """stream_test.cpp:204-252 with the tile kernel behind every call (the device library uses it from a few KB on)."""
    f = FX[ch]
    dec = E.StreamDecoder(parallel=True)
    for size in list(range(5, 24)) + list(range(24, 1025, 53)):
        for target in (0, 3, 4):
            px, desc = H.stream_decode(dec, size, f["qoi"], target)
            assert np.array_equal(px, H.retarget(f["raw"], ch, target)), (size, target)
        px, _ = H.stream_decode(dec, size, f["qoi_incomplete"])
        assert px.size != f["raw"].size and np.array_equal(px, f["raw"][: px.size]), size


def test_parallel_stream_decoder_state_by_state():
    """Random input / output sizes over multi-tile streams: (processed, written), the pixels and the complete carried state
    equal the oracle's after every call -- capacity cuts inside tiles and runs, incomplete ops at the end of the input,
    pending runs longer than the output, tiles behind the cut."""
    rng = np.random.default_rng(18)
    serial = calls = 0
    for it in range(36):
        kind = synth.CLASSES[it % len(synth.CLASSES)]
        ch = 3 + (it & 1)
        w, h = int(rng.integers(20, 160)), int(rng.integers(10, 90))
        raw = synth.generate(kind, w, h, ch, seed=900 + it)
        q = Oracle.encode(raw, w, h, ch)
        a, b = E.StreamDecoder(parallel=True, resident=int(rng.integers(1, 4)), seed=it), Oracle.StreamDecoder()
        tgt = [0, 3, 4][it % 3]
        assert a.initialize(q[:14], tgt)[0] == 0 and b.initialize(q[:14], tgt)[0] == 0
        off = 14
        big = it % 3 == 0
        for _ in range(10000):
            if off >= q.size:
                break
            cap = int(rng.integers(4, 300)) if not big else int(rng.integers(2000, 40000))
            take = int(rng.integers(1, 200)) if not big else int(rng.integers(500, 9000))
            oa, ob = np.full(cap + 16, 0xAA, np.uint8), np.full(cap + 16, 0xAA, np.uint8)
            ra, rb = a.decode(oa[:cap], q[off: off + take]), b.decode(ob[:cap], q[off: off + take])
            assert ra == rb, (kind, ch, off, cap, take, ra, rb)
            assert np.array_equal(oa[: ra[2]], ob[: rb[2]]), (kind, ch, off, cap, take)
            assert (oa[cap:] == 0xAA).all()
            assert a.s.run == b.s.run and bytes(a.s.prev) == bytes(b.s.prev) and bytes(a.s.seen) == bytes(b.s.seen), (kind, ch, off, cap, take)
            off += ra[1]
            calls += 1
        serial += a.serial_calls
    assert serial < calls / 2, (serial, calls)  # the tile kernel's result stands for most calls


def test_parallel_stream_decoder_incomplete_op_across_a_tile_boundary():
    """RGBA noise: every op is 5 bytes, the op at byte 895 starts in tile 0 and ends in tile 1.  Inputs that end inside it
    (896..899 bytes) must leave it for the next call although the tile that holds its first byte is not the last one."""
    w, h = 40, 20
    rng = np.random.default_rng(5)
    body = rng.integers(0, 256, size=(w * h, 5), dtype=np.uint8)
    body[:, 0] = 0xFF  # OP_RGBA r g b a, 800 times
    body = body.reshape(-1)
    q = np.concatenate([np.frombuffer(b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([4, 0]), dtype=np.uint8), body,
                        np.array([0, 0, 0, 0, 0, 0, 0, 1], np.uint8)])
    for n in list(range(888, 908)) + [1791, 1792, 1793, 1794, 1795, 1796]:
        for cap in (100000, 4 * 150):
            a, b = E.StreamDecoder(parallel=True), Oracle.StreamDecoder()
            assert a.initialize(q[:14])[0] == 0 and b.initialize(q[:14])[0] == 0
            oa, ob = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
            ra, rb = a.decode(oa, body[:n]), b.decode(ob, body[:n])
            assert ra == rb, (n, cap, ra, rb)
            assert np.array_equal(oa[: ra[2]], ob[: rb[2]])
            assert bytes(a.s.prev) == bytes(b.s.prev) and bytes(a.s.seen) == bytes(b.s.seen) and a.s.run == b.s.run, (n, cap)
