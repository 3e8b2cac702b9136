"""CPU: the product ENCODE kernel (qoipp_b200/csrc/encode_kernel.cuh) stepped by the SIMT emulator and compared
byte for byte with the oracle.  Small tiles (K=1: 256 pixels) force long look-back chains on small images."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle
from qoipp_b200 import synth
from tests import emu_lib as E
from tests import helpers as H

FX = H.fixtures()
SIZES = [(1, 1), (1, 2), (1, 61), (1, 62), (1, 63), (1, 123), (1, 124), (1, 125), (29, 17), (24, 14), (255, 1), (256, 1),
         (257, 1), (33, 31), (128, 20)]


def check(raw, w, h, ch, cs=0, cap=None, **kw):
    e, ref, rn, rok = Oracle.encode_into(raw, w, h, ch, cs, cap=cap)
    assert e == 0
    out, n, ok = E.encode(raw, w, h, ch, cs, cap=cap, **kw)
    assert (n, ok) == (rn, rok), (w, h, ch, cap, n, rn, ok, rok)
    assert np.array_equal(out[:n], ref[:n]), (w, h, ch, cap, int(np.nonzero(out[:n] != ref[:n])[0][0]))
    assert np.all(out[n:] == 0xAA), "bytes beyond `written` must stay untouched"


@pytest.mark.parametrize("ch", [3, 4])
def test_fixtures(ch):  # simple_test.cpp:77-108
    f = FX[ch]
    w, h, _, _ = f["desc"]
    for K in (1, 2, 8):
        out, n, ok = E.encode(f["raw"], w, h, ch, K=K)
        assert ok and np.array_equal(out[:n], f["qoi"])
    out, n, ok = E.encode(f["raw"], w, h, ch, cap=H.CHUNK_BOUNDARY)
    assert not ok and n == H.CHUNK_BOUNDARY and np.array_equal(out[:n], f["qoi"][:n])


@pytest.mark.parametrize("kind", synth.CLASSES)
def test_classes_and_sizes(kind):
    for ch in (3, 4):
        for i, (w, h) in enumerate(SIZES):
            raw = synth.generate(kind, w, h, ch)
            check(raw, w, h, ch, cs=i & 1, K=1 + (i & 1), seed=i)


@pytest.mark.parametrize("kind", ["photo", "long_runs", "hash_collide", "flat0", "palette", "alpha_toggle"])
def test_multi_tile_interleavings(kind):
    for ch in (3, 4):
        raw = synth.generate(kind, 96, 64, ch)
        for seed, resident in ((0, 1), (1, 2), (2, 5), (3, 8)):
            check(raw, 96, 64, ch, K=1, seed=seed, resident=resident)
    raw = synth.generate(kind, 160, 90, 4)
    check(raw, 160, 90, 4, K=8, seed=5)


def test_partial_capacity_sweep():
    rng = np.random.default_rng(3)
    for it, kind in enumerate(synth.CLASSES):
        ch = 3 + (it & 1)
        w, h = 40, 23
        raw = synth.generate(kind, w, h, ch)
        full = Oracle.encode(raw, w, h, ch)
        caps = {0, 13, 14, 15, 18, full.size - 9, full.size - 8, full.size - 1, full.size, full.size + 3}
        caps |= set(int(c) for c in rng.integers(14, full.size, size=12))
        for cap in sorted(c for c in caps if c >= 0):
            check(raw, w, h, ch, cap=cap, K=1, seed=it)


def test_long_run_across_many_tiles():
    # 130 start pixels -> fd fd c5; runs that span tiles and end exactly at tile / 62 boundaries
    for n in (130, 255, 256, 257, 62 * 5, 62 * 5 + 1, 1024, 1025):
        for ch in (3, 4):
            raw = np.tile(np.array([0, 0, 0, 255][:ch], dtype=np.uint8), n)
            check(raw, n, 1, ch, K=1)
            raw2 = raw.copy()
            raw2[-ch] = 9  # run ends at the last pixel
            check(raw2, n, 1, ch, K=1)
            raw3 = np.tile(np.array([7, 7, 7, 255][:ch], dtype=np.uint8), n)
            check(raw3, n, 1, ch, K=1)


def test_random_small_palettes():
    rng = np.random.default_rng(11)
    for it in range(120):
        ch = 3 + (it & 1)
        n = int(rng.integers(1, 700))
        pal = rng.integers(0, 256, size=(int(rng.integers(1, 6)), 4), dtype=np.uint8)
        pal[0] = [0, 0, 0, 255]
        if it % 3 == 0:
            pal[-1] = [0, 0, 0, 0]
        if it % 5 == 0 and pal.shape[0] > 2:
            pal[1] = pal[2] + np.array([64, 0, 0, 0], dtype=np.uint8)  # same slot, different colour
        idx = rng.integers(0, pal.shape[0], size=n)
        idx = np.repeat(idx, rng.integers(1, 4, size=n))[:n]
        raw = np.ascontiguousarray(pal[idx][:, :ch]).reshape(-1)
        check(raw, n, 1, ch, K=1, seed=it)


def test_batch_images_share_one_launch():
    w, h, ch = 31, 19, 4
    imgs = [synth.generate("photo", w, h, ch, seed=0x51F0 + k) for k in range(5)]
    res = E.encode(np.concatenate(imgs), w, h, ch, K=1, n_images=5, seed=9)
    for raw, (out, n, ok) in zip(imgs, res):
        ref = Oracle.encode(raw, w, h, ch)
        assert ok and n == ref.size and np.array_equal(out[:n], ref)


@pytest.mark.parametrize("ch", [3, 4])
def test_stream_encoder_sweep(ch):  # stream_test.cpp:192-201 (every 7th size here; the GPU test runs them all)
    f = FX[ch]
    enc = E.StreamEncoder(K=1)
    for size in list(range(5, 64)) + list(range(64, 1025, 7)):
        got = H.stream_encode(enc, f["desc"], size, f["raw"])
        assert np.array_equal(got, f["qoi"]), size


def test_stream_encoder_matches_oracle_state_by_state():
    rng = np.random.default_rng(5)
    for it in range(40):
        kind = synth.CLASSES[it % len(synth.CLASSES)]
        ch = 3 + (it & 1)
        w, h = int(rng.integers(1, 50)), int(rng.integers(1, 30))
        raw = synth.generate(kind, w, h, ch, seed=77 + it)
        a, b = E.StreamEncoder(K=1, seed=it), Oracle.StreamEncoder()
        hd = np.zeros(14, np.uint8)
        a.initialize(hd, w, h, ch)
        b.initialize(hd.copy(), w, h, ch)
        off = 0
        while off < raw.size:
            cap = int(rng.integers(5, 120))
            take = int(rng.integers(1, 400))
            oa, ob = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
            ra = a.encode(oa, raw[off: off + take])
            rb = b.encode(ob, raw[off: off + take])
            assert ra == rb, (kind, ch, off, cap, take, ra, rb)
            assert np.array_equal(oa[: ra[2]], ob[: rb[2]])
            assert a.s.run == b.s.run and bytes(a.s.prev) == bytes(b.s.prev), (kind, off)
            assert bytes(a.s.seen) == bytes(b.s.seen), (kind, ch, off, cap, take)
            off += ra[1]
            if ra[1] == 0 and ra[2] == 0:
                break
