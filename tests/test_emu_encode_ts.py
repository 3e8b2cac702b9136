"""CPU: the thread-serial ENCODE kernel (qoipp_b200/csrc/encode_ts.cuh, the one-shot fast path) stepped by the SIMT
emulator and compared byte for byte with the oracle.  A tile is 128 threads x 32 pixels = 4096 pixels."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle
from qoipp_b200 import synth
from tests import emu_lib as E
from tests import helpers as H

FX = H.fixtures()
TS = 32  # K == 32 selects encode_ts_kernel in tests/emu/emu_main.cpp
SIZES = [(1, 1), (1, 2), (1, 31), (1, 32), (1, 33), (1, 61), (1, 62), (1, 63), (1, 123), (1, 124), (1, 125), (29, 17), (24, 14),
         (2047, 1), (2048, 2), (4095, 1), (4097, 1), (64, 65)]


def aligned(raw):
    buf = np.empty(raw.size + 64, dtype=np.uint8)
    off = (-buf.ctypes.data) % 64
    a = buf[off: off + raw.size]
    a[:] = raw
    return a


def check(raw, w, h, ch, cs=0, **kw):
    ref = Oracle.encode(raw, w, h, ch, cs)
    out, n, ok = E.encode(aligned(raw), w, h, ch, cs, K=TS, **kw)
    assert ok and n == ref.size, (w, h, ch, n, ref.size)
    assert np.array_equal(out[:n], ref), (w, h, ch, int(np.nonzero(out[:n] != ref)[0][0]))
    assert np.all(out[n:] == 0xAA), "bytes beyond `written` must stay untouched"


@pytest.mark.parametrize("ch", [3, 4])
def test_fixtures(ch):  # simple_test.cpp:77-108
    f = FX[ch]
    w, h, _, _ = f["desc"]
    out, n, ok = E.encode(aligned(f["raw"]), w, h, ch, K=TS)
    assert ok and np.array_equal(out[:n], f["qoi"])


@pytest.mark.parametrize("kind", synth.CLASSES)
def test_classes_and_sizes(kind):
    for ch in (3, 4):
        for i, (w, h) in enumerate(SIZES):
            check(synth.generate(kind, w, h, ch), w, h, ch, cs=i & 1, seed=i)


@pytest.mark.parametrize("kind", ["photo", "long_runs", "hash_collide", "flat0", "palette", "alpha_toggle", "dither"])
def test_multi_tile_interleavings(kind):
    for ch in (3, 4):
        raw = synth.generate(kind, 160, 90, ch)  # 14400 pixels: 3 full tiles + a partial one
        for seed, resident in ((0, 1), (1, 2), (3, 4)):
            check(raw, 160, 90, ch, seed=seed, resident=resident)


def test_long_runs_across_threads_and_tiles():
    # runs that start / end / split (every 62) at thread (32 px) and tile (4096 px) boundaries
    for n in (130, 4096, 4097, 62 * 66, 62 * 66 + 1, 8192 + 31):
        for ch in (3, 4):
            raw = np.tile(np.array([0, 0, 0, 255][:ch], dtype=np.uint8), n)  # start pixel: leading run
            check(raw, n, 1, ch)
            raw2 = raw.copy()
            raw2[-ch] = 9  # run ends at the last pixel
            check(raw2, n, 1, ch)
            raw3 = np.tile(np.array([7, 7, 7, 255][:ch], dtype=np.uint8), n)
            for cut in (31, 32, 33, 4095, 4096):
                if cut < n:
                    r4 = raw3.copy()
                    r4[cut * ch] = 200
                    check(r4, n, 1, ch)


def test_random_small_palettes():
    rng = np.random.default_rng(11)
    for it in range(60):
        ch = 3 + (it & 1)
        n = int(rng.integers(1, 9000))
        pal = rng.integers(0, 256, size=(int(rng.integers(1, 6)), 4), dtype=np.uint8)
        pal[0] = [0, 0, 0, 255]
        if it % 3 == 0:
            pal[-1] = [0, 0, 0, 0]
        if it % 5 == 0 and pal.shape[0] > 2:
            pal[1] = pal[2] + np.array([64, 0, 0, 0], dtype=np.uint8)  # same slot, different colour
        idx = rng.integers(0, pal.shape[0], size=n)
        idx = np.repeat(idx, rng.integers(1, 4, size=n))[:n]
        raw = np.ascontiguousarray(pal[idx][:, :ch]).reshape(-1)
        check(raw, n, 1, ch, seed=it)


def test_slots_written_far_back():
    # a colour whose slot is not touched for many threads / tiles must still hit the index (merge + look-back)
    rng = np.random.default_rng(2)
    for ch in (3, 4):
        n = 3 * 4096 + 77
        base = rng.integers(0, 256, size=(n, 4), dtype=np.uint8)
        base[:, 0] = (base[:, 0] & 0xC0)  # few slots in use
        base[:, 1] = 0
        base[:, 2] = 0
        base[:, 3] = 255
        special = np.array([13, 200, 99, 255], dtype=np.uint8)
        for pos in (0, 5, 40, 700, 4095, 4096, 9000, n - 1):
            base[pos] = special
        check(np.ascontiguousarray(base[:, :ch]).reshape(-1), n, 1, ch)


def test_batch_images_share_one_launch():
    w, h, ch = 80, 64, 4  # 5120 pixels, raw stride 20480 (16-byte multiple)
    imgs = [synth.generate("photo", w, h, ch, seed=0x51F0 + k) for k in range(3)]
    res = E.encode(aligned(np.concatenate(imgs)), w, h, ch, K=TS, n_images=3, seed=9)
    for raw, (out, n, ok) in zip(imgs, res):
        ref = Oracle.encode(raw, w, h, ch)
        assert ok and n == ref.size and np.array_equal(out[:n], ref)


def test_more_than_one_tile_group():
    # tile offsets are sums of 64-tile group totals plus the earlier tiles of the group: 2.x groups per image here
    for ch, kind in ((3, "photo"), (4, "dither")):
        w, h = 410, 333  # 136530 pixels: 134 tiles
        check(synth.generate(kind, w, h, ch), w, h, ch, seed=3, resident=3)
    w, h, ch = 272, 256, 4  # 69632 pixels: exactly 68 tiles, raw stride 16-byte aligned
    imgs = [synth.generate("photo", w, h, ch, seed=7 + k) for k in range(2)]
    res = E.encode(aligned(np.concatenate(imgs)), w, h, ch, K=TS, n_images=2, seed=4)
    for raw, (out, n, ok) in zip(imgs, res):
        ref = Oracle.encode(raw, w, h, ch)
        assert ok and n == ref.size and np.array_equal(out[:n], ref)
