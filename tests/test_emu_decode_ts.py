"""CPU: the thread-serial DECODE fast path (qoipp_b200/csrc/decode_ts.cuh) as round 0, followed by decode_finish_kernel,
stepped by the SIMT emulator and compared with the oracle.  A fast tile is one warp x 68 stream bytes per lane = 2176 bytes.
path >= 1000 means the fast path flagged the image and the general machinery decoded it."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle
from qoipp_b200 import synth
from tests import emu_lib as E
from tests import helpers as H

FX = H.fixtures()
FAST = 2  # force_serial == 2 selects decode_ts_kernel as round 0 in tests/emu/emu_main.cpp
SIZES = [(1, 1), (1, 2), (1, 61), (1, 62), (1, 63), (1, 124), (29, 17), (24, 14), (255, 3), (100, 41)]


def check(qoi, w, h, src_ch, target=0, expect_fast=None, **kw):
    tgt = target or src_ch
    ref = Oracle.decode(qoi, tgt, False)
    px, path = E.decode(qoi, w, h, tgt, force_serial=FAST, **kw)
    if not np.array_equal(px[0], ref):
        bad = int(np.nonzero(px[0] != ref)[0][0]) // tgt
        raise AssertionError(f"{w}x{h} src {src_ch} -> {tgt}: first wrong pixel {bad} (path {path})")
    if expect_fast is True:
        assert path[0] == 0, path
    elif expect_fast is False:
        assert path[0] >= 1000, path
    return path[0]


@pytest.mark.parametrize("ch", [3, 4])
def test_fixtures(ch):  # simple_test.cpp:179-223, 316-322
    f = FX[ch]
    w, h, _, _ = f["desc"]
    for target in (0, 3, 4):
        px, path = E.decode(f["qoi"], w, h, target or ch, force_serial=FAST)
        assert np.array_equal(px[0], H.retarget(f["raw"], ch, target))
    check(f["qoi_incomplete"], w, h, ch, expect_fast=False)  # truncated stream: the general path decodes the zero padding


@pytest.mark.parametrize("kind", synth.CLASSES)
def test_classes_and_sizes(kind):
    for ch in (3, 4):
        for i, (w, h) in enumerate(SIZES):
            raw = synth.generate(kind, w, h, ch)
            q = Oracle.encode(raw, w, h, ch)
            check(q, w, h, ch, target=[0, 3, 4][i % 3], seed=i)


@pytest.mark.parametrize("kind", ["photo", "dither", "noise", "resync", "gradient", "long_runs", "wrap"])
def test_opaque_content_is_verified_by_the_fast_path(kind):
    for ch in (3, 4):
        w, h = 160, 100
        raw = synth.generate(kind, w, h, ch)
        if ch == 4:
            raw = raw.copy()
            raw[3::4] = 255
        q = Oracle.encode(raw, w, h, ch)
        for seed, resident in ((0, 1), (2, 3), (5, 8)):
            check(q, w, h, ch, seed=seed, resident=resident, expect_fast=True)


@pytest.mark.parametrize("kind", ["alpha_toggle", "hash_collide", "palette", "photo"])
def test_transparent_content_is_exact_either_way(kind):
    w, h = 120, 90
    raw = synth.generate(kind, w, h, 4)
    q = Oracle.encode(raw, w, h, 4)
    for target in (0, 3):
        check(q, w, h, 4, target=target, seed=3)


def test_index_heavy_streams():
    # few colours: nearly every op is an OP_INDEX whose writer lies lanes or tiles back
    rng = np.random.default_rng(4)
    for ch in (3, 4):
        for ncol in (2, 5, 17, 40):
            pal = rng.integers(0, 256, size=(ncol, 4), dtype=np.uint8)
            pal[:, 3] = 255
            idx = rng.integers(0, ncol, size=150 * 80)
            raw = np.ascontiguousarray(pal[idx][:, :ch]).reshape(-1)
            check(Oracle.encode(raw, 150, 80, ch), 150, 80, ch, seed=ncol)


def test_batch():
    w, h, ch = 96, 70, 3
    imgs = [synth.generate("photo", w, h, ch, seed=0x51F0 + k) for k in range(4)]
    qs = [Oracle.encode(r, w, h, ch) for r in imgs]
    px, path = E.decode(qs, w, h, ch, force_serial=FAST, seed=2, resident=3)
    for r, p in zip(imgs, px):
        assert np.array_equal(p, r)
    assert all(x == 0 for x in path), path


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
            px, path = E.decode(q, w, h, target or ch, force_serial=FAST)
            assert np.array_equal(px[0], v[f"adv/{name}/out"]), (name, path)
            n += 1
    assert n >= 39


def test_truncated_and_padded_streams():
    for kind in ("photo", "palette", "long_runs"):
        for ch in (3, 4):
            w, h = 64, 37
            raw = synth.generate(kind, w, h, ch)
            q = Oracle.encode(raw, w, h, ch)
            for cut in (q.size - 8, q.size - 9, q.size - 11, q.size // 2, 40, 23):
                if cut > 22:
                    check(q[:cut], w, h, ch)


def test_random_op_soup():
    rng = np.random.default_rng(99)
    tags = np.array([0, 1, 5, 53, 0x6A, 0x55, 0x7F, 0xA0, 0x88, 0x11, 0xC1, 0xC5, 38, 17, 0xFE, 0xFF, 0x80, 0x3F], dtype=np.uint8)
    for it in range(60):
        nb = int(rng.integers(1, 9000))
        body = rng.choice(tags, size=nb) if it % 2 else rng.integers(0, 256, size=nb, dtype=np.uint8)
        ch = 3 + (it & 1)
        w, h = 97, int(rng.integers(1, 80))
        hdr = np.frombuffer(b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([ch, 0]), dtype=np.uint8)
        q = np.concatenate([hdr, body.astype(np.uint8), np.array([0, 0, 0, 0, 0, 0, 0, 1], np.uint8)])
        check(q, w, h, ch, target=[0, 3, 4][it % 3], seed=it)


def test_more_ops_than_pixels_and_late_resync():
    # a parse that never self-synchronises (FE FE FE ..) and surplus ops behind the last pixel
    for ch in (3, 4):
        w, h = 70, 60
        raw = synth.generate("resync", w, h, ch)
        q = Oracle.encode(raw, w, h, ch)
        extra = np.concatenate([q[:-8], np.array([0x6A, 0xFE, 1, 2, 3, 0xC5, 0x22] * 40, np.uint8), q[-8:]])
        check(extra, w, h, ch, seed=1)
