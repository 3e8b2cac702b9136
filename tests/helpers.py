"""Shared test helpers: fixtures, the reference's stream calling protocol, case tables."""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

HEADER, MARKER = 14, 8

# test/source/simple_test.cpp:29-34,53-58 / stream_test.cpp:131-183
DESC3 = (29, 17, 3, 0)
DESC4 = (24, 14, 4, 0)
CHUNK_BOUNDARY = 1007  # simple_test.cpp:24-25


def fixtures():
    z = np.load(os.path.join(GOLDEN, "fixtures.npz"))
    return {
        3: dict(desc=DESC3, raw=z["image_raw_3"], qoi=z["image_qoi_3"], qoi_incomplete=z["image_qoi_3_incomplete"]),
        4: dict(desc=DESC4, raw=z["image_raw_4"], qoi=z["image_qoi_4"], qoi_incomplete=z["image_qoi_4_incomplete"]),
    }


def ref_vectors():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


def to_rgb(raw4):  # test/source/util.hpp:61-75
    return np.ascontiguousarray(raw4.reshape(-1, 4)[:, :3]).reshape(-1)


def to_rgba(raw3):  # test/source/util.hpp:77-92
    px = raw3.reshape(-1, 3)
    return np.concatenate([px, np.full((px.shape[0], 1), 255, dtype=np.uint8)], axis=1).reshape(-1)


def retarget(raw, ch, target):
    if target in (0, ch):
        return raw
    return to_rgb(raw) if target == 3 else to_rgba(raw)


def stream_encode(enc, desc, bufsize, raw):
    """The canonical calling protocol of test/source/stream_test.cpp:43-80."""
    w, h, ch, cs = desc
    head = np.zeros(HEADER, dtype=np.uint8)
    e, n = enc.initialize(head, w, h, ch, cs)
    assert e == 0 and n == HEADER
    parts = [head]
    out = np.zeros(bufsize, dtype=np.uint8)
    off = 0
    guard = 0
    while off < raw.size:
        chunk = raw[off: off + min(bufsize, raw.size - off)]
        e, processed, written = enc.encode(out, chunk)
        assert e == 0, e
        off += processed
        parts.append(out[:written].copy())
        guard += 1
        assert guard < 10 * raw.size + 100, "stream encoder makes no progress"
    extra = MARKER + int(enc.has_run_count())
    tail = np.zeros(extra, dtype=np.uint8)
    e, n = enc.finalize(tail)
    assert e == 0 and n == extra
    parts.append(tail)
    return np.concatenate(parts)


def stream_decode(dec, bufsize, qoi, target=0):
    """stream_test.cpp:82-123: the last 8 bytes are withheld as the end marker."""
    e, desc = dec.initialize(qoi[:HEADER], target)
    assert e == 0, e
    out = np.zeros(bufsize, dtype=np.uint8)
    parts = []
    off, end = HEADER, qoi.size - MARKER
    guard = 0
    while off < end:
        chunk = qoi[off: off + min(bufsize, end - off)]
        e, processed, written = dec.decode(out, chunk)
        assert e == 0, e
        off += processed
        parts.append(out[:written].copy())
        guard += 1
        assert guard < 10 * qoi.size + 100, "stream decoder makes no progress"
    while dec.has_run_count():
        e, n = dec.drain_run(out)
        assert e == 0
        parts.append(out[:n].copy())
    dec.reset()
    return (np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)), desc
