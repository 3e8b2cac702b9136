"""GPU (-m gpu): randomized sweep in the spirit of the reference's libFuzzer harness (example/source/99_fuzz.cpp):
the first bytes of a random blob become the Desc, the rest the payload; one-shot and resumable paths are driven and
every result (bytes, counts, error values, carried state) is compared with the oracle."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from qoipp_b200 import api

    c = api.Context(0)
    yield c
    c.close()


def blob_to_case(rng):
    w = int(rng.integers(0, 90))
    h = int(rng.integers(0, 70))
    ch = int(rng.choice([3, 4, 4, 3, 5, 0]))
    cs = int(rng.choice([0, 1, 0, 1, 2]))
    style = int(rng.integers(0, 4))
    n = max(w * h * max(ch, 1), 0)
    if style == 0:
        data = rng.integers(0, 256, size=n, dtype=np.uint8)
    elif style == 1:  # few colours, long runs
        pal = rng.integers(0, 256, size=(4, 4), dtype=np.uint8)
        idx = np.repeat(rng.integers(0, 4, size=n // max(ch, 1) + 1), rng.integers(1, 90, size=n // max(ch, 1) + 1))[: n // max(ch, 1)]
        data = np.ascontiguousarray(pal[idx][:, : max(ch, 1)]).reshape(-1)[:n]
    elif style == 2:  # smooth
        base = np.cumsum(rng.integers(-3, 4, size=(n // max(ch, 1) + 1, 4)), axis=0).astype(np.uint8)
        data = np.ascontiguousarray(base[:, : max(ch, 1)]).reshape(-1)[:n]
    else:
        data = rng.integers(0, 256, size=int(rng.integers(0, n + 5)), dtype=np.uint8)  # wrong size on purpose
    return w, h, ch, cs, np.ascontiguousarray(data, dtype=np.uint8)


def test_fuzz_one_shot(ctx):
    rng = np.random.default_rng(20261018)
    errors = 0
    for it in range(250):
        w, h, ch, cs, data = blob_to_case(rng)
        cap = None if it % 3 else int(rng.integers(0, 400))
        eo = Oracle.encode_into(data, w, h, ch, cs, cap=cap if cap is not None else max((ch + 1) * w * h + 22, 1))
        eg = ctx.encode_into(data, w, h, ch, cs, cap=cap if cap is not None else max((ch + 1) * w * h + 22, 1))
        assert eg[0] == eo[0], (it, w, h, ch, cs, data.size, eg[0], eo[0])
        if eo[0]:
            errors += 1
            continue
        assert (eg[2], eg[3]) == (eo[2], eo[3]), (it, w, h, ch, cap)
        assert np.array_equal(eg[1][: eg[2]], eo[1][: eo[2]])
        if eo[3]:  # complete stream: decode it, also with random corruption of the chunk bytes
            q = eo[1][: eo[2]].copy()
            for trial in range(2):
                tgt = int(rng.choice([0, 3, 4]))
                flip = bool(rng.integers(0, 2))
                do, dg = Oracle.decode_into(q, tgt, flip), ctx.decode_into(q, tgt, flip)
                assert dg[0] == do[0], (it, trial)
                if do[0] == 0:
                    assert dg[2] == do[2] and np.array_equal(dg[1], do[1]), (it, trial, w, h, ch, tgt, flip)
                if q.size > 30:  # corrupt a few chunk bytes (the header stays valid) and decode again
                    pos = rng.integers(14, q.size - 8, size=3)
                    q[pos] = rng.integers(0, 256, size=3, dtype=np.uint8)
    assert errors > 10  # the sweep did reach the validation paths


def test_fuzz_streams(ctx):
    from qoipp_b200 import api

    rng = np.random.default_rng(77)
    for it in range(40):
        w, h, ch = int(rng.integers(1, 70)), int(rng.integers(1, 50)), int(rng.choice([3, 4]))
        style_rng = np.random.default_rng(it)
        raw = blob_to_case(style_rng)[4]
        raw = np.resize(raw if raw.size else np.zeros(1, np.uint8), w * h * ch).astype(np.uint8)
        ea, eb = api.StreamEncoder(ctx), Oracle.StreamEncoder()
        size = int(rng.integers(5, 700))
        a = H.stream_encode(ea, (w, h, ch, 0), size, raw)
        b = H.stream_encode(eb, (w, h, ch, 0), size, raw)
        assert np.array_equal(a, b), (it, w, h, ch, size)
        da, db = api.StreamDecoder(ctx), Oracle.StreamDecoder()
        tgt = int(rng.choice([0, 3, 4]))
        size = int(rng.integers(5, 700))
        pa, _ = H.stream_decode(da, size, a, tgt)
        pb, _ = H.stream_decode(db, size, a, tgt)
        assert np.array_equal(pa, pb), (it, w, h, ch, size, tgt)
