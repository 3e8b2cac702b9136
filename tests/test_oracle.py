"""CPU: pins oracle/qoi_oracle.c against the reference's golden fixtures, the committed outputs of the
unmodified reference (tests/golden/ref_vectors.npz) and, where it is built, oracle/_ref itself."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle, Ref
from qoipp_b200 import synth
from tests import helpers as H

FX = H.fixtures()
HAVE_REF = Ref.available()
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built on this host")


@pytest.mark.parametrize("ch", [3, 4])
def test_fixture_encode(ch):  # simple_test.cpp:77-83
    f = FX[ch]
    assert np.array_equal(Oracle.encode(f["raw"], *f["desc"]), f["qoi"])


@pytest.mark.parametrize("ch", [3, 4])
def test_fixture_partial_encode(ch):  # simple_test.cpp:85-108
    f = FX[ch]
    e, out, written, complete = Oracle.encode_into(f["raw"], *f["desc"], cap=H.CHUNK_BOUNDARY)
    assert e == 0 and not complete and written == H.CHUNK_BOUNDARY
    assert np.array_equal(out[:written], f["qoi"][: H.CHUNK_BOUNDARY])


@pytest.mark.parametrize("ch", [3, 4])
@pytest.mark.parametrize("target", [0, 3, 4])
def test_fixture_decode(ch, target):  # simple_test.cpp:179-223
    f = FX[ch]
    e, px, desc = Oracle.decode_into(f["qoi"], target)
    assert e == 0
    assert desc == (f["desc"][0], f["desc"][1], target or ch, 0)
    assert np.array_equal(px, H.retarget(f["raw"], ch, target))


@pytest.mark.parametrize("ch", [3, 4])
def test_fixture_header_and_errors(ch):  # simple_test.cpp:282-295
    f = FX[ch]
    assert Oracle.read_header(f["qoi"]) == (0, f["desc"])
    assert Oracle.read_header(np.zeros(0, np.uint8))[0] == 1
    assert Oracle.read_header(np.array([1, 2, 3, 4], np.uint8))[0] == 2
    assert Oracle.decode_into(np.zeros(0, np.uint8))[0] == 1
    assert Oracle.decode_into(f["qoi"][:22])[0] == 2
    bad = f["qoi"].copy()
    bad[0] = 0
    assert Oracle.decode_into(bad)[0] == 4


@pytest.mark.parametrize("ch", [3, 4])
def test_fixture_stream_sweep(ch):  # stream_test.cpp:192-252, every buffer size 5..1024
    f = FX[ch]
    enc, dec = Oracle.StreamEncoder(), Oracle.StreamDecoder()
    for size in range(5, 1025):
        got = H.stream_encode(enc, f["desc"], size, f["raw"])
        assert np.array_equal(got, f["qoi"]), size
        for target in (0, 3, 4):
            px, desc = H.stream_decode(dec, size, f["qoi"], target)
            assert np.array_equal(px, H.retarget(f["raw"], ch, target)), (size, target)
            assert desc[2] == (target or ch)
        px, _ = H.stream_decode(dec, size, f["qoi_incomplete"])
        assert px.size != f["raw"].size and np.array_equal(px, f["raw"][: px.size]), size


def test_committed_reference_vectors():
    v = H.ref_vectors()
    names = list(v.keys())
    W, Hh = 37, 23
    n_enc = n_part = n_adv = 0
    for k in names:
        parts = k.split("/")
        if parts[0] == "enc":
            kind, ch, cs = parts[1], int(parts[2]), int(parts[3])
            raw = synth.generate(kind, W, Hh, ch)
            enc = Oracle.encode(raw, W, Hh, ch, cs)
            assert np.array_equal(enc, v[k]), k
            for target in (0, 3, 4):
                assert np.array_equal(Oracle.decode(enc, target), H.retarget(raw, ch, target)), k
            n_enc += 1
        elif parts[0] == "partial":
            kind, ch, cap = parts[1], int(parts[2]), int(parts[3])
            raw = synth.generate(kind, W, Hh, ch)
            blob = v[k]
            written, complete = (int(x) for x in blob[:16].view(np.uint64))
            e, out, w2, c2 = Oracle.encode_into(raw, W, Hh, ch, 0, cap=cap)
            assert (e, w2, c2) == (0, written, bool(complete)), k
            assert np.array_equal(out[:w2], blob[16:]), k
            n_part += 1
        elif parts[0] == "adv" and parts[2] == "in":
            name = parts[1]
            target = int(v[f"adv/{name}/target"][0])
            assert np.array_equal(Oracle.decode(v[k], target), v[f"adv/{name}/out"]), name
            n_adv += 1
    assert n_enc == 48 and n_part == 144 and n_adv >= 39


SIZES = [(1, 1), (1, 61), (1, 62), (1, 63), (1, 123), (1, 124), (1, 125), (29, 17), (24, 14), (200, 120)]


@needs_ref
@pytest.mark.parametrize("kind", synth.CLASSES)
def test_oracle_equals_reference_synthetic(kind):
    for ch in (3, 4):
        for (w, h) in SIZES:
            raw = synth.generate(kind, w, h, ch)
            for cs in (0, 1):
                a, b = Oracle.encode(raw, w, h, ch, cs), Ref.encode(raw, w, h, ch, cs)
                assert np.array_equal(a, b), (kind, ch, w, h)
            for target in (0, 3, 4):
                assert np.array_equal(Oracle.decode(a, target), Ref.decode(a, target))
            for flip in (True,):
                assert np.array_equal(Oracle.decode(a, 0, flip), Ref.decode(a, 0, flip))
            # truncated streams: both keep decoding the zero padding (simple.cpp:106)
            for cut in (a.size - 8, a.size - 9, max(23, a.size // 2)):
                if cut > 22:
                    assert np.array_equal(Oracle.decode(a[:cut]), Ref.decode(a[:cut])), (kind, ch, w, h, cut)


@needs_ref
def test_oracle_equals_reference_random_partial_and_streams():
    rng = np.random.default_rng(7)
    enc_o, enc_r = Oracle.StreamEncoder(), Ref.StreamEncoder()
    dec_o, dec_r = Oracle.StreamDecoder(), Ref.StreamDecoder()
    for it in range(60):
        kind = synth.CLASSES[it % len(synth.CLASSES)]
        ch = 3 + (it & 1)
        w, h = int(rng.integers(1, 60)), int(rng.integers(1, 40))
        raw = synth.generate(kind, w, h, ch, seed=1000 + it)
        full = Ref.encode(raw, w, h, ch)
        for cap in rng.integers(0, full.size + 4, size=6):
            ro = Oracle.encode_into(raw, w, h, ch, 0, cap=int(cap))
            rr = Ref.encode_into(raw, w, h, ch, 0, cap=int(cap))
            assert (ro[0], ro[2], ro[3]) == (rr[0], rr[2], rr[3]), (kind, ch, w, h, cap)
            assert np.array_equal(ro[1][: ro[2]], rr[1][: rr[2]])
        for size in rng.integers(5, 200, size=3):
            size = int(size)
            assert np.array_equal(H.stream_encode(enc_o, (w, h, ch, 0), size, raw), H.stream_encode(enc_r, (w, h, ch, 0), size, raw))
            for target in (0, 3, 4):
                if size < (target or ch):
                    continue
                a, _ = H.stream_decode(dec_o, size, full, target)
                b, _ = H.stream_decode(dec_r, size, full, target)
                assert np.array_equal(a, b)


@needs_ref
def test_error_codes_match_reference():
    raw = synth.generate("noise", 4, 4, 4)
    for args in [(raw[:0], 4, 4, 4), (raw, 0, 4, 4), (raw, 4, 4, 5), (raw[:-1], 4, 4, 4), (raw, 4, 4, 3)]:
        a = Oracle.encode_into(*args)
        b = Ref.encode_into(*args)
        assert a[0] == b[0] and a[0] != 0, args[1:]
    q = Oracle.encode(raw, 4, 4, 4)
    for bad in (q[:0], q[:10], q[:22]):
        assert Oracle.decode_into(bad)[0] == Ref.decode_err(bad)
    nq = q.copy()
    nq[12] = 7
    assert Oracle.decode_into(nq)[0] == Ref.decode_err(nq) == 5
    # decode_into capacity rule (simple.cpp:467-471): NotEnoughSpace is sized with the source channels
    assert Ref.decode_into(q, 4 * 4 * 4 - 1)[0] == 7 and Oracle.decode_into(q, cap=4 * 4 * 4 - 1)[0] == 7
