#!/usr/bin/env python
"""Regenerates tests/golden/*.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

1. fixtures.npz   -- the reference's own golden vectors, parsed from the comma-separated hex of
                     /root/reference/test/resources/image_{raw,qoi}_{3,4}.txt and the two *_incomplete prefixes
                     (used by test/source/simple_test.cpp:36-70 and stream_test.cpp:131-183).
2. ref_vectors.npz -- outputs of the UNMODIFIED reference (oracle/_ref/libqoipp_ref.so) on small synthetic inputs
                     and on hand-made adversarial streams, so that the pins travel to hosts without the reference.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pyoracle import Ref  # noqa: E402
from qoipp_b200 import synth  # noqa: E402

RES = "/root/reference/test/resources"
OUT = os.path.dirname(os.path.abspath(__file__))


def parse_hex(path):
    return np.array([int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", open(path).read())], dtype=np.uint8)


def header(w, h, ch, cs=0):
    return b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([ch, cs])


def adversarial_streams():
    """(name, qoi bytes, target) -- decoder corner cases listed in SURVEY.md section 8 'pinned by probe'."""
    end = bytes([0, 0, 0, 0, 0, 0, 0, 1])
    rng = np.random.default_rng(0xC0FFEE)
    cases = [
        ("index53_index5", header(2, 1, 4) + bytes([53, 5]) + end, 0),
        ("run62_in_3px", header(3, 1, 3) + bytes([0xC0 | 61]) + end, 0),
        ("too_few_ops", header(8, 4, 4) + bytes([0xFE, 10, 20, 30, 0x55]) + end, 0),
        ("too_few_ops_rgb", header(8, 4, 3) + bytes([0xFE, 10, 20, 30, 0xC3]) + end, 4),
        ("rgba_op_in_rgb_file", header(4, 1, 3) + bytes([0xFF, 1, 2, 3, 77, 0x6A, 0xFE, 9, 9, 9, 0xC0]) + end, 4),
        ("index_unwritten_then_rgb", header(6, 1, 4) + bytes([7, 0xFE, 1, 2, 3, 0, 7, 0xA0, 0x88, 0xC1]) + end, 0),
        ("luma_wrap", header(4, 1, 3) + bytes([0xFE, 250, 3, 128, 0xBF, 0xFF, 0x80, 0x00, 0x7F]) + end, 0),
    ]
    for k in range(24):
        body = rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8).tobytes()
        w, h = 64, len(body) + 2  # >= 62 pixels per body byte: the reference never runs past the image
        ch = 3 + (k & 1)
        cases.append((f"garbage_{k}", header(w, h, ch) + body + end, [0, 3, 4][k % 3]))
    for k in range(8):  # INDEX / small-op soup (chains through the table, unwritten slots)
        body = rng.choice(np.array([0, 1, 5, 53, 0x6A, 0x55, 0x7F, 0xA0, 0x11, 0xC1, 38, 17], dtype=np.uint8), size=60).tobytes()
        cases.append((f"soup_{k}", header(64, 64, 4) + body + end, 0))
    return cases


def main():
    assert Ref.available(), "oracle/_ref/libqoipp_ref.so is missing (make -C oracle)"
    fx = {}
    for n in ("image_raw_3", "image_qoi_3", "image_qoi_3_incomplete", "image_raw_4", "image_qoi_4", "image_qoi_4_incomplete"):
        fx[n] = parse_hex(os.path.join(RES, n + ".txt"))
    np.savez_compressed(os.path.join(OUT, "fixtures.npz"), **fx)

    vec = {}
    W, H = 37, 23
    for kind in synth.CLASSES:
        for ch in (3, 4):
            raw = synth.generate(kind, W, H, ch)
            for cs in (0, 1):
                enc = Ref.encode(raw, W, H, ch, cs)
                vec[f"enc/{kind}/{ch}/{cs}"] = enc
            # partial encodes at three capacities (chunk-boundary rule, simple.cpp:84-88)
            for cap in (13, 14, 15, len(enc) // 2, len(enc) - 8, len(enc) - 1):
                e, out, written, complete = Ref.encode_into(raw, W, H, ch, 0, cap=cap)
                assert e == 0
                vec[f"partial/{kind}/{ch}/{cap}"] = np.concatenate([np.array([written, int(complete)], dtype=np.uint64).view(np.uint8), out[:written]])
    for name, qoi, target in adversarial_streams():
        q = np.frombuffer(qoi, dtype=np.uint8)
        vec[f"adv/{name}/in"] = q
        vec[f"adv/{name}/target"] = np.array([target], dtype=np.uint8)
        vec[f"adv/{name}/out"] = Ref.decode(q, target)
    np.savez_compressed(os.path.join(OUT, "ref_vectors.npz"), **vec)
    print("fixtures:", {k: v.size for k, v in fx.items()})
    print("ref_vectors:", len(vec), "arrays")


if __name__ == "__main__":
    main()
