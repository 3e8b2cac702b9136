"""Development aid: randomized soak of the resumable decode (parallel tile kernel) against the oracle, state by state."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle.pyoracle import Oracle
from qoipp_b200 import api, synth
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = api.Context(0)
t0 = time.time(); it = calls = 0
cache = {}
while time.time() - t0 < budget:
    rng = np.random.default_rng(seed0 + it)
    kind = synth.CLASSES[int(rng.integers(0, len(synth.CLASSES)))]
    ch = int(rng.integers(3, 5)); w, h = int(rng.integers(100, 900)), int(rng.integers(50, 500))
    key = (kind, ch, w, h)
    raw = synth.generate(kind, w, h, ch); q = Oracle.encode(raw, w, h, ch)
    tgt = int(rng.choice([0, 3, 4]))
    a, b = api.StreamDecoder(ctx), Oracle.StreamDecoder()
    assert a.initialize(q[:14], tgt)[0] == 0 and b.initialize(q[:14], tgt)[0] == 0
    off, end = 14, q.size - 8
    while off < end:
        cap = int(rng.integers(3000, 400000)) if rng.integers(0, 4) else int(rng.integers(4, 200))
        take = int(rng.integers(2000, 120000))
        oa, ob = np.full(cap + 32, 0xAA, np.uint8), np.full(cap + 32, 0xAA, np.uint8)
        chunk = q[off: min(off + take, end)]
        sa = (a.s.run, bytes(a.s.prev))
        ra, rb = a.decode(oa[:cap], chunk), b.decode(ob[:cap], chunk)
        calls += 1
        ok = ra == rb and np.array_equal(oa[: ra[2]], ob[: rb[2]]) and (oa[cap:] == 0xAA).all() and a.s.run == b.s.run and bytes(a.s.prev) == bytes(b.s.prev) and bytes(a.s.seen) == bytes(b.s.seen)
        if not ok:
            nd = int((oa[: min(ra[2], rb[2])] != ob[: min(ra[2], rb[2])]).sum())
            seen_bad = [i for i in range(64) if bytes(a.s.seen)[4*i:4*i+4] != bytes(b.s.seen)[4*i:4*i+4]]
            print(f"MISMATCH it={it} seed={seed0 + it} {kind} {w}x{h}x{ch} tgt={tgt} off={off} cap={cap} take={chunk.size} run_in={sa[0]} got={ra} want={rb} differing bytes={nd} "
                  f"run {a.s.run}/{b.s.run} prev {bytes(a.s.prev).hex()}/{bytes(b.s.prev).hex()} seen_bad={seen_bad[:10]}", flush=True)
            sys.exit(1)
        off += ra[1]
    it += 1
print(f"soak ok: {it} streams, {calls} calls in {time.time() - t0:.0f} s")
