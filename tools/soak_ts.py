"""Development aid: randomized soak of the device-pointer paths (thread-serial encode kernels, decode) against the oracle."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle.pyoracle import Oracle
from qoipp_b200 import api, synth

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
t0 = time.time(); n = 0; npx = 0
while time.time() - t0 < budget:
    kind = synth.CLASSES[int(rng.integers(len(synth.CLASSES)))]
    ch = 3 + int(rng.integers(2))
    if rng.random() < 0.3:
        w, h = int(rng.integers(1, 5000)), int(rng.integers(1, 8))
    else:
        w, h = int(rng.integers(1, 700)), int(rng.integers(1, 500))
    raw = synth.generate(kind, w, h, ch, seed=int(rng.integers(1 << 30)))
    ref = Oracle.encode(raw, w, h, ch)
    nb = int(rng.integers(1, 4)) if w * h * ch % 16 == 0 else 1   # batches need 16-byte aligned strides for the fast path
    cap = (ch + 1) * w * h + 22
    stride = (cap + 63) // 64 * 64
    d_raw = torch.from_numpy(np.tile(raw, nb)).cuda()
    d_q = torch.full((stride * nb + 64,), 0xAA, dtype=torch.uint8, device="cuda")
    if nb == 1:
        ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st)
        wr, ok = ctx.encode_status(st)
        sizes = [wr]
        assert ok
    else:
        d_w = torch.zeros(nb, dtype=torch.int64, device="cuda")
        ctx.encode_batch_dev(d_raw, raw.size, nb, w, h, ch, 0, d_q, stride, stride, d_w, st)
        torch.cuda.synchronize()
        sizes = [int(x) for x in d_w.cpu().numpy()]
    q = d_q.cpu().numpy()
    for k in range(nb):
        got = q[k * stride: k * stride + sizes[k]]
        assert sizes[k] == ref.size and np.array_equal(got, ref), (kind, w, h, ch, nb, k, sizes[k], ref.size)
        assert np.all(q[k * stride + sizes[k]: (k + 1) * stride] == 0xAA), "bytes beyond written were touched"
    tgt = [0, 3, 4][int(rng.integers(3))]
    d_out = torch.full((w * h * (tgt or ch) + 32,), 0xAA, dtype=torch.uint8, device="cuda")
    ctx.decode_dev(torch.from_numpy(ref).cuda(), ref.size, w, h, ch, 0, tgt, False, d_out, w * h * (tgt or ch), st)
    ctx.decode_status(st)
    out = d_out.cpu().numpy()
    assert np.array_equal(out[: w * h * (tgt or ch)], Oracle.decode(ref, tgt or ch, False)), (kind, w, h, ch, tgt)
    assert np.all(out[w * h * (tgt or ch):] == 0xAA)
    n += 1; npx += w * h * nb
print(f"soak ok: {n} cases, {npx / 1e6:.1f} Mpixel encoded, {time.time() - t0:.0f} s")
