"""Development aid: decode path (retry rounds / sequential part) and time of alpha-changing content classes."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth
from oracle.pyoracle import Oracle
ctx = api.Context(0)
st = torch.cuda.current_stream().cuda_stream
for kind, w, h in [("hash_collide", 1920, 1080), ("wrap", 1920, 1080), ("alpha_toggle", 1920, 1080), ("photo", 1920, 1080), ("photo", 3840, 2160), ("palette", 1920, 1080)]:
    ch = 4
    raw = synth.generate(kind, w, h, ch)
    q = Oracle.encode(raw, w, h, ch)
    d_q = torch.from_numpy(q).cuda()
    d_out = torch.zeros(raw.size, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.decode_dev(d_q, q.size, w, h, ch, 0, 0, False, d_out, raw.size, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    path = ctx.decode_status(st)
    print(f"{kind:13s} {w}x{h}: path {path} ok {bool(np.array_equal(d_out.cpu().numpy(), raw))} median {np.median(ts[1:]):.1f} us")
