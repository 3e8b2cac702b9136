"""Development aid: device-timed encode and decode of every synthetic content class at one size, with the
decode path that produced the pixels.  usage: python tools/class_sweep.py [w h]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
print(f"{w}x{h}, device-timed, L2 flushed, median of 5; us")
print(f"{'class':14s} ch  E/N    encode    decode  path   GB/s raw (enc / dec)")
for kind in synth.CLASSES:
    for ch in (3, 4):
        raw = synth.generate(kind, w, h, ch)
        d_raw = torch.from_numpy(raw).cuda(); cap = (ch + 1) * w * h + 22
        d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
        te, td = [], []
        for it in range(7):
            e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            flush.fill_(it); e0.record(); ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st); e1.record()
            n, ok = ctx.encode_status(st)
            flush.fill_(it + 50); e2.record(); ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st); e3.record(); torch.cuda.synchronize()
            te.append(e0.elapsed_time(e1) * 1e3); td.append(e2.elapsed_time(e3) * 1e3)
        path = ctx.decode_status(st)
        assert torch.equal(d_out, d_raw), kind
        e, d = np.median(te[2:]), np.median(td[2:])
        print(f"{kind:14s} {ch}  {n / (w * h):5.3f} {e:9.1f} {d:9.1f}  {path:4d}   {raw.size / e / 1e3:7.1f} / {raw.size / d / 1e3:7.1f}", flush=True)
