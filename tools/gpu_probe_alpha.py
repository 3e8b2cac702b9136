"""Scratch probe: decode time of RGBA images whose INDEX ops change alpha (sequential path today)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
for kind, w, h, ch in [("photo", 1920, 1080, 4), ("photo", 512, 512, 4), ("palette", 1920, 1080, 4), ("alpha_toggle", 1920, 1080, 4), ("hash_collide", 1920, 1080, 4), ("wrap", 1920, 1080, 4)]:
    raw = synth.generate(kind, w, h, ch)
    d_raw = torch.from_numpy(raw).cuda(); cap = (ch + 1) * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st); n, ok = ctx.encode_status(st)
    d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
    ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st); path = ctx.decode_status(st)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3): ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st)
    ev[1].record(); torch.cuda.synchronize()
    print(f"{kind} {w}x{h}x{ch}: decode {ev[0].elapsed_time(ev[1])/3*1e3:.0f} us path={path} ok={bool(torch.equal(d_out, d_raw))} E/raw={n/raw.size:.3f}", flush=True)
