"""Scratch timing probe (development only): device-timed decode."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth
from oracle.pyoracle import Oracle

ctx = api.Context(0)
st = torch.cuda.current_stream().cuda_stream
for kind, w, h, ch in [("photo", 3840, 2160, 3), ("photo", 7680, 4320, 3), ("noise", 3840, 2160, 4), ("flat", 7680, 4320, 4), ("gradient", 7680, 4320, 3), ("palette", 3840, 2160, 3), ("dither", 3840, 2160, 3)]:
    raw = synth.generate(kind, w, h, ch)
    d_raw = torch.from_numpy(raw).cuda()
    cap = (ch + 1) * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st)
    n, ok = ctx.encode_status(st)
    d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st)
    path = ctx.decode_status(st)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 20
    ev[0].record()
    for _ in range(reps):
        ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    alg = raw.size + n
    print(f"decode {kind} {w}x{h}x{ch}: {ms*1e3:.1f} us  raw {raw.size/ms/1e6:.1f} GB/s  alg {alg/ms/1e6:.1f} GB/s ({alg/ms/1e6/6548.8*100:.1f}% HBM) path={path} ok={bool(torch.equal(d_out, d_raw))}", flush=True)
