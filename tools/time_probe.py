"""Development aid: device-timed decode of a few workloads with the library named by QOIPP_B200_SO; no correctness check
(for timing experiments with deliberately incomplete kernels); prints the fraction of matching bytes."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth_torch
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for kind, w, h, ch in [("photo", 3840, 2160, 3), ("photo_opaque", 7680, 4320, 4), ("photo", 7680, 4320, 4), ("noise", 3840, 2160, 4)]:
    d_raw = synth_torch.generate(kind, w, h, ch, device="cuda")[0]
    cap = (ch + 1) * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
    ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st); n, ok = ctx.encode_status(st)
    ts = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(it); e0.record(); ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    path = ctx.decode_status(st)
    print(f"{kind:13s} {w}x{h}x{ch}: decode {np.median(ts[2:]):8.1f} us  path={path}  matching bytes {float((d_out == d_raw).float().mean()):.4f}", flush=True)
