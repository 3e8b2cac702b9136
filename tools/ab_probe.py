"""Development aid: device-timed encode+decode of two workloads with the library named by QOIPP_B200_SO."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for kind, w, h, ch in [("photo", 3840, 2160, 3), ("photo", 7680, 4320, 4)]:
    raw = synth.generate(kind, w, h, 3)
    if ch == 4:
        raw = np.concatenate([raw.reshape(-1, 3), np.full((w * h, 1), 255, np.uint8)], axis=1).reshape(-1)
    d_raw = torch.from_numpy(raw).cuda(); cap = (ch + 1) * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
    ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st); n, ok = ctx.encode_status(st)
    te, td = [], []
    for it in range(13):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        flush.fill_(it); ev[0].record(); ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st); ev[1].record()
        flush.fill_(it + 1); ev[2].record(); ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st); ev[3].record()
        torch.cuda.synchronize()
        if it >= 3: te.append(ev[0].elapsed_time(ev[1])); td.append(ev[2].elapsed_time(ev[3]))
    assert torch.equal(d_out, d_raw)
    print(f"{os.environ.get('QOIPP_B200_SO', 'default'):45s} {kind} {w}x{h}x{ch}: encode {np.median(te)*1e3:7.1f} us  decode {np.median(td)*1e3:7.1f} us", flush=True)
