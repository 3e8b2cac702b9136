"""Development aid: a few device-timed decodes of one workload (for ncu captures of the decode kernels)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth
kind, w, h, ch = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else ("photo", 3840, 2160, 3)
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
raw = synth.generate(kind, w, h, 3 if kind == "photo" else ch)
if kind == "photo" and ch == 4:
    raw = np.concatenate([raw.reshape(-1, 3), np.full((w * h, 1), 255, np.uint8)], axis=1).reshape(-1)
d_raw = torch.from_numpy(raw).cuda(); cap = (ch + 1) * w * h + 22
d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st); n, ok = ctx.encode_status(st)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ts = []
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.fill_(it); e0.record(); ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
path = ctx.decode_status(st)
assert torch.equal(d_out, d_raw)
print(f"{kind} {w}x{h}x{ch}: decode {np.median(ts[2:]):.1f} us, stream {n} bytes, path={path}")
