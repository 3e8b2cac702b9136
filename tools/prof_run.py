"""Development aid for ncu captures: a few encode + decode calls of one bench workload (device buffers, C ABI).
usage: python tools/prof_run.py <kind> <w> <h> <ch> [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth_torch
kind, w, h, ch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
d_raw = synth_torch.generate(kind, w, h, ch, device="cuda")[0]
cap = (ch + 1) * w * h + 22
d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for it in range(reps):
    flush.fill_(it)
    ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st)
    n, ok = ctx.encode_status(st)
    flush.fill_(it + 100)
    ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st)
    torch.cuda.synchronize()
assert torch.equal(d_out, d_raw)
print(f"{kind} {w}x{h}x{ch}: {reps} encode + decode calls, stream {n} bytes, decode path {ctx.decode_status(st)}")
