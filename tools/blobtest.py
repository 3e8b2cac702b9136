import sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth
from oracle.pyoracle import Oracle
ctx = api.Context(0)
st = torch.cuda.current_stream().cuda_stream
w,h,ch=1920,1080,4
raw = synth.generate("photo", w, h, ch)
q = Oracle.encode(raw, w, h, ch)
d_q = torch.from_numpy(q).cuda()
d_out = torch.zeros(raw.size, dtype=torch.uint8, device="cuda")
bad=0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    d_out.zero_()
    ctx.decode_dev(d_q, q.size, w, h, ch, 0, 0, False, d_out, raw.size, st)
    path = ctx.decode_status(st)
    ok = bool(np.array_equal(d_out.cpu().numpy(), raw))
    if path or not ok: bad+=1; print("iter", it, "path", path, "ok", ok)
print("bad", bad, "of 20")
