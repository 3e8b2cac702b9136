"""Development aid (-DQB_STATS build): how many tiles of a decode repaired themselves and what sent an image to a retry round."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth_torch
from qoipp_b200._lib import lib
lib.qoipp_b200_debug_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
for kind, w, h, ch in [("photo", 7680, 4320, 4), ("photo", 3840, 2160, 4), ("photo", 512, 512, 4)]:
    d_raw = synth_torch.generate(kind, w, h, ch, device="cuda")[0]
    cap = (ch + 1) * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
    ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st); n, ok = ctx.encode_status(st)
    for it in range(3):
        ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st); torch.cuda.synchronize()
        out = (C.c_uint32 * 4)()
        lib.qoipp_b200_debug_stats(ctx._h, out)
        print(f"{kind} {w}x{h}x{ch}: tiles {(n - 14 + 895) // 896}, path {ctx.decode_status(st)}: repaired or cascaded tiles {out[0]}, a successor must be decoded again because a pixel/alpha, slot or prev word "
              f"changed {out[1]}, because it read a changed table entry / had not published / the scan gave up {out[2]}")
