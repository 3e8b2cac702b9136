"""Development aid: the resumable decode in 64 MiB chunks of input against the one-shot decode of the same 8K stream."""
import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth_torch
from qoipp_b200._lib import lib
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
for kind, w, h, ch in [("photo_opaque", 7680, 4320, 4), ("photo", 7680, 4320, 4), ("noise", 7680, 4320, 4)]:
    d_raw = synth_torch.generate(kind, w, h, ch, device="cuda")[0]
    cap = (ch + 1) * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
    ctx.encode_dev(d_raw, w, h, ch, 0, d_q, cap, st); n, ok = ctx.encode_status(st)
    def oneshot():
        ctx.decode_dev(d_q, n, w, h, ch, 0, 0, False, d_out, d_out.numel(), st)
    seen = np.zeros(66, np.uint32); seen[0] = 0xFF000000; seen[2 + 53] = 0xFF000000
    d_res = torch.zeros(16, dtype=torch.uint8, device="cuda")
    def chunks(chunk):
        s = torch.from_numpy(seen.view(np.uint8).copy()).cuda()
        off, wr = 14, 0
        while off < n - 8:
            take = min(chunk, n - 8 - off)
            e = lib.qoipp_b200_stream_decode_dev(ctx._h, ch, C.c_void_p(s.data_ptr()), C.c_void_p(d_q[off:].data_ptr()), take, C.c_void_p(d_out[wr:].data_ptr()),
                                                 d_out.numel() - wr, C.c_void_p(d_res.data_ptr()), C.c_void_p(st))
            assert e == 0
            p, wn = (int(x) for x in d_res.cpu().numpy().view(np.uint64))  # the host loop needs the counts: one sync per chunk
            off += p; wr += wn
            assert p > 0
        return wr
    for name, fn in (("one-shot", oneshot), ("64 MiB chunks", lambda: chunks(64 << 20)), ("8 MiB chunks", lambda: chunks(8 << 20))):
        ts = []
        for it in range(5):
            d_out.zero_(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            assert torch.equal(d_out, d_raw), name
        print(f"{kind:13s} stream {n / 1e6:6.1f} MB  {name:14s} {np.median(ts[1:]) * 1e6:9.1f} us (host clock, sync per call)", flush=True)
