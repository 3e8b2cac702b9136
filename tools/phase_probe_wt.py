"""Development aid: per-tile phase timing of decode_wt_kernel with the -DQB_TIMING build (build_ab/timing.so):
SM-clock stamps taken by lane 0 at the phase boundaries of every tile (words 72..79 of the tile's carry record)."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import synth
from qoipp_b200._lib import Desc

L = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else "build_ab/timing.so")
ctx = C.c_void_p(); assert L.qoipp_b200_ctx_create(0, C.byref(ctx)) == 0
st = torch.cuda.current_stream().cuda_stream
# (word, slot) in execution order
order = [("stage", 72, 0), ("parse+lb1", 72, 1), ("counts", 73, 0), ("walk loop", 77, 1), ("gather", 76, 0), ("entry nodes", 73, 1),
         ("lb3 slot", 76, 1), ("lastk", 74, 0), ("index", 74, 1), ("transfer", 77, 0), ("state lb", 75, 0), ("emit", 75, 1)]
TB = 896
for kind, w, h, ch in [("photo", 3840, 2160, 3), ("photo", 7680, 4320, 4), ("noise", 3840, 2160, 4)]:
    raw = synth.generate(kind, w, h, ch)
    if ch == 4 and kind == "photo": raw = raw.copy(); raw[3::4] = 255
    d_raw = torch.from_numpy(raw).cuda(); cap = (ch + 1) * w * h + 22
    d_q = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    assert L.qoipp_b200_encode_dev(ctx, C.c_void_p(d_raw.data_ptr()), C.byref(Desc(w, h, ch, 0)), C.c_void_p(d_q.data_ptr()), C.c_uint64(cap), C.c_void_p(st)) == 0
    wr, ok = C.c_uint64(), C.c_int32()
    L.qoipp_b200_encode_status(ctx, C.c_void_p(st), C.byref(wr), C.byref(ok))
    n = wr.value
    d_out = torch.zeros(w * h * ch, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        assert L.qoipp_b200_decode_dev(ctx, C.c_void_p(d_q.data_ptr()), C.c_uint64(n), C.byref(Desc(w, h, ch, 0)), C.c_uint8(0), C.c_int32(0), C.c_void_p(d_out.data_ptr()), C.c_uint64(d_out.numel()), C.c_void_p(st)) == 0
    torch.cuda.synchronize()
    assert torch.equal(d_out, d_raw)
    p, nb = C.c_void_p(), C.c_uint64()
    L.qoipp_b200_debug_carry(ctx, C.byref(p), C.byref(nb))
    ntiles = (n - 14 + TB - 1) // TB
    buf = torch.empty(ntiles * 80 * 8, dtype=torch.uint8, device="cuda")
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(buf.data_ptr()), p, C.c_size_t(buf.numel()), 3)
    words = buf.cpu().numpy().view(np.uint32).reshape(ntiles, 160)
    t = np.stack([words[:, 2 * wd + sl] for _, wd, sl in order], axis=1).astype(np.int64)
    d = np.diff(np.concatenate([np.zeros((ntiles, 1), np.int64), t], axis=1), axis=1)
    mid = slice(ntiles // 4, 3 * ntiles // 4)
    print(f"DECODE {kind} {w}x{h}x{ch}: tiles {ntiles}; cycles per phase (middle half of the tiles), total median {np.median(t[mid, -1]):.0f} cyc")
    for i, (nm, _, _) in enumerate(order):
        print(f"   {nm:14s} median {np.median(d[mid, i]):8.0f}  mean {np.mean(d[mid, i]):8.0f}  p90 {np.percentile(d[mid, i], 90):8.0f}")
