"""Development aid: per-tile phase timing of decode_ts_kernel with the -DQB_TIMING build (libqoipp_b200_timing.so)."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import synth
from qoipp_b200._lib import Desc
L = C.CDLL("qoipp_b200/libqoipp_b200_timing.so")
ctx = C.c_void_p(); assert L.qoipp_b200_ctx_create(0, C.byref(ctx)) == 0
st = torch.cuda.current_stream().cuda_stream
names = ["stage", "parse+lb", "W1+lbs", "W2", "merge", "chase+publish", "state lb", "W3"]
for kind, w, h, ch in [("photo", 3840, 2160, 3), ("photo", 7680, 4320, 4)]:
    raw = synth.generate(kind, w, h, 3)
    if ch == 4: raw = np.concatenate([raw.reshape(-1, 3), np.full((w * h, 1), 255, np.uint8)], axis=1).reshape(-1)
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
    old_tiles = (n - 14 + 2047) // 2048; ntiles = (n - 14 + 2175) // 2176
    buf = torch.empty((old_tiles + ntiles) * 72 * 8, dtype=torch.uint8, device="cuda")
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(buf.data_ptr()), p, C.c_size_t(buf.numel()), 3)
    words = buf.cpu().numpy().view(np.uint32).reshape(old_tiles + ntiles, 144)[old_tiles:]
    t = words[:, 136:144].astype(np.int64)
    t = t[(t[:, -1] > 0) & (t[:, -1] < 10**8)]
    d = np.diff(np.concatenate([np.zeros((t.shape[0], 1), np.int64), t], axis=1), axis=1)
    print(f"DECODE-TS {kind} {w}x{h}x{ch}: tiles {ntiles} ({t.shape[0]} stamped); median total {np.median(t[:, -1]):.0f} cyc")
    for i, nm in enumerate(names):
        print(f"   {nm:18s} median {np.median(d[:, i]):8.0f}  p10 {np.percentile(d[:, i], 10):8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}")
