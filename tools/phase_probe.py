"""Development aid: per-tile phase timing of encode_kernel with the -DQB_TIMING build (libqoipp_b200_timing.so)."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import synth
from qoipp_b200._lib import Desc

L = C.CDLL("qoipp_b200/libqoipp_b200_timing.so")
ctx = C.c_void_p(); assert L.qoipp_b200_ctx_create(0, C.byref(ctx)) == 0
st = torch.cuda.current_stream().cuda_stream
names = ["ticket", "loads", "phaseA", "lookback12", "fixups+C1", "lookback3+stage", "copyout"]
for kind, w, h, ch in [("photo", 3840, 2160, 3), ("photo", 7680, 4320, 4), ("flat", 3840, 2160, 4)]:
    raw = synth.generate(kind, w, h, ch)
    if ch == 4: raw = raw.copy(); raw[3::4] = 255
    d_raw = torch.from_numpy(raw).cuda(); cap = (ch + 1) * w * h + 22
    d_out = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        assert L.qoipp_b200_encode_dev(ctx, C.c_void_p(d_raw.data_ptr()), C.byref(Desc(w, h, ch, 0)), C.c_void_p(d_out.data_ptr()), C.c_uint64(cap), C.c_void_p(st)) == 0
    torch.cuda.synchronize()
    p, n = C.c_void_p(), C.c_uint64()
    L.qoipp_b200_debug_carry(ctx, C.byref(p), C.byref(n))
    ntiles = (w * h + 2047) // 2048
    buf = torch.empty(ntiles * 72 * 8, dtype=torch.uint8, device="cuda")
    import torch.cuda
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(buf.data_ptr()), p, C.c_size_t(buf.numel()), 3)
    words = buf.cpu().numpy().view(np.uint32).reshape(ntiles, 144)
    # stamps: word66=(ticket, phaseA) word67=(lookback12, fixups) word68=(stage, copyout) word69=(loads,)
    t = np.stack([words[:, 132], words[:, 138], words[:, 133], words[:, 134], words[:, 135], words[:, 136], words[:, 137]], axis=1).astype(np.int64)
    d = np.diff(np.concatenate([np.zeros((ntiles, 1), np.int64), t], axis=1), axis=1)
    mid = slice(ntiles // 4, 3 * ntiles // 4)
    print(f"{kind} {w}x{h}x{ch}: tiles {ntiles}; median cycles per phase (middle half of tiles), total {np.median(t[mid, -1]):.0f} cyc")
    for i, nm in enumerate(names):
        print(f"   {nm:18s} median {np.median(d[mid, i]):8.0f}  p90 {np.percentile(d[mid, i], 90):8.0f}")


# ---- decode
dnames = ["ticket+stage", "parse+lb1", "walk+lb2", "records", "writers", "jumping", "state lb3", "values+stores"]
for kind, w, h, ch in [("photo", 3840, 2160, 3), ("photo", 7680, 4320, 4), ("flat", 3840, 2160, 4)]:
    raw = synth.generate(kind, w, h, ch)
    if ch == 4: raw = raw.copy(); raw[3::4] = 255
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
    ntiles = (n - 14 + 2047) // 2048
    buf = torch.empty(ntiles * 72 * 8, dtype=torch.uint8, device="cuda")
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(buf.data_ptr()), p, C.c_size_t(buf.numel()), 3)
    words = buf.cpu().numpy().view(np.uint32).reshape(ntiles, 144)
    t = words[:, 136:144].astype(np.int64)
    d = np.diff(np.concatenate([np.zeros((ntiles, 1), np.int64), t], axis=1), axis=1)
    mid = slice(ntiles // 4, 3 * ntiles // 4)
    print(f"DECODE {kind} {w}x{h}x{ch}: tiles {ntiles}; median cycles per phase, total {np.median(t[mid, -1]):.0f} cyc")
    for i, nm in enumerate(dnames):
        print(f"   {nm:18s} median {np.median(d[mid, i]):8.0f}  p90 {np.percentile(d[mid, i], 90):8.0f}")
