"""Development aid: host-pointer calls with pageable vs page-locked buffers (4K RGB / 8K RGBA photo), wall clock."""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from qoipp_b200 import api, synth
from qoipp_b200._lib import Desc, lib
ctx = api.Context(0)
for (w, h, ch) in ((3840, 2160, 3), (7680, 4320, 4)):
    raw = synth.generate("photo", w, h, 3)
    if ch == 4: raw = np.concatenate([raw.reshape(-1, 3), np.full((w * h, 1), 255, np.uint8)], axis=1).reshape(-1)
    worst = (ch + 1) * w * h + 22
    for mode in ("pageable", "pinned"):
        if mode == "pinned":
            t_raw = torch.from_numpy(raw.copy()).pin_memory(); t_q = torch.empty(worst, dtype=torch.uint8).pin_memory(); t_o = torch.empty(raw.size, dtype=torch.uint8).pin_memory()
            a_raw, a_q, a_o = t_raw.numpy(), t_q.numpy(), t_o.numpy()
        else:
            a_raw, a_q, a_o = raw.copy(), np.empty(worst, np.uint8), np.empty(raw.size, np.uint8)
            a_q[:] = 0; a_o[:] = 0
        def once():
            wr, ok = C.c_uint64(0), C.c_int32(0)
            t0 = time.perf_counter()
            e = lib.qoipp_b200_encode_host(ctx._h, C.c_void_p(a_raw.ctypes.data), a_raw.size, C.byref(Desc(w, h, ch, 0)), C.c_void_p(a_q.ctypes.data), worst, C.byref(wr), C.byref(ok))
            t1 = time.perf_counter()
            assert e == 0 and ok.value
            d = Desc()
            e = lib.qoipp_b200_decode_host(ctx._h, C.c_void_p(a_q.ctypes.data), wr.value, 0, 0, C.c_void_p(a_o.ctypes.data), a_raw.size, C.byref(d))
            t2 = time.perf_counter()
            assert e == 0
            return t1 - t0, t2 - t1, wr.value
        for _ in range(3): once()
        r = [once() for _ in range(8)]
        te, td = np.median([x[0] for x in r]), np.median([x[1] for x in r])
        assert np.array_equal(a_o, raw)
        print(f"{w}x{h}x{ch} {mode:9s}: encode_host {te*1e3:7.3f} ms ({raw.size/te/1e9:5.1f} GB/s raw)  decode_host {td*1e3:7.3f} ms ({raw.size/td/1e9:5.1f} GB/s raw)  E={r[0][2]}")
# plain copies for reference
n = 132710400
hp = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda"); hn = torch.empty(n, dtype=torch.uint8)
for name, src, dst in (("H2D pinned", hp, d), ("D2H pinned", d, hp), ("H2D pageable", hn, d), ("D2H pageable", d, hn)):
    for _ in range(2): dst.copy_(src); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): dst.copy_(src); torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 5
    print(f"{name}: {n/t/1e9:.1f} GB/s")
