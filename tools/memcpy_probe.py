import ctypes as C, time, numpy as np, torch
rt = C.cdll.LoadLibrary("libcudart.so.12")
n = 132710400
d = torch.empty(n, dtype=torch.uint8, device="cuda")
a = np.empty(n, np.uint8); a[:] = 1
t = torch.empty(n, dtype=torch.uint8); t.fill_(1)
def bench(name, hostptr, kind):
    for _ in range(2):
        rt.cudaMemcpy(C.c_void_p(d.data_ptr()), C.c_void_p(hostptr), C.c_size_t(n), kind) if kind == 1 else rt.cudaMemcpy(C.c_void_p(hostptr), C.c_void_p(d.data_ptr()), C.c_size_t(n), kind)
    t0 = time.perf_counter()
    for _ in range(5):
        if kind == 1: rt.cudaMemcpy(C.c_void_p(d.data_ptr()), C.c_void_p(hostptr), C.c_size_t(n), 1)
        else: rt.cudaMemcpy(C.c_void_p(hostptr), C.c_void_p(d.data_ptr()), C.c_size_t(n), 2)
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {n/dt/1e9:.1f} GB/s")
bench("numpy H2D", a.ctypes.data, 1); bench("numpy D2H", a.ctypes.data, 2)
bench("torch H2D", t.data_ptr(), 1); bench("torch D2H", t.data_ptr(), 2)
# register in place
t0 = time.perf_counter(); e = rt.cudaHostRegister(C.c_void_p(a.ctypes.data), C.c_size_t(n), 0); t1 = time.perf_counter()
print("cudaHostRegister", e, f"{(t1-t0)*1e3:.2f} ms")
bench("registered numpy H2D", a.ctypes.data, 1); bench("registered numpy D2H", a.ctypes.data, 2)
t0 = time.perf_counter(); rt.cudaHostUnregister(C.c_void_p(a.ctypes.data)); t1 = time.perf_counter()
print("cudaHostUnregister", f"{(t1-t0)*1e3:.2f} ms")
t0 = time.perf_counter(); e = rt.cudaHostRegister(C.c_void_p(a.ctypes.data), C.c_size_t(n), 0); t1 = time.perf_counter()
print("cudaHostRegister again", e, f"{(t1-t0)*1e3:.2f} ms")
