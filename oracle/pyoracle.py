"""ctypes doors into the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle``  -- oracle/liboracle.so, the plain-C restatement (oracle/qoi_oracle.c).
* ``Ref``     -- oracle/_ref/libqoipp_ref.so, the unmodified reference built from /root/reference
                 (oracle/Makefile).  May be absent; ``Ref.available()`` says so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package (qoipp_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libqoipp_ref.so")
REF_SO_V4 = os.path.join(HERE, "_ref", "libqoipp_ref_v4.so")  # same sources, -march=x86-64-v4 (bench only, AVX-512 hosts)

ERROR_NAMES = {
    0: "Ok", 1: "Empty", 2: "TooShort", 3: "TooBig", 4: "NotQoi", 5: "InvalidDesc", 6: "MismatchedDesc",
    7: "NotEnoughSpace", 8: "NotInitialized", 9: "AlreadyInitialized", 10: "NotRegularFile", 11: "FileExists",
    12: "FileNotExists", 13: "IoError", 14: "BadAlloc",
}

u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> None:
    """Compile the checkers (make -C oracle).  _ref is rebuilt only where /root/reference exists."""
    if force or not os.path.exists(ORACLE_SO) or (os.path.isdir("/root/reference/source") and not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


def _ptr(a):
    if a is None:
        return C.cast(None, u8p)
    return a.ctypes.data_as(u8p)


def _as_u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
    return np.frombuffer(bytes(b), dtype=np.uint8)


class QoDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint8), ("colorspace", C.c_uint8)]


class QoState(C.Structure):
    _fields_ = [("channels", C.c_uint8), ("target", C.c_uint8), ("run", C.c_uint8), ("prev", C.c_uint8 * 4),
                ("seen", C.c_uint8 * 256)]


class Oracle:
    """The plain-C restatement."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            L = C.CDLL(ORACLE_SO)
            L.qo_encode_core.restype = C.c_size_t
            L.qo_encode_core.argtypes = [u8p, C.POINTER(QoDesc), u8p, C.c_size_t, C.c_int, C.POINTER(C.c_int)]
            L.qo_decode_core.restype = None
            L.qo_decode_core.argtypes = [u8p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint8, u8p]
            L.qo_encode_into.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(QoDesc), C.POINTER(C.c_size_t),
                                         C.POINTER(C.c_int)]
            L.qo_decode_into.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_uint8, C.c_int, C.POINTER(QoDesc)]
            L.qo_read_header.argtypes = [u8p, C.c_size_t, C.POINTER(QoDesc)]
            L.qo_worst_size.argtypes = [C.POINTER(QoDesc), C.POINTER(C.c_size_t)]
            L.qo_count_bytes.argtypes = [C.POINTER(QoDesc), C.POINTER(C.c_size_t)]
            for n in ("qo_senc_init", "qo_senc_reset", "qo_sdec_init", "qo_sdec_reset"):
                getattr(L, n).restype = None
                getattr(L, n).argtypes = [C.POINTER(QoState)]
            L.qo_senc_initialize.argtypes = [C.POINTER(QoState), u8p, C.c_size_t, C.POINTER(QoDesc), C.POINTER(C.c_size_t)]
            L.qo_senc_encode.argtypes = [C.POINTER(QoState), u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t),
                                         C.POINTER(C.c_size_t)]
            L.qo_senc_finalize.argtypes = [C.POINTER(QoState), u8p, C.c_size_t, C.POINTER(C.c_size_t)]
            L.qo_sdec_initialize.argtypes = [C.POINTER(QoState), u8p, C.c_size_t, C.c_uint8, C.POINTER(QoDesc)]
            L.qo_sdec_decode.argtypes = [C.POINTER(QoState), u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t),
                                         C.POINTER(C.c_size_t)]
            L.qo_sdec_drain_run.argtypes = [C.POINTER(QoState), u8p, C.c_size_t, C.POINTER(C.c_size_t)]
            cls._lib = L
        return cls._lib

    # ---- one-shot
    @classmethod
    def worst_size(cls, w, h, ch, cs=0):
        out = C.c_size_t(0)
        e = cls.lib().qo_worst_size(C.byref(QoDesc(w, h, ch, cs)), C.byref(out))
        return e, out.value

    @classmethod
    def encode_into(cls, raw, w, h, ch, cs=0, cap=None):
        """-> (err, bytes written (np.uint8 array of length cap, valid prefix = written), written, complete)"""
        raw = _as_u8(raw)
        if cap is None:
            cap = (ch + 1) * w * h + 22
        out = np.full(max(cap, 1), 0xAA, dtype=np.uint8)
        written, complete = C.c_size_t(0), C.c_int(0)
        e = cls.lib().qo_encode_into(_ptr(out), cap, _ptr(raw), raw.size, C.byref(QoDesc(w, h, ch, cs)),
                                     C.byref(written), C.byref(complete))
        return e, out[:cap], written.value, bool(complete.value)

    @classmethod
    def encode(cls, raw, w, h, ch, cs=0) -> np.ndarray:
        e, out, n, ok = cls.encode_into(raw, w, h, ch, cs)
        assert e == 0 and ok, (e, ok)
        return out[:n].copy()

    @classmethod
    def decode_into(cls, qoi, target=0, flip=False, cap=None):
        """-> (err, pixels or None, (w,h,ch,cs))"""
        qoi = _as_u8(qoi)
        d = QoDesc()
        if cap is None:
            hd = QoDesc()
            if cls.lib().qo_read_header(_ptr(qoi), qoi.size, C.byref(hd)) == 0:
                cap = hd.width * hd.height * max(target or hd.channels, hd.channels)
            else:
                cap = 16
        out = np.full(max(cap, 1), 0xAA, dtype=np.uint8)
        e = cls.lib().qo_decode_into(_ptr(out), cap, _ptr(qoi), qoi.size, target, int(flip), C.byref(d))
        if e:
            return e, None, None
        return 0, out[: d.width * d.height * d.channels].copy(), (d.width, d.height, d.channels, d.colorspace)

    @classmethod
    def decode(cls, qoi, target=0, flip=False) -> np.ndarray:
        e, px, _ = cls.decode_into(qoi, target, flip)
        assert e == 0, e
        return px

    @classmethod
    def read_header(cls, qoi):
        qoi = _as_u8(qoi)
        d = QoDesc()
        e = cls.lib().qo_read_header(_ptr(qoi), qoi.size, C.byref(d))
        return e, (d.width, d.height, d.channels, d.colorspace)

    # ---- streams
    class StreamEncoder:
        def __init__(self):
            self.L = Oracle.lib()
            self.s = QoState()
            self.L.qo_senc_init(C.byref(self.s))

        def initialize(self, out, w, h, ch, cs=0):
            n = C.c_size_t(0)
            e = self.L.qo_senc_initialize(C.byref(self.s), _ptr(out), out.size, C.byref(QoDesc(w, h, ch, cs)), C.byref(n))
            return e, n.value

        def encode(self, out, inp):
            p, n = C.c_size_t(0), C.c_size_t(0)
            e = self.L.qo_senc_encode(C.byref(self.s), _ptr(out), out.size, _ptr(inp), inp.size, C.byref(p), C.byref(n))
            return e, p.value, n.value

        def finalize(self, out):
            n = C.c_size_t(0)
            e = self.L.qo_senc_finalize(C.byref(self.s), _ptr(out), out.size, C.byref(n))
            return e, n.value

        def reset(self):
            self.L.qo_senc_reset(C.byref(self.s))

        def has_run_count(self):
            return self.s.run > 0

        def is_initialized(self):
            return self.s.channels != 0

    class StreamDecoder:
        def __init__(self):
            self.L = Oracle.lib()
            self.s = QoState()
            self.L.qo_sdec_init(C.byref(self.s))

        def initialize(self, inp, target=0):
            d = QoDesc()
            e = self.L.qo_sdec_initialize(C.byref(self.s), _ptr(inp), inp.size, target, C.byref(d))
            return e, (d.width, d.height, d.channels, d.colorspace)

        def decode(self, out, inp):
            p, n = C.c_size_t(0), C.c_size_t(0)
            e = self.L.qo_sdec_decode(C.byref(self.s), _ptr(out), out.size, _ptr(inp), inp.size, C.byref(p), C.byref(n))
            return e, p.value, n.value

        def drain_run(self, out):
            n = C.c_size_t(0)
            e = self.L.qo_sdec_drain_run(C.byref(self.s), _ptr(out), out.size, C.byref(n))
            return e, n.value

        def reset(self):
            self.L.qo_sdec_reset(C.byref(self.s))

        def has_run_count(self):
            return self.s.run > 0

        def is_initialized(self):
            return self.s.channels != 0


class Ref:
    """The unmodified reference behind oracle/ref_shim.cpp."""

    _lib = None

    @staticmethod
    def available() -> bool:
        build()
        return os.path.exists(REF_SO)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            L = C.CDLL(REF_SO)
            u64p, ip, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_uint32)
            L.ref_worst_size.argtypes = [C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8, u64p]
            L.ref_read_header.argtypes = [u8p, C.c_uint64, u32p]
            L.ref_encode_into.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8,
                                          u64p, ip]
            L.ref_decode_into.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, C.c_uint8, C.c_int, u32p]
            L.ref_decode.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, C.c_uint8, C.c_int, u32p]
            for n in ("ref_senc_new", "ref_sdec_new"):
                getattr(L, n).restype = C.c_void_p
                getattr(L, n).argtypes = []
            for n in ("ref_senc_delete", "ref_sdec_delete", "ref_senc_reset", "ref_sdec_reset"):
                getattr(L, n).restype = None
                getattr(L, n).argtypes = [C.c_void_p]
            L.ref_senc_has_run.argtypes = [C.c_void_p]
            L.ref_sdec_run_count.argtypes = [C.c_void_p]
            L.ref_senc_initialize.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8, u64p]
            L.ref_senc_encode.argtypes = [C.c_void_p, u8p, C.c_uint64, u8p, C.c_uint64, u64p, u64p]
            L.ref_senc_finalize.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p]
            L.ref_sdec_initialize.argtypes = [C.c_void_p, u8p, C.c_uint64, C.c_uint8, u32p]
            L.ref_sdec_decode.argtypes = [C.c_void_p, u8p, C.c_uint64, u8p, C.c_uint64, u64p, u64p]
            L.ref_sdec_drain_run.argtypes = [C.c_void_p, u8p, C.c_uint64, u64p]
            cls._bind_bench(L)
            cls._lib = L
        return cls._lib

    @staticmethod
    def _bind_bench(L):
        L.ref_bench.argtypes = [C.POINTER(u8p), C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8, C.c_int, C.c_int,
                                C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]

    _bench_lib = None

    @classmethod
    def bench_lib(cls):
        """Library for the timed CPU baseline: the x86-64-v4 build when this host has AVX-512, else the portable one."""
        if cls._bench_lib is None:
            so, march = REF_SO, "x86-64-v3"
            try:
                flags = next(ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")).split()
                if os.path.exists(REF_SO_V4) and all(f in flags for f in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl")):
                    so, march = REF_SO_V4, "x86-64-v4"
            except Exception:
                pass
            L = C.CDLL(so)
            cls._bind_bench(L)
            cls._bench_lib = (L, march)
        return cls._bench_lib

    @classmethod
    def bench(cls, images, w, h, ch, cs=0, threads=0, warmups=3, reps=5):
        """The reference's own benchmark method (example/source/04_bench.cpp:445-510,733-754) on `images` (list of raw arrays):
        encode_into a pre-allocated, pre-touched worst_size buffer and decode_into a pre-allocated buffer, thread-per-image.
        Returns dict(enc_s, dec_s, enc_bytes, threads, march): seconds for `reps` passes over all images per direction."""
        L, march = cls.bench_lib()
        imgs = [_as_u8(i) for i in images]
        arr = (u8p * len(imgs))(*[_ptr(i) for i in imgs])
        es, ds, nb, tu = C.c_double(0), C.c_double(0), C.c_uint64(0), C.c_int(0)
        e = L.ref_bench(arr, len(imgs), imgs[0].size, w, h, ch, cs, threads, warmups, reps, C.byref(es), C.byref(ds), C.byref(nb), C.byref(tu))
        if e != 0:
            raise RuntimeError(f"ref_bench failed: {e}")
        return {"enc_s": es.value, "dec_s": ds.value, "enc_bytes": nb.value, "threads": tu.value, "march": march}

    @classmethod
    def worst_size(cls, w, h, ch, cs=0):
        out = C.c_uint64(0)
        e = cls.lib().ref_worst_size(w, h, ch, cs, C.byref(out))
        return e, out.value

    @classmethod
    def encode_into(cls, raw, w, h, ch, cs=0, cap=None):
        raw = _as_u8(raw)
        if cap is None:
            cap = (ch + 1) * w * h + 22
        out = np.full(max(cap, 1), 0xAA, dtype=np.uint8)
        written, complete = C.c_uint64(0), C.c_int(0)
        e = cls.lib().ref_encode_into(_ptr(out), cap, _ptr(raw), raw.size, w, h, ch, cs, C.byref(written), C.byref(complete))
        return e, out[:cap], written.value, bool(complete.value)

    @classmethod
    def encode(cls, raw, w, h, ch, cs=0) -> np.ndarray:
        e, out, n, ok = cls.encode_into(raw, w, h, ch, cs)
        assert e == 0 and ok, (e, ok)
        return out[:n].copy()

    @classmethod
    def decode(cls, qoi, target=0, flip=False) -> np.ndarray:
        """qoipp::decode (allocating) -- safe for every target."""
        qoi = _as_u8(qoi)
        e, hd = cls.read_header(qoi)
        cap = hd[0] * hd[1] * 4 if e == 0 else 16
        out = np.full(max(cap, 1), 0xAA, dtype=np.uint8)
        d = (C.c_uint32 * 4)()
        e = cls.lib().ref_decode(_ptr(out), cap, _ptr(qoi), qoi.size, target, int(flip), d)
        assert e == 0, e
        return out[: d[0] * d[1] * d[2]].copy()

    @classmethod
    def decode_err(cls, qoi, target=0, flip=False) -> int:
        qoi = _as_u8(qoi)
        out = np.zeros(16, dtype=np.uint8)
        d = (C.c_uint32 * 4)()
        e = cls.lib().ref_decode(_ptr(out), 0, _ptr(qoi), qoi.size, target, int(flip), d)
        return e

    @classmethod
    def decode_into(cls, qoi, cap, target=0, flip=False):
        qoi = _as_u8(qoi)
        out = np.full(max(cap, 1), 0xAA, dtype=np.uint8)
        d = (C.c_uint32 * 4)()
        e = cls.lib().ref_decode_into(_ptr(out), cap, _ptr(qoi), qoi.size, target, int(flip), d)
        return e, out[:cap], tuple(d)

    @classmethod
    def read_header(cls, qoi):
        qoi = _as_u8(qoi)
        d = (C.c_uint32 * 4)()
        e = cls.lib().ref_read_header(_ptr(qoi), qoi.size, d)
        return e, tuple(d)

    class StreamEncoder:
        def __init__(self):
            self.L = Ref.lib()
            self.p = self.L.ref_senc_new()

        def __del__(self):
            if getattr(self, "p", None):
                self.L.ref_senc_delete(self.p)
                self.p = None

        def initialize(self, out, w, h, ch, cs=0):
            n = C.c_uint64(0)
            e = self.L.ref_senc_initialize(self.p, _ptr(out), out.size, w, h, ch, cs, C.byref(n))
            return e, n.value

        def encode(self, out, inp):
            p, n = C.c_uint64(0), C.c_uint64(0)
            e = self.L.ref_senc_encode(self.p, _ptr(out), out.size, _ptr(inp), inp.size, C.byref(p), C.byref(n))
            return e, p.value, n.value

        def finalize(self, out):
            n = C.c_uint64(0)
            e = self.L.ref_senc_finalize(self.p, _ptr(out), out.size, C.byref(n))
            return e, n.value

        def reset(self):
            self.L.ref_senc_reset(self.p)

        def has_run_count(self):
            return bool(self.L.ref_senc_has_run(self.p))

    class StreamDecoder:
        def __init__(self):
            self.L = Ref.lib()
            self.p = self.L.ref_sdec_new()

        def __del__(self):
            if getattr(self, "p", None):
                self.L.ref_sdec_delete(self.p)
                self.p = None

        def initialize(self, inp, target=0):
            d = (C.c_uint32 * 4)()
            e = self.L.ref_sdec_initialize(self.p, _ptr(inp), inp.size, target, d)
            return e, tuple(d)

        def decode(self, out, inp):
            p, n = C.c_uint64(0), C.c_uint64(0)
            e = self.L.ref_sdec_decode(self.p, _ptr(out), out.size, _ptr(inp), inp.size, C.byref(p), C.byref(n))
            return e, p.value, n.value

        def drain_run(self, out):
            n = C.c_uint64(0)
            e = self.L.ref_sdec_drain_run(self.p, _ptr(out), out.size, C.byref(n))
            return e, n.value

        def reset(self):
            self.L.ref_sdec_reset(self.p)

        def has_run_count(self):
            return self.L.ref_sdec_run_count(self.p) > 0
