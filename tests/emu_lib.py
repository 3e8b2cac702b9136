"""ctypes door into tests/emu/libqb_emu.so: the product kernels stepped by the CPU SIMT emulator.
TEST INFRASTRUCTURE ONLY -- lets kernel logic be checked against the oracle on hosts without a GPU."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
SO = os.environ.get("QB_EMU_SO", os.path.join(HERE, "libqb_emu.so"))  # tools/emu_asan.sh points this at the ASan + UBSan build
u8p = C.POINTER(C.c_uint8)
_lib = None


class State(C.Structure):  # qoipp_b200_state
    _fields_ = [("channels", C.c_uint8), ("target", C.c_uint8), ("run", C.c_uint8), ("reserved", C.c_uint8),
                ("prev", C.c_uint8 * 4), ("seen", C.c_uint8 * 256)]


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        L = C.CDLL(SO)
        L.emu_encode.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint8, u8p, C.c_uint64,
                                 C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_uint64]
        L.emu_stream_encode.argtypes = [C.POINTER(State), u8p, C.c_uint64, u8p, C.c_uint64, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_uint64]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(u8p)


def encode(raw, w, h, ch, cs=0, cap=None, K=1, resident=4, seed=0, n_images=1):
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    worst = (ch + 1) * w * h + 22
    if cap is None:
        cap = worst
    stride = (max(cap, 1) + 15) // 16 * 16 + 16
    out = np.full(stride * n_images, 0xAA, dtype=np.uint8)
    written = (C.c_uint64 * n_images)()
    complete = (C.c_int * n_images)()
    rc = lib().emu_encode(_p(raw), w * h * ch, n_images, w, h, ch, cs, _p(out), stride, cap, written, complete, K, resident, seed)
    assert rc == 0, rc
    if n_images == 1:
        return out[:cap], int(written[0]), bool(complete[0])
    return [(out[i * stride: i * stride + cap], int(written[i]), bool(complete[i])) for i in range(n_images)]


class StreamEncoder:
    """Host logic of StreamEncoder (initialize/finalize are header/marker writes) over the emulated kernel."""

    def __init__(self, K=1, resident=4, seed=0):
        self.s = State()
        self.K, self.resident, self.seed = K, resident, seed
        self.reset()

    def reset(self):
        C.memset(C.byref(self.s), 0, C.sizeof(self.s))
        self.s.prev[3] = 255

    def initialize(self, out, w, h, ch, cs=0):
        if self.s.channels:
            return 9, 0
        if out.size == 0:
            return 1, 0
        if out.size < 14:
            return 2, 0
        out[:14] = np.frombuffer(b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([ch, cs]), dtype=np.uint8)
        self.s.channels = ch
        return 0, 14

    def encode(self, out, inp):
        if not self.s.channels:
            return 8, 0, 0
        if out.size == 0 or inp.size == 0:
            return 1, 0, 0
        if out.size < 5:
            return 2, 0, 0
        p, n = C.c_uint64(0), C.c_uint64(0)
        inp = np.ascontiguousarray(inp)
        rc = lib().emu_stream_encode(C.byref(self.s), _p(inp), inp.size, _p(out), out.size, C.byref(p), C.byref(n), self.K,
                                     self.resident, self.seed)
        assert rc == 0
        return 0, p.value, n.value

    def has_run_count(self):
        return self.s.run > 0

    def finalize(self, out):
        if not self.s.channels:
            return 8, 0
        if out.size == 0:
            return 1, 0
        need = 8 + (self.s.run > 0)
        if out.size < need:
            return 2, 0
        k = 0
        if self.s.run:
            out[0] = 0xC0 | (self.s.run - 1)
            k = 1
        out[k: k + 8] = [0, 0, 0, 0, 0, 0, 0, 1]
        self.reset()
        return 0, need


def _setup_decode():
    L = lib()
    L.emu_decode.argtypes = [u8p, C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, C.c_int, u8p, C.c_uint64,
                             C.POINTER(C.c_int), C.c_int, C.c_int, C.c_uint64]
    L.emu_stream_decode.argtypes = [C.POINTER(State), u8p, C.c_uint64, u8p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int, C.c_int,
                                    C.c_uint64, C.POINTER(C.c_int)]
    return L


def decode(qoi_list, w, h, target, flip=False, force_serial=False, resident=4, seed=0):
    """Decode one or several streams of equal shape in one launch -> (list of pixel arrays, list of path flags)."""
    if isinstance(qoi_list, np.ndarray):
        qoi_list = [qoi_list]
    L = _setup_decode()
    n = len(qoi_list)
    offs = np.zeros(n + 1, dtype=np.uint64)
    pad = 3  # deliberately misaligned stream starts
    blob = np.zeros(sum(q.size + pad for q in qoi_list) + 64, dtype=np.uint8)
    pos = 1
    starts = []
    for i, q in enumerate(qoi_list):
        blob[pos: pos + q.size] = q
        starts.append(pos)
        pos += q.size + pad
    # batch offsets must be contiguous ranges [offsets[k], offsets[k+1]) -> lay the streams out back to back instead
    blob2 = np.concatenate([np.zeros(1, np.uint8)] + [np.ascontiguousarray(q, dtype=np.uint8) for q in qoi_list] + [np.zeros(64, np.uint8)])
    o = 1
    for i, q in enumerate(qoi_list):
        offs[i] = o
        o += q.size
    offs[n] = o
    stride = w * h * target + 5
    out = np.full(stride * n + 16, 0xAA, dtype=np.uint8)
    path = (C.c_int * n)()
    rc = L.emu_decode(_p(blob2), offs.ctypes.data_as(C.POINTER(C.c_uint64)), n, w, h, target, int(flip), _p(out), stride, path,
                      int(force_serial), resident, seed)
    assert rc == 0, rc
    px = [out[i * stride: i * stride + w * h * target].copy() for i in range(n)]
    for i in range(n):
        assert np.all(out[i * stride + w * h * target: (i + 1) * stride] == 0xAA), "decode wrote past the image"
    return px, [int(x) for x in path]


class StreamDecoder:
    def __init__(self, parallel=False, resident=3, seed=0):
        """parallel: decode_wt_stream_kernel for every call (the device library does so from a few KB on)"""
        self.s = State()
        self.parallel, self.resident, self.seed = parallel, resident, seed
        self.serial_calls = 0
        self.reset()

    def reset(self):
        C.memset(C.byref(self.s), 0, C.sizeof(self.s))
        self.s.prev[3] = 255

    def initialize(self, inp, target=0):
        from oracle.pyoracle import Oracle  # header parsing is host code; reuse the checker's rule table here
        if self.s.channels:
            return 9, None
        e, d = Oracle.read_header(inp)
        if e:
            return e, None
        ch = target or d[2]
        self.s.channels = self.s.target = ch
        self.s.seen[53 * 4 + 3] = 255
        return 0, (d[0], d[1], ch, d[3])

    def decode(self, out, inp):
        if not self.s.channels:
            return 8, 0, 0
        if out.size == 0:
            return 1, 0, 0
        if out.size < self.s.channels:
            return 2, 0, 0
        L = _setup_decode()
        p, n = C.c_uint64(0), C.c_uint64(0)
        inp = np.ascontiguousarray(inp)
        used = C.c_int(0)
        L.emu_stream_decode(C.byref(self.s), _p(inp) if inp.size else _p(np.zeros(1, np.uint8)), inp.size, _p(out), out.size,
                            C.byref(p), C.byref(n), int(self.parallel), self.resident, self.seed, C.byref(used))
        self.serial_calls += used.value
        self.seed += 1
        return 0, p.value, n.value

    def has_run_count(self):
        return self.s.run > 0

    def drain_run(self, out):
        ch = self.s.channels
        k = min(self.s.run, out.size // ch)
        px = np.frombuffer(bytes(self.s.prev), dtype=np.uint8)[:ch]
        out[: k * ch] = np.tile(px, k)
        self.s.run -= k
        return 0, k * ch
