"""ctypes door into tests/emu/libqb_emu.so: the product kernels stepped by the CPU SIMT emulator.
TEST INFRASTRUCTURE ONLY -- lets kernel logic be checked against the oracle on hosts without a GPU."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
SO = os.path.join(HERE, "libqb_emu.so")
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
