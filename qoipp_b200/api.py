"""Python driver over the C ABI -- mirrors the reference's one-shot and stream interfaces
(include/qoipp/simple.hpp, stream.hpp) closely enough that the parity tests read like the reference's own.

Host data are numpy uint8 arrays; device data are anything with a ``data_ptr()`` (torch tensors).
Errors are the integer qoipp::Error values (0 = ok); nothing here computes pixels or bytes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Desc, State, lib

HEADER_SIZE, MARKER_SIZE = 14, 8


class QoiError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"qoipp_b200 error {code}: {lib.qoipp_b200_error_string(code).decode()}")
        self.code = code


def _np_ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _dev_ptr(t):
    return C.c_void_p(t.data_ptr())


def worst_size(w, h, ch, cs=0) -> int:
    out = C.c_uint64(0)
    e = lib.qoipp_b200_worst_size(C.byref(Desc(w, h, ch, cs)), C.byref(out))
    if e:
        raise QoiError(e)
    return out.value


def read_header(qoi: np.ndarray):
    d = Desc()
    q = np.ascontiguousarray(qoi, dtype=np.uint8)
    e = lib.qoipp_b200_read_header(q.ctypes.data_as(C.POINTER(C.c_uint8)), q.size, C.byref(d))
    return e, (d.width, d.height, d.channels, d.colorspace)


class Context:
    """One per driving thread (qoipp_b200_ctx)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        e = lib.qoipp_b200_ctx_create(device, C.byref(self._h))
        if e:
            raise QoiError(e)
        self.device = device

    def close(self):
        if self._h:
            lib.qoipp_b200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- qoipp::encode_into(ByteSpan, ByteCSpan, Desc)
    def encode_into(self, raw: np.ndarray, w, h, ch, cs=0, cap=None):
        """-> (err, out[:cap], written, complete); bytes beyond `written` are left as 0xAA"""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        if cap is None:
            cap = (ch + 1) * w * h + 22
        out = np.full(max(cap, 1), 0xAA, dtype=np.uint8)
        written, complete = C.c_uint64(0), C.c_int32(0)
        e = lib.qoipp_b200_encode_host(self._h, _np_ptr(raw), raw.size, C.byref(Desc(w, h, ch, cs)), _np_ptr(out), cap,
                                       C.byref(written), C.byref(complete))
        return e, out[:cap], written.value, bool(complete.value)

    def encode(self, raw, w, h, ch, cs=0) -> np.ndarray:
        e, out, n, ok = self.encode_into(raw, w, h, ch, cs)
        if e:
            raise QoiError(e)
        assert ok
        return out[:n].copy()

    # ---- device-pointer (timed) path
    def encode_dev(self, d_raw, w, h, ch, cs, d_out, cap, stream=0):
        e = lib.qoipp_b200_encode_dev(self._h, _dev_ptr(d_raw), C.byref(Desc(w, h, ch, cs)), _dev_ptr(d_out), cap, C.c_void_p(stream))
        if e:
            raise QoiError(e)

    def encode_status(self, stream=0):
        written, complete = C.c_uint64(0), C.c_int32(0)
        e = lib.qoipp_b200_encode_status(self._h, C.c_void_p(stream), C.byref(written), C.byref(complete))
        if e:
            raise QoiError(e)
        return written.value, bool(complete.value)

    def encode_batch_dev(self, d_raw, raw_stride, n_images, w, h, ch, cs, d_out, out_stride, out_cap, d_written=None, stream=0):
        e = lib.qoipp_b200_encode_batch_dev(self._h, _dev_ptr(d_raw), raw_stride, n_images, C.byref(Desc(w, h, ch, cs)), _dev_ptr(d_out),
                                            out_stride, out_cap, _dev_ptr(d_written) if d_written is not None else None,
                                            C.c_void_p(stream))
        if e:
            raise QoiError(e)

    # ---- qoipp::decode_into(ByteSpan, ByteCSpan, target, flip)
    def decode_into(self, qoi: np.ndarray, target=0, flip=False, cap=None):
        """-> (err, pixels or None, (w, h, ch, cs))"""
        qoi = np.ascontiguousarray(qoi, dtype=np.uint8)
        if cap is None:
            e, hd = read_header(qoi)
            cap = hd[0] * hd[1] * max(target or hd[2], hd[2]) if e == 0 else 16
        out = np.full(max(cap, 1), 0xAA, dtype=np.uint8)
        d = Desc()
        e = lib.qoipp_b200_decode_host(self._h, _np_ptr(qoi), qoi.size, target, int(flip), _np_ptr(out), cap, C.byref(d))
        if e:
            return e, None, None
        return 0, out[: d.width * d.height * d.channels].copy(), (d.width, d.height, d.channels, d.colorspace)

    def decode(self, qoi, target=0, flip=False) -> np.ndarray:
        e, px, _ = self.decode_into(qoi, target, flip)
        if e:
            raise QoiError(e)
        return px

    def decode_dev(self, d_qoi, size, w, h, ch, cs, target, flip, d_out, cap, stream=0):
        e = lib.qoipp_b200_decode_dev(self._h, _dev_ptr(d_qoi), size, C.byref(Desc(w, h, ch, cs)), target, int(flip), _dev_ptr(d_out), cap,
                                      C.c_void_p(stream))
        if e:
            raise QoiError(e)

    def decode_status(self, stream=0) -> int:
        path = C.c_int32(0)
        e = lib.qoipp_b200_decode_status(self._h, C.c_void_p(stream), C.byref(path))
        if e:
            raise QoiError(e)
        return path.value

    def decode_status_batch(self, n_images: int, stream=0) -> np.ndarray:
        paths = np.zeros(n_images, dtype=np.int32)
        e = lib.qoipp_b200_decode_status_batch(self._h, C.c_void_p(stream), paths.ctypes.data_as(C.POINTER(C.c_int32)), n_images)
        if e:
            raise QoiError(e)
        return paths

    def decode_batch_dev(self, d_qoi, offsets: np.ndarray, w, h, ch, cs, target, d_out, out_stride, stream=0):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        e = lib.qoipp_b200_decode_batch_dev(self._h, _dev_ptr(d_qoi), offsets.ctypes.data_as(C.POINTER(C.c_uint64)), offsets.size - 1,
                                            C.byref(Desc(w, h, ch, cs)), target, _dev_ptr(d_out), out_stride, C.c_void_p(stream))
        if e:
            raise QoiError(e)


class StreamEncoder:
    """qoipp::StreamEncoder (include/qoipp/stream.hpp:23-116): initialize / encode / finalize / reset.
    The header and end-marker writes are host byte copies; encode() is the resumable kernel."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.s = State()
        self.reset()

    def reset(self):
        C.memset(C.byref(self.s), 0, C.sizeof(self.s))
        self.s.prev[3] = 255

    def is_initialized(self):
        return self.s.channels != 0

    def has_run_count(self):
        return self.s.run > 0

    def initialize(self, out: np.ndarray, w, h, ch, cs=0):  # stream.cpp:113-136
        if self.s.channels:
            return 9, 0
        if out.size == 0:
            return 1, 0
        if out.size < HEADER_SIZE:
            return 2, 0
        n = C.c_uint64(0)
        e = lib.qoipp_b200_count_bytes(C.byref(Desc(w, h, ch, cs)), C.byref(n))
        if e:
            return e, 0
        out[:HEADER_SIZE] = np.frombuffer(b"qoif" + w.to_bytes(4, "big") + h.to_bytes(4, "big") + bytes([ch, cs]), dtype=np.uint8)
        self.s.channels = ch
        return 0, HEADER_SIZE

    def encode(self, out: np.ndarray, inp: np.ndarray):
        inp = np.ascontiguousarray(inp)
        p, n = C.c_uint64(0), C.c_uint64(0)
        e = lib.qoipp_b200_stream_encode_host(self.ctx._h, C.byref(self.s), _np_ptr(inp), inp.size, _np_ptr(out), out.size, C.byref(p),
                                              C.byref(n))
        return e, p.value, n.value

    def finalize(self, out: np.ndarray):  # stream.cpp:241-267
        if not self.s.channels:
            return 8, 0
        if out.size == 0:
            return 1, 0
        need = MARKER_SIZE + (self.s.run > 0)
        if out.size < need:
            return 2, 0
        k = 0
        if self.s.run:
            out[0] = 0xC0 | (self.s.run - 1)
            k = 1
        out[k: k + MARKER_SIZE] = [0, 0, 0, 0, 0, 0, 0, 1]
        self.reset()
        return 0, need


class StreamDecoder:
    """qoipp::StreamDecoder (include/qoipp/stream.hpp:133-244)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.s = State()
        self.reset()

    def reset(self):
        C.memset(C.byref(self.s), 0, C.sizeof(self.s))
        self.s.prev[3] = 255

    def is_initialized(self):
        return self.s.channels != 0

    def has_run_count(self):
        return self.s.run > 0

    def initialize(self, inp: np.ndarray, target=0):  # stream.cpp:290-310
        if self.s.channels:
            return 9, None
        e, d = read_header(inp)
        if e:
            return e, None
        n = C.c_uint64(0)
        e = lib.qoipp_b200_count_bytes(C.byref(Desc(*d)), C.byref(n))
        if e:
            return e, None
        ch = target or d[2]
        self.s.channels = self.s.target = ch
        self.s.seen[53 * 4 + 3] = 255  # slot of the start pixel {0,0,0,255}
        return 0, (d[0], d[1], ch, d[3])

    def decode(self, out: np.ndarray, inp: np.ndarray):
        inp = np.ascontiguousarray(inp)
        p, n = C.c_uint64(0), C.c_uint64(0)
        e = lib.qoipp_b200_stream_decode_host(self.ctx._h, C.byref(self.s), _np_ptr(inp), inp.size, _np_ptr(out), out.size, C.byref(p),
                                              C.byref(n))
        return e, p.value, n.value

    def drain_run(self, out: np.ndarray):  # stream.cpp:426-447: replicate the pending pixel, no kernel needed
        if not self.s.channels:
            return 8, 0
        if out.size == 0:
            return 1, 0
        ch = self.s.channels
        k = min(self.s.run, out.size // ch)
        px = np.frombuffer(bytes(self.s.prev), dtype=np.uint8)[:ch]
        out[: k * ch] = np.tile(px, k)
        self.s.run -= k
        return 0, k * ch
