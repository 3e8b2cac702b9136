"""ctypes loader for libqoipp_b200.so (the C ABI of include/qoipp_b200.h).

There is no fallback: if the shared library is missing or a symbol declared in the header is not exported,
importing this module raises.  ``declared_symbols()`` parses the header so tests can check the export list.
"""
from __future__ import annotations

import ctypes as C
import os
import re

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SO_PATH = os.environ.get("QOIPP_B200_SO", os.path.join(PKG, "libqoipp_b200.so"))  # override: A/B builds during development
HEADER = os.path.join(ROOT, "include", "qoipp_b200.h")

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)


class Desc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint8), ("colorspace", C.c_uint8)]


class State(C.Structure):
    _fields_ = [("channels", C.c_uint8), ("target", C.c_uint8), ("run", C.c_uint8), ("reserved", C.c_uint8),
                ("prev", C.c_uint8 * 4), ("seen", C.c_uint8 * 256)]


def declared_symbols() -> list[str]:
    text = open(HEADER).read()
    return sorted(set(re.findall(r"\b(qoipp_b200_[a-z0-9_]+)\s*\(", text)))


def _load():
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                          "qoipp_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, pp = C.c_void_p, C.POINTER(C.c_void_p)
    dp = C.POINTER(Desc)
    sig = {
        "qoipp_b200_version": (C.c_int32, []),
        "qoipp_b200_error_string": (C.c_char_p, [C.c_int32]),
        "qoipp_b200_device_count": (C.c_int32, []),
        "qoipp_b200_current_device": (C.c_int32, []),
        "qoipp_b200_ctx_create": (C.c_int32, [C.c_int32, pp]),
        "qoipp_b200_ctx_destroy": (C.c_int32, [vp]),
        "qoipp_b200_count_bytes": (C.c_int32, [dp, u64p]),
        "qoipp_b200_worst_size": (C.c_int32, [dp, u64p]),
        "qoipp_b200_read_header": (C.c_int32, [u8p, C.c_uint64, dp]),
        "qoipp_b200_encode_dev": (C.c_int32, [vp, vp, dp, vp, C.c_uint64, vp]),
        "qoipp_b200_encode_status": (C.c_int32, [vp, vp, u64p, i32p]),
        "qoipp_b200_encode_host": (C.c_int32, [vp, vp, C.c_uint64, dp, vp, C.c_uint64, u64p, i32p]),
        "qoipp_b200_encode_staged": (C.c_int32, [vp, vp, C.c_uint64, dp, u64p]),
        "qoipp_b200_decode_staged": (C.c_int32, [vp, vp, C.c_uint64, C.c_uint8, C.c_int32, dp, u64p]),
        "qoipp_b200_fetch_staged": (C.c_int32, [vp, vp, C.c_uint64]),
        "qoipp_b200_encode_batch_dev": (C.c_int32, [vp, vp, C.c_uint64, C.c_uint32, dp, vp, C.c_uint64, C.c_uint64, vp, vp]),
        "qoipp_b200_encode_batch_host": (C.c_int32, [vp, vp, C.c_uint64, C.c_uint32, dp, vp, C.c_uint64, C.c_uint64, u64p]),
        "qoipp_b200_decode_batch_strided_dev": (C.c_int32, [vp, vp, C.c_uint64, u64p, C.c_uint32, dp, C.c_uint8, vp, C.c_uint64, vp]),
        "qoipp_b200_decode_batch_host": (C.c_int32, [vp, vp, C.c_uint64, u64p, C.c_uint32, dp, C.c_uint8, vp, C.c_uint64]),
        "qoipp_b200_stream_encode_dev": (C.c_int32, [vp, C.c_uint8, vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp]),
        "qoipp_b200_stream_encode_host": (C.c_int32, [vp, C.POINTER(State), vp, C.c_uint64, vp, C.c_uint64, u64p, u64p]),
        "qoipp_b200_decode_dev": (C.c_int32, [vp, vp, C.c_uint64, dp, C.c_uint8, C.c_int32, vp, C.c_uint64, vp]),
        "qoipp_b200_decode_status": (C.c_int32, [vp, vp, i32p]),
        "qoipp_b200_decode_status_batch": (C.c_int32, [vp, vp, i32p, C.c_uint32]),
        "qoipp_b200_decode_host": (C.c_int32, [vp, vp, C.c_uint64, C.c_uint8, C.c_int32, vp, C.c_uint64, dp]),
        "qoipp_b200_decode_batch_dev": (C.c_int32, [vp, vp, u64p, C.c_uint32, dp, C.c_uint8, vp, C.c_uint64, vp]),
        "qoipp_b200_stream_decode_dev": (C.c_int32, [vp, C.c_uint8, vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp]),
        "qoipp_b200_stream_decode_host": (C.c_int32, [vp, C.POINTER(State), vp, C.c_uint64, vp, C.c_uint64, u64p, u64p]),
    }
    old_build = os.environ.get("QOIPP_B200_AB_OLD_BUILD") == "1"  # A/B timing against a library built from an older commit
    for name in declared_symbols():
        if not hasattr(L, name) and not old_build:
            raise ImportError(f"{SO_PATH} does not export {name} (declared in include/qoipp_b200.h)")
    for name, (res, args) in sig.items():
        if old_build and not hasattr(L, name):
            continue
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


lib = _load()
