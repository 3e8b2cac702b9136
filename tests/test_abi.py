"""CPU: the C-ABI library loads and exports every symbol include/qoipp_b200.h declares; host-only entry points
(no device work) behave like the reference's helpers; device entry points fail loudly without a GPU."""
import ctypes as C
import subprocess
import os

import numpy as np
import pytest

from oracle.pyoracle import Oracle
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g

    g.build()
    from qoipp_b200 import _lib

    return _lib


def test_exports_every_declared_symbol(L):
    names = L.declared_symbols()
    assert len(names) >= 18
    out = subprocess.run(["nm", "-D", "--defined-only", L.SO_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for n in names:
        assert n in exported, n
    assert L.lib.qoipp_b200_version() == 100


def test_no_torch_or_cuda_types_in_the_header():
    import re

    text = open(os.path.join(ROOT, "include", "qoipp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # signatures only: comments may name what is excluded
    for word in ("torch", "at::", "cudaStream_t", "#include <cuda"):
        assert word not in text


def test_host_helpers_match_the_oracle(L):
    f = H.fixtures()
    for ch in (3, 4):
        d = L.Desc()
        q = f[ch]["qoi"]
        assert L.lib.qoipp_b200_read_header(q.ctypes.data_as(C.POINTER(C.c_uint8)), q.size, C.byref(d)) == 0
        assert (d.width, d.height, d.channels, d.colorspace) == f[ch]["desc"]
    for bad, want in ((np.zeros(0, np.uint8), 1), (np.array([1, 2, 3], np.uint8), 2), (np.arange(14, dtype=np.uint8), 4)):
        d = L.Desc()
        assert L.lib.qoipp_b200_read_header(bad.ctypes.data_as(C.POINTER(C.c_uint8)), bad.size, C.byref(d)) == want == Oracle.read_header(bad)[0]
    for (w, h, ch, cs) in ((7, 9, 3, 0), (16384, 16384, 4, 1), (0, 1, 3, 0), (1, 1, 5, 0), (1, 1, 4, 2), (0xFFFFFFFF, 0xFFFFFFFF, 4, 0)):
        out = C.c_uint64(0)
        e = L.lib.qoipp_b200_worst_size(C.byref(L.Desc(w, h, ch, cs)), C.byref(out))
        eo, vo = Oracle.worst_size(w, h, ch, cs)
        assert e == eo and (e != 0 or out.value == vo)
    assert L.lib.qoipp_b200_error_string(7) == b"Buffer does not have enough space"


def test_device_calls_fail_loudly_without_a_gpu(L):
    if L.lib.qoipp_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    e = L.lib.qoipp_b200_ctx_create(0, C.byref(h))
    assert e < 0 and not h.value  # a negative cudaError_t, never a silent CPU path
    from qoipp_b200 import api

    with pytest.raises(api.QoiError):
        api.Context(0)
