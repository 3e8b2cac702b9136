"""qoipp_b200 -- B200-native QOI codec behind mrizaln/qoipp's API (hot path only).

The product is ``csrc/`` (sm_100a CUDA kernels + the C ABI declared in ``include/qoipp_b200.h``) and
the C++20 ``qoipp::`` host API above it.  The Python modules here are thin drivers for tests and
``bench.py``: ``_lib`` loads the shared library through ctypes, ``synth`` makes deterministic inputs.
There is no CPU fallback: importing ``_lib`` without the built library raises.
"""
__version__ = "0.1.0"
