"""GPU (-m gpu): builds and runs tests/cxx/test_api, the C++ replay of the reference's own test files against the
B200-backed qoipp:: API (include/qoipp/qoipp.hpp -> libqoipp.so -> libqoipp_b200.so)."""
import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cxx_api_replays_reference_tests(tmp_path):
    subprocess.run(["make", "-C", os.path.join(ROOT, "qoipp_b200", "csrc", "cxx")], check=True, stdout=subprocess.DEVNULL)
    z = np.load(os.path.join(H.GOLDEN, "fixtures.npz"))
    for k in z.keys():
        z[k].tofile(tmp_path / f"{k}.bin")
    r = subprocess.run([os.path.join(ROOT, "tests", "cxx", "test_api"), str(tmp_path), "1"], capture_output=True, text=True, timeout=900)
    print(r.stdout[-2000:], r.stderr[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert " 0 failed" in r.stdout
