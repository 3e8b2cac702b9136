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


@pytest.mark.gpu
def test_qoiconv_round_trip_on_disk(tmp_path):
    """examples/qoiconv (02_conv.cpp analogue): raw -> .qoi -> raw through the file overloads of the C++ API."""
    from oracle.pyoracle import Oracle
    from qoipp_b200 import synth

    subprocess.run(["make", "-C", os.path.join(ROOT, "qoipp_b200", "csrc", "cxx")], check=True, stdout=subprocess.DEVNULL)
    tool = os.path.join(ROOT, "examples", "qoiconv")
    for ch in (3, 4):
        w, h = 300, 217
        raw = synth.generate("photo", w, h, ch)
        raw.tofile(tmp_path / "in.raw")
        r = subprocess.run([tool, "encode", str(tmp_path / "in.raw"), str(w), str(h), str(ch), str(tmp_path / "a.qoi")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert np.array_equal(np.fromfile(tmp_path / "a.qoi", dtype=np.uint8), Oracle.encode(raw, w, h, ch))
        r = subprocess.run([tool, "info", str(tmp_path / "a.qoi")], capture_output=True, text=True)
        assert r.returncode == 0 and f"{w} x {h}, {ch} channels" in r.stdout, r.stdout + r.stderr
        r = subprocess.run([tool, "decode", str(tmp_path / "a.qoi"), str(tmp_path / "out.raw")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert np.array_equal(np.fromfile(tmp_path / "out.raw", dtype=np.uint8), raw)
    r = subprocess.run([tool, "info", str(tmp_path / "in.raw")], capture_output=True, text=True)
    assert r.returncode == 1 and "Not a QOI file" in r.stderr


@pytest.mark.gpu
def test_thread_per_gpu_example_serves_every_gpu(tmp_path):
    """examples/batch_multi_gpu: one host thread per GPU selects its device with qoipp::b200::set_device and runs the ordinary
    qoipp::encode / qoipp::decode; image k of B on GPU floor(k*G/B) (all GPUs of the box; one is enough for the test to mean
    something: the example checks qoipp::b200::device() inside every call)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "qoipp_b200", "csrc", "cxx")], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(ROOT, "examples", "batch_multi_gpu"), "24", "256", "192", "4"], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout.splitlines()[-1]


@pytest.mark.gpu
def test_several_host_threads_drive_one_gpu(tmp_path):
    """The free functions are re-entrant (SURVEY 8(b) threading): six host threads encode and decode at the same time on GPU 0,
    each through its own context and stream; every result is checked."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "qoipp_b200", "csrc", "cxx")], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(ROOT, "examples", "batch_multi_gpu"), "96", "640", "360", "4", "1", "6"], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "96 images on 1 GPUs" in r.stdout and "0 failures" in r.stdout.splitlines()[-1]
