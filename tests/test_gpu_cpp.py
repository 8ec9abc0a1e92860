"""Builds and runs tests/cpp/test_cabi_classes.cpp on the B200: the reference-named C++ classes of
include/fheb200.hpp (over the C ABI) against the CPU oracle, in the style of the reference's own
stand-alone C++ tests."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_cpp_classes_over_cabi():
    lib_dir = os.path.join(ROOT, "node-fhe-accelerate_b200")
    exe = os.path.join(ROOT, "build", "test_cabi_classes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_cabi_classes.cpp"),
                           os.path.join(ROOT, "oracle", "fhe_oracle.c"), "-L" + lib_dir, "-lfheb200", "-Wl,-rpath," + lib_dir])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "CPP CABI TESTS OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_node_addon_bulk_entry_points_under_the_mock_host():
    """The Node-API addon (addon/fheb_addon.cc) built as a .node shared object and driven through a mock Node-API host on
    the GPU: initialize, detectHardware, setDevices, NttProcessor, tallyVotes, modAddBatch against the oracle."""
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fheb200
    from test_cabi_cpu import _build_addon_and_host

    out = os.path.join(ROOT, "build")
    os.makedirs(out, exist_ok=True)
    host, node = _build_addon_and_host(fheb200, out)
    r = subprocess.run([host, node, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "NAPI HOST GPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
