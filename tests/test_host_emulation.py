"""Runs the CUDA pass bodies on the CPU (tests/host_emulation/*.cpp) against the C oracle.

The device code in node-fhe-accelerate_b200/csrc/{ntt_core,boot_core}.cuh is written as
host/device functions of (thread id, thread count); the emulators execute the threads of each
barrier-separated phase sequentially.  This validates index math, twiddle ordering, swizzling,
lazy-range bookkeeping and the blind-rotation algebra without a GPU.  It is a check of the
product's source, not a CPU execution path of the product: nothing here is shipped or timed.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_pipeline_schedule():
    """Pure host logic of the host-buffer pipeline (chunk schedule), compiled with g++ against the CUDA headers only."""
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "check_pipeline")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", cuda_inc, "-o", exe,
                           os.path.join(ROOT, "tests", "host_emulation", "check_pipeline.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PIPELINE SCHEDULE OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name,ok", [("emulate_ntt", "HOST EMULATION OK"), ("emulate_boot", "BOOT EMULATION OK")])
def test_emulation(name, ok):
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, name)
    src = os.path.join(ROOT, "tests", "host_emulation", name + ".cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, src, os.path.join(ROOT, "oracle", "fhe_oracle.c")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and ok in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
