"""Builds libfheb200.so (the C-ABI library of include/fheb200.h) in-tree with nvcc for sm_100a.

    python node-fhe-accelerate_b200/build.py [--force] [--verbose]

Every csrc/*.cu is compiled to build/obj/*.o (in parallel) and linked into
node-fhe-accelerate_b200/libfheb200.so.  nvcc cross-compiles without a GPU.  Objects are rebuilt
when their source or any header is newer.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libfheb200.so")
# experiments: FHEB_BUILD_TAG=<tag> FHEB_BUILD_DEFINES="-DX=1 ..." builds libfheb200_<tag>.so from separate objects
_TAG = os.environ.get("FHEB_BUILD_TAG", "")
if _TAG:
    OBJ = os.path.join(ROOT, "build", "obj_" + _TAG)
    LIB = os.path.join(HERE, f"libfheb200_{_TAG}.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-std=c++17", "-O3", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I", os.path.join(ROOT, "include"),
] + os.environ.get("FHEB_BUILD_DEFINES", "").split()


def _newest_header() -> float:
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.hpp")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, obj: str, verbose: bool) -> str:
    cmd = [NVCC, *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        with open(obj + ".log", "w") as f:
            f.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdr = _newest_header()
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr):
            jobs.append((s, o))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for f in [ex.submit(_compile, s, o, verbose) for s, o in jobs]:
                f.result()
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
