"""Build libhammock_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The shared library is the product: it exports the C ABI of include/hammock_b200.h.
nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhammock_b200.so")
SOURCES = [os.path.join(CSRC, "hmk_engine.cu")]
HEADERS = [os.path.join(CSRC, f) for f in ("hmk_common.h", "hmk_resolve.h", "hmk_kernels.cuh")] + [
    os.path.join(ROOT, "include", "hammock_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libhammock_b200.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """defines/out: instrumented variants for profiling sessions, e.g. defines=("HMK_DEBUG", "HMK_RESOLVE_TIMING"),
    out=".../libhammock_b200_timing.so" (load it with HMK_LIB=...); the product is the plain build."""
    if out == LIB and not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-pthread", "-shared", "-o", out] + [f"-D{d}" for d in defines] + SOURCES
    # the image exports CC/CXX wrappers that lack parts of the toolchain; use the system g++
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return out


HOST_SRC = os.path.join(HERE, "host_cpp", "hammock_greedy.cpp")
HOST_HDR = os.path.join(HERE, "host_cpp", "hammock_host.hpp")
HOST_BIN = os.path.join(HERE, "hammock_greedy")


def build_host(force: bool = False) -> str:
    """The native `greedy`-mode driver (C++ host side above the C ABI); links libhammock_b200.so."""
    build(force=False)
    if not force and os.path.exists(HOST_BIN) and all(
            os.path.getmtime(HOST_BIN) >= os.path.getmtime(f) for f in (HOST_SRC, HOST_HDR, LIB)):
        return HOST_BIN
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-Wall", "-o", HOST_BIN, HOST_SRC, "-L", HERE, "-lhammock_b200",
                           "-Wl,-rpath,$ORIGIN"])
    return HOST_BIN


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
    print(build_host(force="--force" in sys.argv))
