"""In-tree build of libpdsat_b200.so (sm_100a only) with plain nvcc.

The .so lands next to this file (git-ignored, but it travels with the gpurun snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpdsat_b200.so")
SOURCES = ["pdsat_cuda.cu"]
HEADERS = ["kernels.cuh", "host_db.hpp", "mt_jump.hpp", "handoff.hpp", os.path.join("..", "host", "cdcl.hpp"),
           os.path.join("..", "..", "include", "pdsat_cuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
]


def find_nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA path cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


HOST_DIR = os.path.join(HERE, "host")
HOST_BIN = os.path.join(HERE, "pd-sat-b200")
HOST_SOURCES = ["pd_sat_main.cpp", "dimacs.hpp", "predict.hpp", "search.hpp"]


def build_host(force: bool = False, verbose: bool = False) -> str:
    """C++ host driver (reference-style CLI) linked against the C ABI only."""
    stale = force or not os.path.exists(HOST_BIN) or any(
        os.path.getmtime(os.path.join(HOST_DIR, f)) > os.path.getmtime(HOST_BIN) for f in HOST_SOURCES) or \
        os.path.getmtime(LIB) > os.path.getmtime(HOST_BIN)
    if not stale:
        return HOST_BIN
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-o", HOST_BIN, os.path.join(HOST_DIR, "pd_sat_main.cpp"),
           "-L" + HERE, "-lpdsat_b200", "-Wl,-rpath,$ORIGIN"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return HOST_BIN


def build(force: bool = False, verbose: bool = False, extra=()) -> str:
    if force or needs_build():
        cmd = [find_nvcc()] + NVCC_FLAGS + list(extra) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    build_host(force, verbose)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True, extra=sys.argv[1:])
