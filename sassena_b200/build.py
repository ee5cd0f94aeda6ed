"""Builds libsassena_b200.so (CUDA kernels + C-ABI + C++ host layer) in-tree with nvcc for sm_100a.

Usage: python -m sassena_b200.build [--force]
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsassena_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--shared",
]


def sources():
    srcs = sorted(glob.glob(os.path.join(CSRC, "kernels", "*.cu")))
    srcs += sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    srcs += sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    return srcs


def _deps():
    d = sources()
    for pat in ("kernels/*.hpp", "kernels/*.cuh", "host/*.hpp", "*.hpp"):
        d += glob.glob(os.path.join(CSRC, pat))
    d.append(os.path.join(HERE, "..", "include", "sassena_b200.h"))
    d.append(os.path.join(HERE, "..", "include", "sassena_host.h"))
    return [p for p in d if os.path.exists(p)]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper compiler; nvcc should use the system g++
    cmd = [nvcc, "-ccbin", "/usr/bin/g++"] + NVCC_FLAGS + ["-I", os.path.join(HERE, "..", "include"), "-I", CSRC,
                                                           "-o", LIB] + sources()
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
