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
]
OBJDIR = os.path.join(HERE, "build")  # git-ignored; one object per source so that an edit recompiles one file


def sources():
    srcs = sorted(glob.glob(os.path.join(CSRC, "kernels", "*.cu")))
    srcs += sorted(glob.glob(os.path.join(CSRC, "*.cu"))) + sorted(glob.glob(os.path.join(CSRC, "*.cpp")))
    srcs += sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    return srcs


def _headers():
    d = []
    for pat in ("kernels/*.hpp", "kernels/*.cuh", "host/*.hpp", "*.hpp"):
        d += glob.glob(os.path.join(CSRC, pat))
    d.append(os.path.join(HERE, "..", "include", "sassena_b200.h"))
    d.append(os.path.join(HERE, "..", "include", "sassena_host.h"))
    d.append(os.path.abspath(__file__))
    return [p for p in d if os.path.exists(p)]


def _obj(src):
    return os.path.join(OBJDIR, os.path.relpath(src, CSRC).replace(os.sep, "_") + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in deps)


def needs_build() -> bool:
    return _stale(LIB, sources() + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = _headers()
    # the image exports CC/CXX pointing at a wrapper compiler; nvcc should use the system g++
    base = [nvcc, "-ccbin", "/usr/bin/g++"] + NVCC_FLAGS + ["-I", os.path.join(HERE, "..", "include"), "-I", CSRC]

    def compile_one(src):
        obj = _obj(src)
        if force or _stale(obj, [src] + hdrs):
            cmd = base + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd, env=env)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = base + ["--shared", "-o", LIB] + objs + ["-ldl"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
