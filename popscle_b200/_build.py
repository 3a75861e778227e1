"""In-tree build of libpopscle_b200.so (nvcc, sm_100a only) and of the C++ CLI host."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "popscle_b200.cu")
LIB = os.path.join(HERE, "libpopscle_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-strict-aliasing", "-diag-suppress", "550"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def cuda_sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh", ".inl"))] + [
        os.path.join(ROOT, "include", "popscle_b200.h")]


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and _newer(LIB, cuda_sources()):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    subprocess.check_call(cmd, cwd=ROOT)
    return LIB


def host_sources():
    d = os.path.join(HERE, "host")
    if not os.path.isdir(d):
        return []
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cpp", ".h"))]


def build_host(force: bool = False) -> str | None:
    """The `popscle` CLI host (C++17, zlib), linked against libpopscle_b200.so."""
    srcs = [s for s in host_sources() if s.endswith(".cpp")]
    if not srcs:
        return None
    exe = os.path.join(HERE, "popscle")
    if not force and _newer(exe, host_sources() + [LIB]):
        return exe
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe] + srcs + [
        "-L", HERE, "-lpopscle_b200", "-lz", "-pthread", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd, cwd=ROOT)
    return exe
