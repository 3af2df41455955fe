"""In-tree build of the native pieces (sm_100a only; nvcc cross-compiles without a GPU).

  libppcsr_b200.so         CUDA kernels + the C-ABI of include/ppcsr_b200.h
  host/parallel-packed-csr the reference-compatible CLI (PCSR / PPPCSR / ThreadPool* classes + main.cpp flags)
  host/test_host           the reference's unit tests restated against those classes

The built files are git-ignored but travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libppcsr_b200.so")
CLI = os.path.join(PKG, "host", "parallel-packed-csr")
HOST_TEST = os.path.join(PKG, "host", "test_host")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(sub: str, exts=(".cu", ".cuh", ".h", ".cpp")) -> list[str]:
    d = os.path.join(PKG, sub)
    out = [os.path.join(ROOT, "include", "ppcsr_b200.h")]
    for f in sorted(os.listdir(d)):
        if f.endswith(exts):
            out.append(os.path.join(d, f))
    return out


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return exe


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources("csrc")
    if not force and _newer(LIB, srcs):
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(PKG, "csrc", "capi.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB


def build_host(force: bool = False) -> list[str]:
    """C++ host mirror of the reference's classes + CLI, linked against libppcsr_b200.so."""
    hdir = os.path.join(PKG, "host")
    if not os.path.isdir(hdir) or not os.path.exists(os.path.join(hdir, "main.cpp")):
        return []
    build_library(force=False)
    srcs = _sources("host")
    common = [os.path.join(hdir, f) for f in ("PCSR.cpp", "PPPCSR.cpp", "thread_pool.cpp", "thread_pool_pppcsr.cpp")]
    flags = ["-std=c++17", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"), "-I", hdir]
    link = ["-L", PKG, "-lppcsr_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/.."]
    out = []
    for target, main in ((CLI, "main.cpp"), (HOST_TEST, "test_host.cpp")):
        if not os.path.exists(os.path.join(hdir, main)):
            continue
        if force or not _newer(target, srcs + [LIB]):
            subprocess.run(["g++", *flags, "-o", target, os.path.join(hdir, main), *common, *link], check=True)
        out.append(target)
    return out


def build_all(force: bool = False) -> None:
    build_library(force=force)
    build_host(force=force)


if __name__ == "__main__":
    import sys

    build_all(force="--force" in sys.argv)
    print(LIB)
