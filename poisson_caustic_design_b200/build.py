"""Builds the native pieces in-tree (no JIT cache):

* ``libpcd_b200.so``   -- CUDA kernels + C ABI (include/pcd.h), nvcc, sm_100a only
* ``libpcd_host.so``   -- host-only C++ (PNG ingest, OBJ/JSON/SVG writers, CLI argument handling): no CUDA
* ``caustic_design``   -- the CLI executable (reference main.cpp flags), linked against both

    python -m poisson_caustic_design_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
INCLUDE = os.path.join(ROOT, "include")

LIB_CUDA = os.path.join(PKG, "libpcd_b200.so")
LIB_HOST = os.path.join(PKG, "libpcd_host.so")
CLI = os.path.join(PKG, "caustic_design")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # -fmad=false: every a*b+c keeps the reference's two roundings, so stage kernels fed the oracle's
    # inputs agree with it to the last bit wherever the summation order is the same (parity)
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(d: str, exts: tuple[str, ...]) -> list[str]:
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources(CSRC, (".cu",))
    deps = srcs + _sources(CSRC, (".h", ".cuh")) + [os.path.join(INCLUDE, "pcd.h")]
    if not force and _newer(LIB_CUDA, deps):
        return LIB_CUDA
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-o", LIB_CUDA] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_CUDA


def build_host(force: bool = False, verbose: bool = False) -> tuple[str, str]:
    srcs = [s for s in _sources(HOST, (".cpp",)) if not s.endswith("main.cpp")]
    deps = _sources(HOST, (".cpp", ".h")) + [os.path.join(INCLUDE, "pcd.h")]
    if not srcs:
        return LIB_HOST, CLI
    cxx = shutil.which("g++") or "g++"
    common = [cxx, "-std=c++17", "-O2", "-fPIC", "-Wall", "-I", INCLUDE, "-I", HOST]
    if force or not _newer(LIB_HOST, deps + [LIB_CUDA]):
        cmd = common + ["-shared", "-o", LIB_HOST] + srcs + ["-L", PKG, "-lpcd_b200", "-lz", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    main_cpp = os.path.join(HOST, "main.cpp")
    if os.path.exists(main_cpp) and (force or not _newer(CLI, deps + [LIB_HOST])):
        cmd = common + ["-o", CLI, main_cpp, "-L", PKG, "-lpcd_host", "-lpcd_b200", "-lz", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB_HOST, CLI


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    print("built:", LIB_CUDA)
