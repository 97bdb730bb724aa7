"""In-tree nvcc build of librabitq_b200.so (sm_100a only) and of the C++ CLI and HTTP-service twins.

The shared library travels to the GPU box with the repo snapshot; nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librabitq_b200.so")
CLI = os.path.join(HERE, "rabitq_cli")
SERVICE = os.path.join(HERE, "rabitq_service")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # an FMA exists only where the source writes fmaf(): bit-exact parity with the AVX2 reference
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: rabitq_b200 has no non-CUDA build")
    return nvcc


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources() -> list[str]:
    hdr = os.path.join(ROOT, "include", "rabitq_b200.h")
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".cpp", ".hpp"))] + [hdr]


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    if force or _stale(LIB, srcs):
        tmp = LIB + ".tmp"  # built aside and renamed: a repo snapshot never sees a half-written library
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", tmp, os.path.join(CSRC, "rabitq_capi.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd, cwd=ROOT)
        os.replace(tmp, LIB)
    cli_src = os.path.join(CSRC, "cli.cpp")
    if os.path.exists(cli_src) and (force or _stale(CLI, srcs + [LIB])):
        gxx = shutil.which("g++") or "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-o", CLI, cli_src, "-I", os.path.join(ROOT, "include"),
                               "-L", HERE, "-lrabitq_b200", "-Wl,-rpath,$ORIGIN"], cwd=ROOT)
    svc_src = os.path.join(CSRC, "service.cpp")
    if os.path.exists(svc_src) and (force or _stale(SERVICE, srcs + [LIB])):
        gxx = shutil.which("g++") or "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-pthread", "-o", SERVICE, svc_src, "-I", os.path.join(ROOT, "include"),
                               "-L", HERE, "-lrabitq_b200", "-Wl,-rpath,$ORIGIN"], cwd=ROOT)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
