"""Builds the CUDA shared library in-tree: spruce_b200/lib/libspruce_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  -fmad=false is REQUIRED: the kernels reproduce the reference's individually
rounded FP64 arithmetic, and the only fused operations are the explicit fma() calls of the exact division
(csrc/exact_math.cuh).
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc"
LIB = PKG / "lib" / "libspruce_b200.so"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550"]


def sources():
    return [SRC / "capi.cu"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(SRC.glob("*.cu")) + list(SRC.glob("*.cuh")) + [PKG.parent / "include" / "spruce_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB), *map(str, sources())]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)   # let nvcc pick the system host compiler
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout.decode())
    if r.returncode:
        raise RuntimeError("nvcc failed (%d)" % r.returncode)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
