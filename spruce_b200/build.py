"""Builds the CUDA shared library in-tree: spruce_b200/lib/libspruce_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  -fmad=false is REQUIRED for capi.cu: the kernels reproduce the reference's individually
rounded FP64 arithmetic, and the only fused operations are the explicit fma() calls of the exact division
(csrc/exact_math.cuh).  stage_relaxed.cu is the one unit compiled with contraction: the opt-in relaxed stage kernel.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc"
LIB = PKG / "lib" / "libspruce_b200.so"
COMMON_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-diag-suppress", "550"]
# translation units and their floating-point contraction: everything is exact (-fmad=false) except the opt-in relaxed stage kernel
UNITS = [("capi.cu", "-fmad=false"), ("stage_relaxed.cu", "-fmad=true")]
OBJ = PKG / "lib" / "obj"


def sources():
    return [SRC / name for name, _ in UNITS]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(SRC.glob("*.cu")) + list(SRC.glob("*.cuh")) + list(SRC.glob("*.hpp")) + [PKG.parent / "include" / "spruce_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if a source is newer than it.  Several processes may call this at once (the ranks of a torchrun launch): a file lock lets one of them
    compile while the others wait and then find the library up to date."""
    if not force and not needs_build():
        return LIB
    import fcntl
    LIB.parent.mkdir(parents=True, exist_ok=True)
    with open(LIB.parent / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():          # another process built it while this one waited
                return LIB
            return _compile(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _compile(verbose: bool) -> Path:
    LIB.parent.mkdir(parents=True, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)   # let nvcc pick the system host compiler
    OBJ.mkdir(parents=True, exist_ok=True)

    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout.decode())
        if r.returncode:
            raise RuntimeError("nvcc failed (%d)" % r.returncode)

    objs = []
    procs = []
    for name, fmad in UNITS:                     # the units compile concurrently
        obj = OBJ / (Path(name).stem + ".o")
        cmd = [nvcc, *COMMON_FLAGS, *os.environ.get("SPRUCE_NVCC_FLAGS", "").split(), fmad, "-c", "-o", str(obj), str(SRC / name)]   # SPRUCE_NVCC_FLAGS: experiments (-D...)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)))
        objs.append(obj)
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out.decode())
        if pr.returncode:
            raise RuntimeError("nvcc failed (%d): %s" % (pr.returncode, " ".join(cmd)))
    run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB), *map(str, objs)])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
