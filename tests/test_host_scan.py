"""spruce_b200/csrc/host_scan.hpp: the zero-plane test of spruce_grid_upload (which planes are identically +-0 decides which instance of the stage kernel runs, so a
wrong answer would be a wrong result).  Against numpy on planes below and above the multi-thread threshold: all zero, negative zeros, a single non-zero value at
every kind of position (first, last, each thread's chunk edges), NaN, a denormal; every thread count; timing of the 4096^2 case for the record."""
import ctypes as C
import subprocess
import time
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "host_scan_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libhost_scan_check.so"


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC, ROOT / "spruce_b200" / "csrc" / "host_scan.hpp"]
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O3", "-shared", "-fPIC", "-pthread", "-o", str(LIB), str(SRC)], check=True)
    L = C.CDLL(str(LIB))
    L.plane_nonzero.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
    return L


def ask(lib, a, threads=8):
    return bool(lib.plane_nonzero(a.ctypes.data, a.size, threads))


@pytest.mark.parametrize("n", [1, 5, 4096, 4097, 300 * 211, (1 << 20) - 1, 1 << 20, (1 << 21) + 12345])
def test_zero_and_single_nonzero_positions(lib, n):
    a = np.zeros(n)
    for threads in (1, 2, 3, 8):
        assert not ask(lib, a, threads)
        a[::3] = -0.0                                           # negative zeros are zeros
        assert not ask(lib, a, threads)
        chunk = (n + threads - 1) // threads
        for pos in {0, n - 1, n // 2, min(chunk, n - 1), max(chunk - 1, 0), min(2 * chunk, n - 1), min(4095, n - 1), min(4096, n - 1)}:
            for val in (1.0, -3.0e-300, 5e-324, np.nan, np.inf):
                a[pos] = val
                assert ask(lib, a, threads), (n, threads, pos, val)
                a[pos] = 0.0
        assert not ask(lib, a, threads)


def test_random_planes_agree_with_numpy(lib):
    rng = np.random.default_rng(3)
    for _ in range(40):
        n = int(rng.integers(1, 3 << 20))
        a = np.zeros(n)
        k = int(rng.integers(0, 4))
        if k:
            a[rng.integers(0, n, k)] = rng.standard_normal(k)
        assert ask(lib, a, int(rng.integers(1, 9))) == bool(np.any(a != 0.0) or np.any(np.isnan(a)))


def test_scan_time_of_a_4096_squared_plane(lib, capsys):
    a = np.zeros(4096 * 4096)
    ask(lib, a, 8)
    t = {}
    for threads in (1, 8):
        t0 = time.perf_counter()
        for _ in range(3):
            assert not ask(lib, a, threads)
        t[threads] = (time.perf_counter() - t0) / 3
    with capsys.disabled():
        print("\n[host_scan] 4096^2 zero plane: %.1f ms on one thread, %.1f ms on up to 8" % (t[1] * 1e3, t[8] * 1e3))
    assert t[8] < 1.5 * t[1] + 0.005
