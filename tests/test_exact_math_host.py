"""The device arithmetic header of the stage kernels (spruce_b200/csrc/exact_math.cuh), compiled for the HOST by tests/hostcheck/exact_math_check.cpp:
  * ddiv(a, b, RN(1/b)) == RN(a/b) on random and adversarial operands (the sequence itself is also proven in tests/test_exact_division.py);
  * upwind_face / upwind_face_sel / upwind_face_far reproduce the reference's upwindSurface selection (derivs.cpp:47-68) bit for bit, ties, equal
    neighbours, zero and negative face velocities, NaN operands and non-uniform cell sizes included.
A regression guard for kernel work: these are the functions an optimisation of the stage kernel touches first, and this check needs no GPU."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "exact_math_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libexact_math_check.so"


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    hdr = ROOT / "spruce_b200" / "csrc" / "exact_math.cuh"
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)   # fma() calls stay explicit fmas
    return C.CDLL(str(LIB))


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def same(a, b):
    return (a == b) | (np.isnan(a) & np.isnan(b))


def test_ddiv_is_the_correctly_rounded_quotient(lib):
    rng = np.random.default_rng(11)
    n = 2_000_000
    a = rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)
    b = (1.0 + rng.random(n)) * 10.0 ** rng.integers(-20, 20, n) * rng.choice([-1.0, 1.0], n)
    a[:1000] = 0.0                                        # zero numerators
    a[1000:2000] = b[1000:2000] * (1.0 + 2.0 ** -52)      # quotients next to 1
    a[2000:3000] = b[2000:3000] * 3.0                     # exact quotients
    out = np.zeros(n)
    lib.exact_math_div(C.c_int(n), vp(a), vp(b), vp(out))
    assert np.array_equal(out, a / b)


def face_cases(rng, n):
    q = rng.standard_normal((n, 4)) * 10.0 ** rng.integers(-3, 4, (n, 1))
    k = n // 8
    q[:k, 2] = q[:k, 1]                                   # q0 == qm1 (the <= tie)
    q[k:2 * k, 0] = q[k:2 * k, 1]                         # flat upwind side
    q[2 * k:3 * k] = np.round(q[2 * k:3 * k])             # many exact equalities between d1, d2, d3
    q[3 * k:3 * k + 50, rng.integers(0, 4, 50)] = np.nan  # NaN operands
    q[3 * k + 50:3 * k + 100] = 0.0                       # all zero
    q[3 * k + 100:3 * k + 200, 1:3] *= -0.0               # signed zeros
    h = 0.5 * (1.0 + rng.random((n, 4))) * 10.0 ** rng.integers(6, 9, (n, 1))
    h[4 * k:5 * k] = h[4 * k:5 * k, :1]                   # uniform grid
    vf = rng.standard_normal(n)
    vf[5 * k:5 * k + 200] = 0.0
    vf[5 * k + 200:5 * k + 300] = -0.0
    vf[5 * k + 300:5 * k + 320] = np.nan
    return np.ascontiguousarray(q), np.ascontiguousarray(h), np.ascontiguousarray(vf)


def test_barton_face_forms_equal_the_reference_formula(lib):
    rng = np.random.default_rng(5)
    n = 1_000_000
    q, h, vf = face_cases(rng, n)
    out = np.zeros((n, 5))
    lib.exact_math_faces(C.c_int(n), vp(q), vp(h), vp(vf), vp(out))
    ref, d2, f1, f2, f3 = out.T
    finite_in = np.isfinite(q).all(axis=1)
    assert same(d2[finite_in], d2[finite_in]).all() and not np.isnan(d2[finite_in]).any(), "a linear interpolation d2 differs between the forms"
    assert same(f1, ref).all(), "upwind_face differs from the reference formula in %d cases" % int((~same(f1, ref)).sum())
    assert same(f2, ref).all(), "upwind_face_sel differs in %d cases" % int((~same(f2, ref)).sum())
    moving = (vf > 0.0) | (vf < 0.0)                      # upwind_face_far returns the limited value also for vf == 0; the kernel multiplies it by vf
    assert same(f3[moving], ref[moving]).all(), "upwind_face_far differs in %d cases" % int((~same(f3[moving], ref[moving])).sum())
    still = ~moving & finite_in & np.isfinite(vf)
    assert np.all((f3[still] * vf[still] == 0.0) & (ref[still] * vf[still] == 0.0))


def test_all_zero8(lib):
    z = np.zeros(8)
    assert lib.exact_math_all_zero8(vp(z)) == 1
    z[3] = -0.0
    assert lib.exact_math_all_zero8(vp(z)) == 1
    z[5] = 5e-324
    assert lib.exact_math_all_zero8(vp(z)) == 0
    z[5] = np.nan
    assert lib.exact_math_all_zero8(vp(z)) == 0
