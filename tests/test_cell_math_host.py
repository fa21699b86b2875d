"""spruce_b200/csrc/cell_math.cuh compiled for the HOST (tests/hostcheck/cell_math_check.cpp):
  * cell_dt == the reference's recomputeDT (idealmhd.cpp:279-304) bit for bit, including the |v| shortcut for sqrt(v*v) and the shared reciprocal;
  * dt_can_skip is SOUND: a cell it lets the stage kernel skip never has dt <= 1/R, for thresholds from far below to far above the cells' dt
    (this is what makes the pruned minimum equal the full minimum whenever k_dt_validate accepts the window);
  * density_floor == the reference's enforceMinimums + derived-variable round trip for rho."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "cell_math_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libcell_math_check.so"
M_I, GAMMA = 1.6726e-24, 5.0 / 3.0


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC, ROOT / "spruce_b200" / "csrc" / "cell_math.cuh", ROOT / "spruce_b200" / "csrc" / "exact_math.cuh"]
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    return C.CDLL(str(LIB))


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def states(rng, n):
    """coronal to UCNP-like magnitudes, sub- and super-Alfvenic, zero and tiny velocities, zero field"""
    nn = 10.0 ** rng.uniform(6, 12, n)
    rho = nn * M_I
    T = 10.0 ** rng.uniform(3, 7, n)
    e = nn * 2 * 1.3807e-16 * T / (GAMMA - 1.0)
    cs = np.sqrt(GAMMA * (GAMMA - 1.0) * e / rho)
    v = cs[:, None] * rng.standard_normal((n, 2)) * 10.0 ** rng.uniform(-3, 1.5, (n, 1))
    b = rng.standard_normal((n, 3)) * 10.0 ** rng.uniform(-2, 2.5, (n, 1))
    k = n // 10
    v[:k] = 0.0; v[k:2 * k, 0] = 0.0; b[2 * k:3 * k] = 0.0; b[3 * k:4 * k, 2] = 0.0
    v[4 * k:4 * k + 100] *= 1e-200                                       # underflowing v*v: the sqrt(v*v) path must be taken
    dx = 10.0 ** rng.uniform(6, 8, n); dy = dx * (0.5 + rng.random(n))
    return np.ascontiguousarray(np.column_stack([rho, rho * v[:, 0], rho * v[:, 1], e, b[:, 0], b[:, 1], b[:, 2], dx, dy]))


def test_cell_dt_equals_reference_formula(lib):
    rng = np.random.default_rng(3)
    n = 500_000
    u = states(rng, n)
    out = np.zeros((n, 2)); skip = np.zeros(n, dtype=np.int32)
    lib.cell_math_dt(C.c_int(n), vp(u), C.c_double(M_I), C.c_double(GAMMA), C.c_double(0.0), vp(out), vp(skip))
    assert np.array_equal(out[:, 0], out[:, 1]), "cell_dt differs from recomputeDT in %d cells" % int((out[:, 0] != out[:, 1]).sum())
    assert not skip.any(), "R = 0 (no threshold known) must evaluate every cell"


@pytest.mark.parametrize("quantile", [0.001, 0.05, 0.5, 0.95])
def test_dt_can_skip_is_sound(lib, quantile):
    rng = np.random.default_rng(17)
    n = 500_000
    u = states(rng, n)
    u[:, 7:9] = 10.0 ** 7 * (0.5 + rng.random((n, 2)))                   # comparable cell sizes, so that a single threshold separates the cells
    out = np.zeros((n, 2)); skip = np.zeros(n, dtype=np.int32)
    lib.cell_math_dt(C.c_int(n), vp(u), C.c_double(M_I), C.c_double(GAMMA), C.c_double(0.0), vp(out), vp(skip))
    thr = float(np.quantile(out[:, 1], quantile))                        # F * (previous global minimum) in the kernel
    lib.cell_math_dt(C.c_int(n), vp(u), C.c_double(M_I), C.c_double(GAMMA), C.c_double(1.0 / thr), vp(out), vp(skip))
    skipped = skip.astype(bool)
    assert skipped.any(), "the test should skip something at this threshold"
    assert np.all(out[skipped, 1] > thr), "dt_can_skip dropped %d cells whose dt is not above the threshold" % int((out[skipped, 1] <= thr).sum())
    kept_above = (~skipped) & (out[:, 1] > thr)
    # not a requirement, but the point of the test: most cells above the threshold are skipped
    assert kept_above.sum() < 0.5 * (out[:, 1] > thr).sum() + 10


def test_density_floor_equals_reference_round_trip(lib):
    rng = np.random.default_rng(23)
    n = 500_000
    n_min = 1.0e7
    rho_u = M_I * 10.0 ** rng.uniform(5, 11, n)
    rho_u[:1000] = -rho_u[:1000]                                         # negative densities after a step are floored
    rho_u[1000:1100] = 0.0
    out = np.zeros((n, 4))
    lib.cell_math_floor(C.c_int(n), vp(rho_u), C.c_double(M_I), C.c_double(n_min), vp(out))
    assert np.array_equal(out[:, 0], out[:, 2]) and np.array_equal(out[:, 1], out[:, 3])
