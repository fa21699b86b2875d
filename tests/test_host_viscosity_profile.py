"""Viscosity::getBoundaryViscosity in the host shell (spruce_b200/host/viscosity_profile.hpp, compiled here with g++): all four shapes -- gaussian, exp,
exp_elliptical, gaussian_elliptical -- bit for bit against the restatement of tests/golden_util.py, which tests/test_oracle_vs_live_reference.py pins to live
runs of the reference binary (av_mixed, av_boundary_exp, av_boundary_exp_elliptical, av_boundary_gaussian_elliptical); an unknown shape is refused."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import boundary_viscosity_profile, same_bits
from spruce_b200 import synthetic

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "viscosity_profile_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libviscosity_profile_check.so"
DP = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC, ROOT / "spruce_b200" / "host" / "viscosity_profile.hpp", ROOT / "spruce_b200" / "host" / "grid.hpp"]
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    return C.CDLL(str(LIB))


@pytest.mark.parametrize("shape", ["gaussian", "exp", "exp_elliptical", "gaussian_elliptical"])
@pytest.mark.parametrize("nx,ny,strength,length", [(31, 27, 0.8, 6.0e8), (24, 40, 2.5, 1.5e9), (26, 23, 0.3, 2.0e8)])
def test_profile_equals_the_pinned_restatement(lib, shape, nx, ny, strength, length):
    s = synthetic.stratified_loop(nx, ny)
    x, y = np.ascontiguousarray(s["planes"]["pos_x"]), np.ascontiguousarray(s["planes"]["pos_y"])
    out = np.zeros((nx, ny))
    assert lib.boundary_viscosity_profile(x.ctypes.data_as(DP), y.ctypes.data_as(DP), nx, ny, C.c_double(strength), C.c_double(length), shape.encode(), out.ctypes.data_as(DP)) == 0
    ref = boundary_viscosity_profile(x, y, strength, length, shape)
    assert same_bits(out, ref)
    assert out.max() <= strength and out.min() >= 0.0 and out.max() > 0.0


def test_unknown_shape_is_refused(lib):
    x = np.zeros((6, 6))
    assert lib.boundary_viscosity_profile(x.ctypes.data_as(DP), x.ctypes.data_as(DP), 6, 6, C.c_double(1.0), C.c_double(1.0), b"hexagonal", x.copy().ctypes.data_as(DP)) == 1
