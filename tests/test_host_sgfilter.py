"""The host-resident sg_filtering module of the host shell (spruce_b200/host/sgfilter.hpp, compiled here with g++) against the oracle's restatement of
SGFilter::singleVarSavitzkyGolay -- which tests/test_oracle_vs_live_reference.py pins to live runs of the reference binary -- bit for bit, on random
planes, for wall and periodic iteration bounds."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import same_bits

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "sgfilter_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libsgfilter_check.so"


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC, ROOT / "spruce_b200" / "host" / "sgfilter.hpp", ROOT / "spruce_b200" / "host" / "grid.hpp"]
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    return C.CDLL(str(LIB))


@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("periodic", "periodic")), (("fixed", "open"), ("reflect", "fixed")), (("periodic", "periodic"), ("fixed", "open")),
                                   (("open", "open"), ("periodic", "periodic"))])
def test_host_filter_equals_oracle_filter(lib, xb, yb):
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    nx, ny = 31, 27                                        # ydim <= xdim: every tap reads grid(j, j) (the reference's quirk), a valid row index is required
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb)
    rng = np.random.default_rng(11)
    xl, xu = (0, nx - 1) if xb[0] == "periodic" else (2, nx - 3)
    yl, yu = (0, ny - 1) if yb[0] == "periodic" else (2, ny - 3)
    for _ in range(3):
        p = np.ascontiguousarray(rng.standard_normal((nx, ny)) * 10.0 ** rng.uniform(-3, 3))
        ref = o.sg_filter(p)
        got = p.copy()
        lib.sgfilter_apply(got.ctypes.data_as(C.c_void_p), nx, ny, xl, xu, yl, yu, int(yb[0] == "periodic"))
        assert same_bits(got, ref)
        assert not np.array_equal(got, p)
    o.close()
