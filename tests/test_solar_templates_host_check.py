"""The product's host-side template builders of the pointwise solar source-term modules (spruce_b200/csrc/solar_templates.hpp -- plain
C++ that capi.cu includes) compiled with g++ and compared bit for bit with the planes the CPU oracle builds; the oracle's modules are
pinned against live reference runs (tests/test_oracle_vs_live_reference.py).  Also checks that a slab builds exactly its rows."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import mismatch, same_bits
from oracle.oracle import Oracle
from spruce_b200 import synthetic

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "solar_templates_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libsolar_templates_check.so"


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    hdr = ROOT / "spruce_b200" / "csrc" / "solar_templates.hpp"
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-std=c++17", "-O3", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)   # -O3: as nvcc's host pass
    return C.CDLL(str(LIB))


def dp(a):
    return a.ctypes.data_as(C.c_void_p)


BOUNDS = [(("periodic", "periodic"), ("fixed", "fixed")), (("fixed", "open"), ("periodic", "periodic")), (("periodic", "periodic"), ("periodic", "periodic")),
          (("reflect", "open"), ("fixed", "open"))]


@pytest.mark.parametrize("xb,yb", BOUNDS)
@pytest.mark.parametrize("kind", ["localized_heating", "mass_injection"])
def test_positive_templates_equal_oracle(lib, xb, yb, kind):
    nx, ny = 29, 23
    s = synthetic.stratified_loop(nx, ny)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="euler", density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    peak = 1.0e-3 if kind == "localized_heating" else 2.0e6
    kw = dict(start_time=0.0, duration=50.0, stddev_x=3.0, stddev_y=4.5, center_x=2.5, center_y=20.0)       # near the edges: the periodic images matter
    kw["max_heating_rate" if kind == "localized_heating" else "max_injection_rate"] = peak
    o.add_small_module(kind, **kw)
    o.step()
    ref = o.small_module_plane(0)
    assert ref is not None and np.count_nonzero(ref) > 0
    out = np.zeros((nx, ny))
    lib.tmpl_positive(nx, ny, 0, nx, int(xb[0] == "periodic"), int(yb[0] == "periodic"), C.c_double(peak), C.c_double(3.0), C.c_double(4.5), C.c_double(2.5), C.c_double(20.0), dp(out))
    assert same_bits(out, ref), mismatch(out, ref)
    r0, nl = 7, 11                                                                                           # a slab builds exactly its rows
    part = np.zeros((nl, ny))
    lib.tmpl_positive(nx, ny, r0, nl, int(xb[0] == "periodic"), int(yb[0] == "periodic"), C.c_double(peak), C.c_double(3.0), C.c_double(4.5), C.c_double(2.5), C.c_double(20.0), dp(part))
    assert same_bits(part, ref[r0:r0 + nl])
    o.close()


@pytest.mark.parametrize("xb,yb", BOUNDS)
@pytest.mark.parametrize("angle,dirs", [(0.0, (1.0, 0.5)), (20.0, (-1.0, 0.5)), (-35.0, (0.3, -2.0))])
def test_momentum_templates_equal_oracle(lib, xb, yb, angle, dirs):
    nx, ny = 27, 24
    s = synthetic.stratified_loop(nx, ny)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="euler", density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o.add_small_module("momentum_injection", start_time=0.0, duration=50.0, max_accel=1.0e3, stddev_x=4.0, stddev_y=3.0, center_x=13.0, center_y=9.0,
                       dir_x=dirs[0], dir_y=dirs[1], template_angle=angle, oscillatory=0.0, oscillation_period=3.0)
    rx, ry = o.small_module_plane(0, 0), o.small_module_plane(0, 1)
    ox, oy = np.zeros((nx, ny)), np.zeros((nx, ny))
    lib.tmpl_momentum(nx, ny, 0, nx, int(xb[0] == "periodic"), int(yb[0] == "periodic"), C.c_double(4.0), C.c_double(3.0), C.c_double(13.0), C.c_double(9.0),
                      C.c_double(dirs[0]), C.c_double(dirs[1]), C.c_double(angle), dp(ox), dp(oy))
    assert same_bits(ox, rx), mismatch(ox, rx)
    assert same_bits(oy, ry), mismatch(oy, ry)
    o.close()


BC = {"periodic": 0, "open": 1, "fixed": 2, "reflect": 3, "open_moc": 4, "open_ucnp": 5}
OUTFLOW = [
    (("periodic", "periodic"), ("fixed", "open_moc"), "y_bound_2", "exp", 6.0e8, 3.0e8),
    (("periodic", "periodic"), ("fixed", "open"), "y_bound_2", "gaussian", 5.0e8, 0.0),
    (("open_moc", "fixed"), ("reflect", "reflect"), "x_bound_1", "flat", 4.0e8, 2.0e8),
    (("fixed", "open"), ("fixed", "fixed"), "x_bound_2", "exp", 7.0e8, 2.5e8),
    (("fixed", "fixed"), ("open_moc", "fixed"), "y_bound_1", "gaussian", 3.0e8, 1.0e8),
]


@pytest.mark.parametrize("xb,yb,boundary,shape,length,feather", OUTFLOW)
def test_boundary_outflow_template_and_window_equal_oracle(lib, xb, yb, boundary, shape, length, feather):
    nx, ny = 31, 28
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="euler", density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    bnd = {"x_bound_1": 0, "x_bound_2": 1, "y_bound_1": 2, "y_bound_2": 3}[boundary]
    shp = {"exp": 0, "gaussian": 1, "flat": 2}[shape]
    target = 2.0e6
    o.add_small_module("boundary_outflow", max_accel=2.0e3, falloff_length=length, boundary=bnd, falloff_shape=shp, feather_length=feather, field_aligned_mode=1.0,
                       dynamic_mode=1.0, dynamic_time=10.0, dynamic_target_speed=target)
    ref = o.small_module_plane(0)
    assert ref is not None and np.count_nonzero(ref) > 0
    X = np.ascontiguousarray(s["planes"]["pos_x"], dtype=np.float64); Y = np.ascontiguousarray(s["planes"]["pos_y"], dtype=np.float64)
    out = np.zeros((nx, ny)); win = (C.c_int * 4)()
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    lib.tmpl_outflow(nx, ny, bc, dp(X), dp(Y), C.c_double(length), C.c_double(feather), bnd, shp, dp(out), win)
    assert same_bits(out, ref), mismatch(out, ref)
    # the window, through the quantity it exists for: max over it of the signed field-aligned speed (boundaryoutflow.cpp:215-236), on a developed flow
    for _ in range(3):
        o.step()
    xl, xu, yl, yu = list(win)
    hx, hy, vx, vy = (o.get(v)[xl:xu + 1, yl:yu + 1] for v in ("b_hat_x", "b_hat_y", "v_x", "v_y"))
    cur = hx * vx + hy * vy
    flip = (hx > 0.0) if bnd == 0 else (hx < 0.0) if bnd == 1 else (hy < 0.0) if bnd == 3 else (hy > 0.0)
    cur = np.where(flip, cur * -1.0, cur)
    mine = max(-1.0 * target, float(np.max(cur))) if cur.size else -1.0 * target
    assert mine == o.outflow_mean(0), (mine, o.outflow_mean(0), list(win))
    o.close()
