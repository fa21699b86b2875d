"""The arithmetic of the two host-resident UCNP modules of the host shell (spruce_b200/host/ucnp_modules.hpp, compiled here with g++) against the oracle's
restatements of CoulombExplosion::postIterateModule and GlobalTemperature::postIterateModule -- which tests/test_oracle_vs_live_reference.py pins to live runs
of the reference binary -- bit for bit: the force planes and the momentum update of one coulomb_explosion hook, the temperature after the diffusion sub-steps,
and the refusal of a grid on which the reference aborts (an empty radial bin)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import same_bits
from oracle.oracle import Oracle
from spruce_b200 import synthetic

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "ucnp_modules_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libucnp_modules_check.so"
DP = C.POINTER(C.c_double)
LAP = C.CFUNCTYPE(None, DP, DP)


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC, ROOT / "spruce_b200" / "host" / "ucnp_modules.hpp", ROOT / "spruce_b200" / "host" / "grid.hpp"]
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    L = C.CDLL(str(LIB))
    L.ucnp_coulomb_force.argtypes = [DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, DP, DP]
    L.ucnp_diffuse_temperature.argtypes = [DP, DP, DP, DP, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, LAP]
    return L


def dp(a):
    return a.ctypes.data_as(DP)


KW = dict(density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1.0e-30)


@pytest.mark.parametrize("nx,ny,t,xb,yb", [(83, 79, 0.0, ("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp")), (90, 101, 4.0e-7, ("periodic", "periodic"), ("fixed", "fixed"))])
def test_coulomb_force_equals_oracle(lib, nx, ny, t, xb, yb):
    s = synthetic.ucnp_cloud_mhd(nx, ny, drift=30.0)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, **KW)
    o.add_small_module("coulomb_explosion", timescale=1.0e-6, lengthscale=0.17, strength=2.0e-3)
    o.set_time(t)
    n, mx0, my0 = o.get("n").copy(), o.get("mom_x").copy(), o.get("mom_y").copy()
    dt = 3.0e-8
    o.small_module_hooks(2, dt)                                            # the post-iterate hook alone
    fx, fy = np.zeros((nx, ny)), np.zeros((nx, ny))
    x, y = np.ascontiguousarray(s["planes"]["pos_x"]), np.ascontiguousarray(s["planes"]["pos_y"])
    assert lib.ucnp_coulomb_force(dp(x), dp(y), dp(n), nx, ny, t, 1.0e-6, 0.17, 2.0e-3, dp(fx), dp(fy)) == 0
    assert same_bits(fx, o.coulomb_plane(0, "F_x")) and same_bits(fy, o.coulomb_plane(0, "F_y"))
    assert np.abs(fx).max() > 0.0
    # the shell's update (host/module.cpp: CoulombExplosion::postIterateModule): mom += F * dt, then propagateChanges, which leaves the momenta of the interior as they are
    xl, xu = (0, nx) if xb[0] == "periodic" else (3, nx - 3)
    yl, yu = (0, ny) if yb[0] == "periodic" else (3, ny - 3)
    assert same_bits((mx0 + fx * dt)[xl:xu, yl:yu], o.get("mom_x")[xl:xu, yl:yu]) and same_bits((my0 + fy * dt)[xl:xu, yl:yu], o.get("mom_y")[xl:xu, yl:yu])
    o.close()


def test_coulomb_refuses_a_grid_with_an_empty_radial_bin(lib):
    nx = ny = 41                                                           # the reference aborts here (101 radial bins over 21 x 21 distinct lattice radii)
    s = synthetic.ucnp_cloud_mhd(nx, ny)
    x, y, n = (np.ascontiguousarray(s["planes"][k]) for k in ("pos_x", "pos_y", "rho"))
    f = np.zeros((nx, ny))
    assert lib.ucnp_coulomb_force(dp(x), dp(y), dp(n), nx, ny, 0.0, 1.0e-6, 0.2, 1.0e-3, dp(f), dp(f.copy())) == 1
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=("open_ucnp", "open_ucnp"), yb=("open_ucnp", "open_ucnp"), **KW)
    o.add_small_module("coulomb_explosion", timescale=1.0e-6, lengthscale=0.2, strength=1.0e-3)
    o.small_module_hooks(2, 1.0e-8)
    with pytest.raises(RuntimeError):
        o.coulomb_plane(0, "F_x")
    o.close()


@pytest.mark.parametrize("xb,yb,strength", [(("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp"), 3.7), (("periodic", "periodic"), ("fixed", "reflect"), 1.0), (("fixed", "fixed"), ("periodic", "periodic"), 5.2)])
def test_temperature_diffusion_equals_oracle(lib, xb, yb, strength):
    nx, ny = 37, 30
    s = synthetic.ucnp_cloud_mhd(nx, ny)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, **KW)
    o.add_small_module("global_temperature", gt_strength=strength, gt_use_diffusion=1.0)
    temp, n = o.get("temp").copy(), o.get("n").copy()
    dt = 2.0e-8
    lap = LAP(lambda q, out: np.ctypeslib.as_array(out, (nx, ny)).__setitem__(slice(None), o.operator("laplacian", 0, np.ctypeslib.as_array(q, (nx, ny)))))
    mask = np.zeros((nx, ny))
    xl, xu = (0, nx) if xb[0] == "periodic" else (2, nx - 2)
    yl, yu = (0, ny) if yb[0] == "periodic" else (2, ny - 2)
    mask[xl:xu, yl:yu] = 1.0
    d_x, d_y = np.ascontiguousarray(s["planes"]["d_x"]), np.ascontiguousarray(s["planes"]["d_y"])
    steps = lib.ucnp_diffuse_temperature(dp(temp), dp(d_x), dp(d_y), dp(mask), nx, ny, dt, 0.2, strength, lap)
    assert steps == int(strength)
    o.small_module_hooks(2, dt)
    # the module then sets thermal_energy = n * K_B * temp / (gamma - 1) (global_temperature.cpp:90) and propagates
    e = n * 1.3807e-16 * temp / (s["adiabatic_index"] - 1.0)
    assert same_bits(e[xl:xu, yl:yu], o.get("thermal_energy")[xl:xu, yl:yu])
    assert not np.array_equal(temp, s["planes"]["temp"])
    o.close()
