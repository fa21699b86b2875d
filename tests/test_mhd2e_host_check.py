"""The product's IdealMHD2E path -- per-cell functions (spruce_b200/csrc/mhd2e_cells.cuh) in the product's stage order (mhd2e_step.hpp), both
compiled for the HOST by tests/hostcheck/mhd2e_host_check.cpp with loops in place of kernel launches -- against the CPU restatement
(oracle/ideal_mhd2e_oracle.inc, pinned to live reference runs): step-size history, every evolved plane, dt and the right-hand side, bit for bit."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import mismatch, same_bits
from oracle.oracle import EVOLVED_2E, Oracle2E
from test_oracle_vs_live_reference import e2_cases, e2_state

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "mhd2e_host_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libmhd2e_host_check.so"
BC = {"periodic": 0, "open": 1, "fixed": 2, "reflect": 3, "open_moc": 4, "open_ucnp": 5}
TI = {"euler": 0, "rk2": 1, "rk4": 2}


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC, ROOT / "spruce_b200" / "csrc" / "mhd2e_cells.cuh", ROOT / "spruce_b200" / "csrc" / "mhd2e_step.hpp"]
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    L = C.CDLL(str(LIB))
    L.mhd2e_host_run.restype = C.c_int
    return L


@pytest.mark.parametrize("k,xb,yb,integrator,nx,ny,loop,nmin", e2_cases())
def test_product_mhd2e_steps_equal_oracle(lib, k, xb, yb, integrator, nx, ny, loop, nmin):
    s = e2_state(nx, ny, loop)
    floors = dict(density_min=nmin, temp_min=1.0e4, thermal_energy_min=1.0e-6) if loop else dict(density_min=1.0, temp_min=1.0, thermal_energy_min=1.0e-30)
    o = Oracle2E(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator=integrator, **floors)
    nsteps = 5
    ref_steps = [o.step() for _ in range(nsteps)]
    names = ["rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "be_x", "be_y", "grav_x", "grav_y"]
    planes = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in names]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    arr = (C.c_void_p * 11)(*[p.ctypes.data for p in planes])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    out = np.zeros((7, nx, ny)); dt = np.zeros((nx, ny)); steps = np.zeros(nsteps); rhs = np.zeros((7, nx, ny))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.mhd2e_host_run(arr, vp(dx), vp(dy), C.c_int(nx), C.c_int(ny), bc, C.c_int(TI[integrator]), C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]),
                            C.c_double(0.2), C.c_double(floors["density_min"]), C.c_double(floors["temp_min"]), C.c_double(floors["thermal_energy_min"]),
                            C.c_double(1.0), C.c_double(0.5), C.c_int(nsteps), vp(out), vp(dt), vp(steps), vp(rhs))
    assert rc == 0
    assert [float(x).hex() for x in steps] == [x.hex() for x in ref_steps]
    for v, nm in enumerate(EVOLVED_2E):
        assert same_bits(out[v], o.get(nm)), "case %d %s: %s" % (k, nm, mismatch(out[v], o.get(nm)))
    assert same_bits(dt, o.get("dt")), "case %d dt: %s" % (k, mismatch(dt, o.get("dt")))
    k_ref = o.rhs()
    for v, nm in enumerate(EVOLVED_2E):
        assert same_bits(rhs[v], k_ref[v]), "case %d d(%s)/dt: %s" % (k, nm, mismatch(rhs[v], k_ref[v]))
    o.close()


@pytest.mark.parametrize("n_ranks", [2, 3])
@pytest.mark.parametrize("k,xb,yb,integrator,nx,ny,loop,nmin", e2_cases())
def test_product_mhd2e_slab_steps_equal_oracle(lib, k, xb, yb, integrator, nx, ny, loop, nmin, n_ranks):
    """The slab form: every rank holds its rows plus two halo rows per side (NaN where nothing is resident), addresses them by global row through
    shifted pointers as mhd2e_host.cuh does, runs each phase, and the halo rows travel after every stage (also those of the primary state when a
    wall-type boundary pass wrote it).  Result: the oracle's planes and step sizes, bit for bit."""
    s = e2_state(nx, ny, loop)
    floors = dict(density_min=nmin, temp_min=1.0e4, thermal_energy_min=1.0e-6) if loop else dict(density_min=1.0, temp_min=1.0, thermal_energy_min=1.0e-30)
    o = Oracle2E(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator=integrator, **floors)
    nsteps = 5
    ref_steps = [o.step() for _ in range(nsteps)]
    names = ["rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "be_x", "be_y", "grav_x", "grav_y"]
    planes = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in names]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    arr = (C.c_void_p * 11)(*[p.ctypes.data for p in planes])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    out = np.zeros((7, nx, ny)); steps = np.zeros(nsteps)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.mhd2e_host_run_slabs.restype = C.c_int
    rc = lib.mhd2e_host_run_slabs(arr, vp(dx), vp(dy), C.c_int(nx), C.c_int(ny), bc, C.c_int(TI[integrator]), C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]),
                                  C.c_double(0.2), C.c_double(floors["density_min"]), C.c_double(floors["temp_min"]), C.c_double(floors["thermal_energy_min"]),
                                  C.c_double(1.0), C.c_double(0.5), C.c_int(nsteps), C.c_int(n_ranks), vp(out), vp(steps))
    assert rc == 0
    assert [float(x).hex() for x in steps] == [x.hex() for x in ref_steps]
    for v, nm in enumerate(EVOLVED_2E):
        assert same_bits(out[v], o.get(nm)), "case %d %d ranks %s: %s" % (k, n_ranks, nm, mismatch(out[v], o.get(nm)))
    o.close()
