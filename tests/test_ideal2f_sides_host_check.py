"""The ordered two-fluid boundary passes of the product (spruce_b200/csrc/ideal2f_sides.cuh, used when an open_ucnp side meets a fixed / reflect side)
compiled for the host and compared with the CPU restatement's updateGhostZones (oracle/ideal2f_oracle.inc, pinned to the reference) on random planes,
whole domain and slab by slab (halo rows must stay untouched)."""
import ctypes as C
import itertools
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import mismatch, same_bits
from oracle.oracle import Oracle2F, VARS_2F
from spruce_b200 import synthetic

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "ideal2f_sides_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libideal2f_sides_check.so"
BC = {"periodic": 0, "open": 1, "fixed": 2, "reflect": 3, "open_moc": 4, "open_ucnp": 5}
EVOLVED = ["i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_thermal_energy", "e_thermal_energy", "E_x", "E_y", "E_z", "bi_x", "bi_y", "bi_z"]


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    hdr = ROOT / "spruce_b200" / "csrc" / "ideal2f_sides.cuh"
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    L = C.CDLL(str(LIB))
    L.tf2_host_sides.restype = C.c_int
    return L


SIDES = ["fixed", "reflect", "open_ucnp"]
COMBOS = [(xb, yb) for xb in itertools.product(SIDES, repeat=2) for yb in itertools.product(SIDES, repeat=2)][::3] + \
         [(("periodic", "periodic"), ("open_ucnp", "reflect")), (("reflect", "open_ucnp"), ("periodic", "periodic"))]


@pytest.mark.parametrize("n_ranks", [1, 3])
@pytest.mark.parametrize("xb,yb", COMBOS)
def test_ordered_two_fluid_boundary_passes_equal_oracle(lib, xb, yb, n_ranks):
    nx, ny = 19, 17
    s = synthetic.ucnp_cloud(nx, ny, drift=2.0e3, bfield=5.0)
    o = Oracle2F(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="euler", setup=False)
    rng = np.random.default_rng(hash((xb, yb)) % (2 ** 32))
    for v in EVOLVED:
        o.view(v)[...] = rng.standard_normal((nx, ny)) + (3.0 if "rho" in v or "thermal" in v else 0.0)
    mine = [o.get(v) for v in EVOLVED]
    before = [m.copy() for m in mine]
    o.apply_ghosts()
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    ptr = (C.c_void_p * 14)(*[m.ctypes.data for m in mine])
    assert lib.tf2_host_sides(ptr, C.c_int(nx), C.c_int(ny), bc, C.c_int(n_ranks)) == 0, "a slab wrote into a halo row"
    changed = 0
    for v, m, b0 in zip(EVOLVED, mine, before):
        assert same_bits(m, o.get(v)), "%s %s %s: %s" % (xb, yb, v, mismatch(m, o.get(v)))
        changed += int((m != b0).sum())
    assert changed > 0
    o.close()
