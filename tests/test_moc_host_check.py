"""The product's per-cell open_moc arithmetic (spruce_b200/csrc/moc_kernels.cuh, the functions the CUDA kernel k_moc_stage calls),
compiled for the HOST by tests/hostcheck/moc_host_check.cpp and compared bit for bit with the CPU oracle's right-hand side in the
evolved ghost cells -- the part of the open_moc device path that can be proven without a GPU.  The oracle itself is pinned against
live reference runs (tests/test_oracle_vs_live_reference.py).  Not a product path: the library has no CPU fallback."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import mismatch, same_bits
from oracle.oracle import Oracle
from spruce_b200 import synthetic

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "moc_host_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libmoc_host_check.so"
BC = {"periodic": 0, "open": 1, "fixed": 2, "reflect": 3, "open_moc": 4, "open_ucnp": 5}
EVOLVED = ["rho", "mom_x", "mom_y", "mom_z", "thermal_energy", "bi_x", "bi_y", "bi_z"]


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    hdr = ROOT / "spruce_b200" / "csrc" / "moc_kernels.cuh"
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        # x86-64 baseline ISA has no FMA; -ffp-contract=off states it anyway (the CUDA build uses -fmad=false)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    L = C.CDLL(str(LIB))
    L.moc_host_terms.restype = C.c_int
    return L


def host_terms(L, o, xb, yb, visc, ion_mass, adiabatic_index):
    names = ["n", "mom_x", "mom_y", "mom_z", "thermal_energy", "bi_x", "bi_y", "bi_z", "be_x", "be_y", "be_z", "grav_x", "grav_y"]
    planes = [np.ascontiguousarray(o.get(v), dtype=np.float64) for v in names]
    nx, ny = planes[0].shape
    dx = np.ascontiguousarray(o.get("d_x")[:, 0]); dy = np.ascontiguousarray(o.get("d_y")[0, :])
    arr = (C.c_void_p * 13)(*[p.ctypes.data for p in planes])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    k = np.zeros((8, nx, ny)); owned = np.zeros((nx, ny), dtype=np.uint8)
    cnt = L.moc_host_terms(arr, dx.ctypes.data_as(C.c_void_p), dy.ctypes.data_as(C.c_void_p), C.c_int(nx), C.c_int(ny), bc,
                           C.c_double(ion_mass), C.c_double(adiabatic_index), C.c_double(visc), k.ctypes.data_as(C.c_void_p), owned.ctypes.data_as(C.c_void_p))
    return k, owned.astype(bool), cnt


def dt_bounds(xb, yb, nx, ny):
    lo = lambda b: 0 if b in ("periodic", "open_moc") else 2
    hi = lambda b, n: n - 1 if b in ("periodic", "open_moc") else n - 3
    return lo(xb[0]), hi(xb[1], nx), lo(yb[0]), hi(yb[1], ny)


CASES = [
    ("y2_moc", ("periodic", "periodic"), ("fixed", "open_moc"), 0.0, 26, 23),
    ("y1_moc_visc", ("periodic", "periodic"), ("open_moc", "fixed"), 0.3, 24, 25),
    ("x1_moc", ("open_moc", "reflect"), ("fixed", "open"), 0.0, 27, 22),
    ("all_moc_visc", ("open_moc", "open_moc"), ("open_moc", "open_moc"), 0.1, 25, 24),
    ("x2_moc_y2_moc", ("fixed", "open_moc"), ("open", "open_moc"), 0.0, 23, 26),
    ("x_moc_y_periodic_visc", ("open_moc", "open_moc"), ("periodic", "periodic"), 0.2, 22, 27),
    ("x1_moc_only", ("open_moc", "fixed"), ("reflect", "reflect"), 0.0, 31, 19),
]


@pytest.mark.parametrize("name,xb,yb,gvisc,nx,ny", CASES, ids=[c[0] for c in CASES])
def test_product_moc_cell_terms_equal_oracle_rhs(lib, name, xb, yb, gvisc, nx, ny):
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    kw = dict(xb=xb, yb=yb, integrator="rk2", density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    o.set_global_viscosity(gvisc)
    n_owned = 0
    for it in range(4):                                       # the flow develops: inflow / outflow switches change sign along the strips
        xl, xu, yl, yu = dt_bounds(xb, yb, nx, ny)
        dxp, dyp, dt = o.get("d_x"), o.get("d_y"), o.get("dt")
        term = (1.0 / (1.0 / (dxp * dxp) + 1.0 / (dyp * dyp))) / dt
        visc = gvisc * 0.5 * float(np.min(term[xl:xu + 1, yl:yu + 1]))                    # idealmhd.cpp:90
        k_ref = o.rhs()
        k, owned, cnt = host_terms(lib, o, xb, yb, visc, s["ion_mass"], s["adiabatic_index"])
        assert cnt == int(owned.sum()) and cnt > 0
        n_owned = cnt
        mask = np.zeros((nx, ny), dtype=bool)                                               # ghost cells = outside the interior bounds
        ilo = lambda b: 0 if b == "periodic" else 2
        ihi = lambda b, n: n - 1 if b == "periodic" else n - 3
        mask[ilo(xb[0]):ihi(xb[1], nx) + 1, ilo(yb[0]):ihi(yb[1], ny) + 1] = True
        assert not np.any(owned & mask), "an interior cell was claimed by an open_moc side"
        for v, nm in enumerate(EVOLVED):
            ref = np.where(owned, k_ref[v], 0.0)
            assert same_bits(np.where(owned, k[v], 0.0), ref), "%s it %d d(%s)/dt: %s" % (name, it, nm, mismatch(np.where(owned, k[v], 0.0), ref))
            ghost_not_owned = (~mask) & (~owned)
            assert np.all(k_ref[v][ghost_not_owned] == 0.0), "the oracle evolves a ghost cell the product does not claim (%s)" % nm
        o.step()
    assert n_owned > 0
    o.close()


@pytest.mark.parametrize("name,xb,yb,gvisc,nx,ny", CASES, ids=[c[0] for c in CASES])
def test_product_moc_euler_update_equals_oracle_step(lib, name, xb, yb, gvisc, nx, ny):
    """advance_cell + cell_dt_plain (the tail of k_moc_stage): one euler step of the evolved ghost cells -- increment, floors, the
    momentum zeroing of fixed / reflect sides that sweep into the strip, n round trip -- and their new dt, against the oracle's step."""
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    floors = dict(density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="euler", **floors)
    o.set_global_viscosity(gvisc)
    lib.moc_host_euler.restype = C.c_int
    names = ["n", "mom_x", "mom_y", "mom_z", "thermal_energy", "bi_x", "bi_y", "bi_z", "be_x", "be_y", "be_z", "grav_x", "grav_y"]
    for it in range(3):
        planes = [np.ascontiguousarray(o.get(v), dtype=np.float64) for v in names]
        dxp, dyp, dt = o.get("d_x"), o.get("d_y"), o.get("dt")
        xl, xu, yl, yu = dt_bounds(xb, yb, nx, ny)
        visc = gvisc * 0.5 * float(np.min(((1.0 / (1.0 / (dxp * dxp) + 1.0 / (dyp * dyp))) / dt)[xl:xu + 1, yl:yu + 1]))
        dx = np.ascontiguousarray(dxp[:, 0]); dy = np.ascontiguousarray(dyp[0, :])
        arr = (C.c_void_p * 13)(*[p.ctypes.data for p in planes])
        bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
        out = np.zeros((8, nx, ny)); dt_out = np.zeros((nx, ny)); owned = np.zeros((nx, ny), dtype=np.uint8)
        step = o.step()
        cnt = lib.moc_host_euler(arr, dx.ctypes.data_as(C.c_void_p), dy.ctypes.data_as(C.c_void_p), C.c_int(nx), C.c_int(ny), bc, C.c_double(s["ion_mass"]),
                                 C.c_double(s["adiabatic_index"]), C.c_double(visc), C.c_double(floors["density_min"]), C.c_double(floors["thermal_energy_min"]),
                                 C.c_double(step), out.ctypes.data_as(C.c_void_p), dt_out.ctypes.data_as(C.c_void_p), owned.ctypes.data_as(C.c_void_p))
        own = owned.astype(bool)
        assert cnt == int(own.sum()) and cnt > 0
        for v, nm in enumerate(names[:8]):
            ref = np.where(own, o.get(nm), 0.0)
            assert same_bits(np.where(own, out[v], 0.0), ref), "%s it %d %s: %s" % (name, it, nm, mismatch(np.where(own, out[v], 0.0), ref))
        have_dt = own & (dt_out >= 0.0)
        assert np.any(have_dt)
        assert same_bits(np.where(have_dt, dt_out, 0.0), np.where(have_dt, o.get("dt"), 0.0)), "%s it %d dt: %s" % (name, it, mismatch(np.where(have_dt, dt_out, 0.0), np.where(have_dt, o.get("dt"), 0.0)))
    o.close()


@pytest.mark.parametrize("name,xb,yb,gvisc,nx,ny", CASES, ids=[c[0] for c in CASES])
def test_strip_kernel_thread_mapping_visits_every_evolved_ghost_cell_once(lib, name, xb, yb, gvisc, nx, ny):
    """thread_cell / thread_owns (what k_moc_save and k_moc_stage index with): every cell some open_moc side evolves is owned by exactly one
    thread (corner cells belong to two sides), no other cell by any, and threads past the end map to nothing."""
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="euler", density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    _, owned, _ = host_terms(lib, o, xb, yb, 0.0, s["ion_mass"], s["adiabatic_index"])
    visits = np.zeros((nx, ny), dtype=np.int32)
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    T = lib.moc_host_thread_visits(C.c_int(nx), C.c_int(ny), bc, visits.ctypes.data_as(C.c_void_p))
    assert T == 4 * (nx + ny)
    assert np.array_equal(visits, owned.astype(np.int32))
    o.close()


@pytest.mark.parametrize("n_ranks", [2, 3, 5])
@pytest.mark.parametrize("name,xb,yb,gvisc,nx,ny", CASES, ids=[c[0] for c in CASES])
def test_slab_decomposed_evaluation_equals_oracle_rhs(lib, name, xb, yb, gvisc, nx, ny, n_ranks):
    """The slab form of the open_moc path: each rank holds its rows plus two halo rows (NaN beyond a physical x boundary), addresses them by
    global row through shifted pointers as moc_field() does, and runs the strip kernels' thread loop.  Every evolved ghost cell must be owned
    by exactly one thread of exactly one rank, written only by the rank that holds it, and carry the oracle's right-hand side bit for bit."""
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o.set_global_viscosity(gvisc)
    for _ in range(2):
        o.step()
    xl, xu, yl, yu = dt_bounds(xb, yb, nx, ny)
    dxp, dyp, dt = o.get("d_x"), o.get("d_y"), o.get("dt")
    visc = gvisc * 0.5 * float(np.min(((1.0 / (1.0 / (dxp * dxp) + 1.0 / (dyp * dyp))) / dt)[xl:xu + 1, yl:yu + 1]))
    k_ref = o.rhs()
    _, owned, _ = host_terms(lib, o, xb, yb, visc, s["ion_mass"], s["adiabatic_index"])
    names = ["n", "mom_x", "mom_y", "mom_z", "thermal_energy", "bi_x", "bi_y", "bi_z", "be_x", "be_y", "be_z", "grav_x", "grav_y"]
    planes = [np.ascontiguousarray(o.get(v), dtype=np.float64) for v in names]
    dx = np.ascontiguousarray(dxp[:, 0]); dy = np.ascontiguousarray(dyp[0, :])
    arr = (C.c_void_p * 13)(*[p.ctypes.data for p in planes])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    k = np.zeros((8, nx, ny)); visits = np.zeros((nx, ny), dtype=np.int32)
    lib.moc_host_terms_slabs.restype = C.c_int
    rc = lib.moc_host_terms_slabs(arr, dx.ctypes.data_as(C.c_void_p), dy.ctypes.data_as(C.c_void_p), C.c_int(nx), C.c_int(ny), bc, C.c_double(s["ion_mass"]),
                                  C.c_double(s["adiabatic_index"]), C.c_double(visc), C.c_int(n_ranks), k.ctypes.data_as(C.c_void_p), visits.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert np.array_equal(visits, owned.astype(np.int32))
    for v, nm in enumerate(EVOLVED):
        ref = np.where(owned, k_ref[v], 0.0)
        assert same_bits(np.where(owned, k[v], 0.0), ref), "%s %d ranks d(%s)/dt: %s" % (name, n_ranks, nm, mismatch(np.where(owned, k[v], 0.0), ref))
    o.close()


@pytest.mark.parametrize("nx,ny", [(26, 23), (21, 25), (24, 24)])
@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("fixed", "open_moc")), (("open_moc", "open_moc"), ("open_moc", "open_moc")), (("open_moc", "reflect"), ("open_moc", "open"))])
def test_product_moc_limiters_equal_oracle(lib, xb, yb, nx, ny):
    """limit_line (moc_b_limiting / moc_mom_limiting, idealmhd.cpp:107-223) on random planes against the oracle's limiter passes: sides in order, corner
    cells clamped twice, negative references, and the reference's y_bound_2 index quirk (grids with xdim-4 inside and outside the clamped columns)."""
    if nx - 4 >= ny:
        pytest.skip("the reference aborts on this shape")
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="euler", density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    lim = dict(b_limiting=True, b_lower=0.7, b_upper=1.2, mom_limiting=True, mom_lower=-0.3, mom_upper=1.5)
    o.set_moc_limiting(**lim)
    rng = np.random.default_rng(nx * 100 + ny)
    names = ["mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z"]
    for v in names:
        o.view(v)[...] = rng.standard_normal((nx, ny)) * (1.0e-3 if v.startswith("mom") else 5.0)
    mut = [o.get(v) for v in names]
    be = [np.ascontiguousarray(o.get(v)) for v in ("be_x", "be_y", "be_z")]
    before = [m.copy() for m in mut]
    o.apply_moc_thresholding()
    pm = (C.c_void_p * 6)(*[m.ctypes.data for m in mut]); pb = (C.c_void_p * 3)(*[b.ctypes.data for b in be])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    lib.moc_host_limit(pm, pb, C.c_int(nx), C.c_int(ny), bc, C.c_int(1), C.c_double(lim["b_lower"]), C.c_double(lim["b_upper"]),
                       C.c_int(1), C.c_double(lim["mom_lower"]), C.c_double(lim["mom_upper"]))
    changed = 0
    for v, m, b0 in zip(names, mut, before):
        assert same_bits(m, o.get(v)), "%s: %s" % (v, mismatch(m, o.get(v)))
        changed += int((m != b0).sum())
    assert changed > 0
    o.close()
