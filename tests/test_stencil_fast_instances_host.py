"""The FAST instances of the module stencil helpers (spruce_b200/csrc/module_kernels.cuh: deep_interior, rdT / Dx / Dy / D2x / D2y <true>) in k_pv_stage, executed on
the host from the kernel's own source (assembled like tests/test_module_kernels_emulated.py): with the fast instances on, every output plane of a physical-viscosity
stage -- thermal energy, momenta, stage velocities and temperature -- must equal the all-general run bit for bit, for wall and periodic sides, heating and force, with
and without the gradient correction.  (The general instances are the GPU-validated ones; k_tc_stage and k_2f_stage have the same check in
tests/test_capi_hooks_emulated.py and tests/test_ideal2f_kernels_emulated.py, against the oracle.)"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import same_bits
from spruce_b200 import synthetic
from test_module_kernels_emulated import BC, BLOCK_MIN, BUILD, CSRC, PRELUDE, ROOT, cut, vp

LIB = BUILD / "libstencil_fast.so"


@pytest.fixture(scope="module")
def emu():
    mk = (CSRC / "mhd_kernels.cuh").read_text()
    mo = (CSRC / "module_kernels.cuh").read_text()
    ca = (CSRC / "capi.cu").read_text()
    main = (ROOT / "tests" / "hostcheck" / "kernel_emu_main.inc").read_text()
    text = "".join([PRELUDE, "#include <vector>\nnamespace spruce {\n",
                    cut(mk, "constexpr int HALO", "enum { KM_NONE", include_end=True),
                    cut(mk, "__device__ __forceinline__ FaceGeom load_face_geom", "// is global row g / column j inside"),
                    BLOCK_MIN,
                    cut(mk, "// is global row g / column j inside", "// block-wide NaN-ignoring minimum"),
                    cut(mk, "struct PropArgs {", "// Ghost cells of the non-periodic sides"), "\n",
                    cut(mk, "struct StepCtl {", "// begin: step = epsilon"),
                    cut(mo, "constexpr double kKappa0", "}  // namespace spruce"),
                    cut(ca, "struct HostAxis {", "struct TwoFluid;"),
                    cut(ca, "void build_axis(", "int upload_tables("),
                    "}  // namespace spruce\n",
                    cut(main, "template <class K, class... A>", "// launch_propagate(d, 0) of capi.cu"),
                    (ROOT / "tests" / "hostcheck" / "stencil_fast_check.inc").read_text()])
    BUILD.mkdir(exist_ok=True)
    src = BUILD / "stencil_fast.cpp"
    if not LIB.exists() or not src.exists() or src.read_text() != text:
        LIB.unlink(missing_ok=True)                # a failed compile must not leave the previous library behind
        src.write_text(text)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(CSRC), "-o", str(LIB), str(src)], check=True)
    L = C.CDLL(str(LIB))
    L.emu_create.restype = C.c_void_p
    return L


BOUNDS = [(("fixed", "open"), ("reflect", "open")), (("periodic", "periodic"), ("fixed", "open")), (("open", "open"), ("periodic", "periodic")), (("periodic", "periodic"), ("periodic", "periodic"))]


@pytest.mark.parametrize("heating,force,gc,half,final", [(1, 1, 0, 1.0, 1), (1, 1, 1, 0.5, 0), (0, 1, 1, 1.0, 1), (1, 0, 0, 1.0, 1)])
@pytest.mark.parametrize("xb,yb", BOUNDS)
def test_physical_viscosity_stage_fast_instances_equal_the_general_ones(emu, xb, yb, heating, force, gc, half, final):
    nx, ny = 29, 26
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    P = s["planes"]
    rng = np.random.default_rng(5)
    dx = np.ascontiguousarray(P["d_x"][:, 0]); dy = np.ascontiguousarray(P["d_y"][0, :])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    n = np.ascontiguousarray(P["rho"] / s["ion_mass"])
    zeros = np.zeros((nx, ny))
    ev = [n, zeros, zeros, zeros, zeros, zeros, zeros, zeros]
    st = [np.ascontiguousarray(P[k], dtype=np.float64) for k in ("be_x", "be_y", "be_z", "grav_x", "grav_y")]
    h = C.c_void_p(emu.emu_create(C.c_int(nx), C.c_int(ny), bc, C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]), C.c_double(1.0e7), C.c_double(1.0e4), C.c_double(1.0e-6),
                                  C.c_double(0.2), vp(dx), vp(dy), (C.c_void_p * 8)(*[a.ctypes.data for a in ev]), (C.c_void_p * 5)(*[a.ctypes.data for a in st])))
    v = [np.ascontiguousarray(1.0e6 * rng.standard_normal((nx, ny))) for _ in range(3)]
    T = np.ascontiguousarray(P["temp"] * (1.0 + 0.3 * rng.random((nx, ny))))
    b = rng.standard_normal((3, nx, ny)); b /= np.sqrt((b * b).sum(axis=0))
    bh = [np.ascontiguousarray(b[k]) for k in range(3)]
    cg = np.ascontiguousarray(1.0e-16 * (0.2 + rng.random((nx, ny))))
    e0 = np.ascontiguousarray(n * 1.3807e-16 * T / (s["adiabatic_index"] - 1.0))
    mom0 = [np.ascontiguousarray(P["rho"] * v[k]) for k in range(3)]
    ptrs = lambda arrs: (C.c_void_p * 3)(*[a.ctypes.data for a in arrs])
    res = []
    for fast in (1, 0):
        e = e0.copy(); mom = [m.copy() for m in mom0]; v_out = [np.zeros((nx, ny)) for _ in range(3)]; T_out = np.zeros((nx, ny))
        emu.emu_pv_stage(h, C.c_int(fast), ptrs(v), vp(T), ptrs(bh), vp(n), vp(cg), C.c_double(1.0e-16), C.c_double(0.05), C.c_double(half), C.c_int(heating), C.c_int(force), C.c_int(gc),
                         C.c_int(final), vp(e), ptrs(mom), ptrs(v_out), vp(T_out))
        res.append([e, *mom, *v_out, T_out])
    for a, b2 in zip(res[0], res[1]):
        assert same_bits(a, b2)
    changed = sum(not np.array_equal(a, b0) for a, b0 in zip(res[0][:4], [e0, *mom0]))
    assert changed >= (1 if final else 0) and not np.array_equal(res[0][7], T) or not heating
    emu.emu_destroy(h)
