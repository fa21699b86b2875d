"""The two-fluid set on the host through the PRODUCT'S OWN kernels and launch code (ideal2f_kernels.cuh, ideal2f_host.cuh, ideal2f_sides.cuh; launches rewritten
to block / thread loops as in tests/test_mhd2e_kernels_emulated.py).  The point is the launch sequence written after the GPU time was spent: an open_ucnp side
next to fixed / reflect sides runs the stage without its primary tail, then the literal ordered side passes, the floors kernel and a full dt pass.  The
GPU-validated sequences (all-periodic / ucnp-only / wall-only) run too, as a check of the emulation itself.  Everything bit for bit against the restatement."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import mismatch, same_bits
from oracle.oracle import EVOLVED_2F, Oracle2F
from spruce_b200 import synthetic
from test_mhd2e_kernels_emulated import BC, DOMAIN, TI
from test_module_kernels_emulated import BLOCK_MIN, PRELUDE, cut

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "spruce_b200" / "csrc"
BUILD = ROOT / "tests" / "hostcheck" / "_build"
LIB = BUILD / "libkernel_emu_2f.so"

EXTRA_DOMAIN = r'''
static int static_slot(const char *name)
{
    if (!strcmp(name, "be_x")) return S_BEX; if (!strcmp(name, "be_y")) return S_BEY; if (!strcmp(name, "be_z")) return S_BEZ;
    if (!strcmp(name, "grav_x")) return S_GX; if (!strcmp(name, "grav_y")) return S_GY;
    return -1;
}
static int h2d_plane(spruce_domain *d, double *dev, const double *host) { std::memcpy(dev, host, sizeof(double) * (size_t)d->P.nx * d->P.ny); return SPRUCE_OK; }      // local rows only
static int d2h_plane(spruce_domain *d, double *host, const double *dev) { std::memcpy(host, dev, sizeof(double) * (size_t)d->P.nx * d->P.ny); return SPRUCE_OK; }
static int exchange_plane(spruce_domain *, double *) { return SPRUCE_OK; }
'''


def assemble():
    mk = (CSRC / "mhd_kernels.cuh").read_text()
    mo = (CSRC / "module_kernels.cuh").read_text()
    ca = (CSRC / "capi.cu").read_text()
    tk = (CSRC / "ideal2f_kernels.cuh").read_text()
    th = (CSRC / "ideal2f_host.cuh").read_text()
    e2inc = (ROOT / "tests" / "hostcheck" / "kernel_emu_2e.inc").read_text()
    body = cut(th, "struct TfSideArgs {", "int tf_time_derivatives(")      # includes tf_initial_exchange, tf_upload, tf_download
    body, n = re.subn(r"(\w+)<<<(.+?), (\w+), 0, d->stream>>>\(", r"launch3(\1, \2, \3, ", body)
    assert n >= 8 and "<<<" not in body
    domain = DOMAIN.replace("struct OneFluid2E;", "struct OneFluid2E;\nstruct TwoFluid;").replace(
        "OneFluid2E *e2 = nullptr;", "OneFluid2E *e2 = nullptr; TwoFluid *tf = nullptr; double *scratch_out = nullptr; bool is_setup = false, any_ucnp = false, any_primary_ghost = false, fast_interior = true;")
    return "".join([PRELUDE, '#include "ideal2f_sides.cuh"\nnamespace spruce {\n',
                    cut(mk, "constexpr int HALO", "enum { KM_NONE", include_end=True),
                    cut(mk, "__device__ __forceinline__ FaceGeom load_face_geom", "// is global row g / column j inside"),
                    BLOCK_MIN,
                    cut(mk, "// is global row g / column j inside", "// block-wide NaN-ignoring minimum"),
                    cut(mk, "struct StepCtl {", "// the rare fallback of the skip test"),
                    cut(mo, "constexpr double kKappa0", "struct TcParams {"),
                    "}  // namespace spruce\n",
                    tk[tk.index('#include "module_kernels.cuh"') + len('#include "module_kernels.cuh"'):],
                    "namespace spruce {\n", cut(ca, "struct HostAxis {", "struct TwoFluid;"), cut(ca, "void build_axis(", "int upload_tables("), "}  // namespace spruce\n",
                    domain, EXTRA_DOMAIN, body, cut(e2inc, "static void emu_axis(", "// in: GLOBAL planes"),
                    (ROOT / "tests" / "hostcheck" / "kernel_emu_2f.inc").read_text()])


@pytest.fixture(scope="module")
def emu():
    BUILD.mkdir(exist_ok=True)
    src = BUILD / "kernel_emu_2f.cpp"
    text = assemble()
    if not LIB.exists() or not src.exists() or src.read_text() != text:
        LIB.unlink(missing_ok=True)                # a failed compile must not leave the previous library behind
        src.write_text(text)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-I", str(CSRC), "-I", str(ROOT / "include"), "-o", str(LIB), str(src)], check=True)
    L = C.CDLL(str(LIB))
    L.emu2f_run.restype = C.c_int
    return L


CASES = [
    # launch sequences written after the GPU time was spent: open_ucnp next to wall-type sides (ordered side passes)
    (("open_ucnp", "open_ucnp"), ("reflect", "reflect"), "rk2", 37, 31, True),
    (("reflect", "open_ucnp"), ("fixed", "open_ucnp"), "rk4", 30, 35, True),
    (("open_ucnp", "fixed"), ("open_ucnp", "reflect"), "euler", 34, 29, True),
    (("periodic", "periodic"), ("open_ucnp", "reflect"), "rk2", 28, 36, True),
    # GPU-validated sequences, as a check of the emulation
    (("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp"), "rk2", 33, 30, False),
    (("periodic", "periodic"), ("periodic", "periodic"), "rk4", 26, 24, False),
    (("fixed", "reflect"), ("periodic", "periodic"), "euler", 27, 25, False),
]


@pytest.mark.parametrize("xb,yb,integ,nx,ny,ordered", CASES)
@pytest.mark.parametrize("eic", [False, True])
@pytest.mark.parametrize("world", [1, 2, 3])                          # 2, 3: slabs as host threads, the peer transport a staging copy between them
def test_two_fluid_launch_code_runs_whole_steps_equal_to_oracle(emu, xb, yb, integ, nx, ny, ordered, eic, world):
    s = synthetic.ucnp_cloud(nx, ny, drift=2.0e3, bfield=5.0)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1e-30)
    o = Oracle2F(s["planes"], s["ion_mass"], s["adiabatic_index"], remove_curl_terms=False, eic=eic, **kw)
    nsteps = 5
    ref = [o.step() for _ in range(nsteps)]
    up = [k for k in s["planes"] if k not in ("d_x", "d_y", "pos_x", "pos_y")]
    planes = [np.ascontiguousarray(s["planes"][k], dtype=np.float64) for k in up]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    outs = EVOLVED_2F + ["dt", "dt_i", "e_temp"]
    out = np.zeros((len(outs), nx, ny)); steps = np.zeros(nsteps); flag = C.c_int()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu.emu2f_run(C.c_int(world), (C.c_char_p * len(up))(*[k.encode() for k in up]), (C.c_void_p * len(up))(*[p.ctypes.data for p in planes]), C.c_int(len(up)), vp(dx), vp(dy), C.c_int(nx), C.c_int(ny),
                       (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]]), C.c_int(TI[integ]), C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]), C.c_double(0.2),
                       C.c_double(1.0), C.c_double(1.0e-3), C.c_double(1e-30), C.c_int(0), C.c_int(int(eic)), C.c_int(nsteps),
                       (C.c_char_p * len(outs))(*[k.encode() for k in outs]), C.c_int(len(outs)), vp(out), vp(steps), C.byref(flag))
    assert rc == 0, rc
    assert bool(flag.value) == ordered
    assert [float(x).hex() for x in steps] == [x.hex() for x in ref]
    for k, v in enumerate(outs):
        assert same_bits(out[k], o.get(v)), "%s: %s" % (v, mismatch(out[k], o.get(v)))
    o.close()
