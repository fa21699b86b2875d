"""IdealMHD2E on the host through the PRODUCT'S OWN launch code: mhd2e_host.cuh's kernels and its orchestration (E2DeviceExec::stage, e2_finish,
e2_launch_propagate, e2_enqueue_step, e2_geometry ...) are cut from the source, `k<<<grid, block, 0, stream>>>(args)` is rewritten textually to a loop
over blocks and threads, and the result is compiled with g++ over a minimal stand-in for spruce_domain.  Whole time steps must equal the CPU restatement
(pinned to the reference) bit for bit: step sizes, the seven evolved planes, the dt plane.  tests/test_mhd2e_host_check.py proves the per-cell functions
and the stage order; this adds the kernels' thread mapping, argument plumbing, set rotation and the launch sequence itself."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import mismatch, same_bits
from oracle.oracle import EVOLVED_2E, Oracle2E
from test_module_kernels_emulated import BLOCK_MIN, PRELUDE, cut
from test_oracle_vs_live_reference import e2_cases, e2_state

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "spruce_b200" / "csrc"
BUILD = ROOT / "tests" / "hostcheck" / "_build"
LIB = BUILD / "libkernel_emu_2e.so"
BC = {"periodic": 0, "open": 1, "fixed": 2, "reflect": 3, "open_moc": 4, "open_ucnp": 5}
TI = {"euler": 0, "rk2": 1, "rk4": 2}

DOMAIN = r'''
#include "spruce_b200.h"
#include <cstdlib>
#include <utility>
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
template <class K, class... A>
static void launch3(K k, dim3 g, unsigned bx, A... a)
{
    gridDim = {g.x, g.y, 1}; blockDim = {bx, 1, 1};
    for (unsigned y = 0; y < g.y; y++) for (unsigned x = 0; x < g.x; x++) for (unsigned t = 0; t < bx; t++) { blockIdx = {x, y, 0}; threadIdx = {t, 0, 0}; k(a...); }
}
#define CUDA_TRY(x) do { (void)(x); } while (0)
static inline int cudaGetLastError() { return 0; }
using namespace spruce;
struct OneFluid2E;
// the members of capi.cu's spruce_domain that the IdealMHD2E code touches
struct spruce_domain {
    spruce_config cfg; DomainParams P; double *stat[NSTATIC] = {nullptr}; StepCtl *ctl = nullptr; double *dt_hist = nullptr; int stream = 0; long long launches = 0;
    OneFluid2E *e2 = nullptr; std::vector<double> dxg, dyg;
};
static int fail(int code, const char *, ...) { return code; }
static int alloc_plane(spruce_domain *d, double **out)
{
    const int halo = d->cfg.n_ranks > 1 ? HALO : 0;                  // slabs carry two halo rows on each side
    double *b = (double *)std::calloc((size_t)(d->P.nx + 2 * halo) * d->P.pitch, sizeof(double));
    *out = b ? b + (size_t)halo * d->P.pitch : nullptr;
    return b ? SPRUCE_OK : SPRUCE_ERR_CUDA;
}
// the peer transport between slabs, for ranks that are host threads of this process (as in tests/test_capi_hooks_emulated.py)
#include <condition_variable>
#include <mutex>
#include <thread>
struct Barrier {
    std::mutex m; std::condition_variable cv; int n = 1, count = 0, gen = 0;
    void wait() { std::unique_lock<std::mutex> l(m); const int g = gen; if (++count == n) { gen++; count = 0; cv.notify_all(); } else cv.wait(l, [&] { return g != gen; }); }
};
static Barrier g_bar;
static int g_world = 1;
static std::vector<double> g_edge[8][2];
static unsigned long long g_dt[8];
static int peer_exchange(spruce_domain *d, double *const *U, void *)
{
    if (g_world == 1) return SPRUCE_OK;
    const int r = d->cfg.rank, W = g_world;
    const size_t rows = (size_t)HALO * d->P.pitch;
    g_edge[r][0].resize(NEV * rows); g_edge[r][1].resize(NEV * rows);
    for (int v = 0; v < NEV; v++) {
        std::memcpy(&g_edge[r][0][v * rows], U[v], rows * sizeof(double));
        std::memcpy(&g_edge[r][1][v * rows], U[v] + (size_t)(d->P.nx - HALO) * d->P.pitch, rows * sizeof(double));
    }
    g_bar.wait();
    const int lo = r > 0 ? r - 1 : (d->P.xper ? W - 1 : -1), hi = r < W - 1 ? r + 1 : (d->P.xper ? 0 : -1);
    for (int v = 0; v < NEV; v++) {
        if (lo >= 0) std::memcpy(U[v] - rows, &g_edge[lo][1][v * rows], rows * sizeof(double));
        if (hi >= 0) std::memcpy(U[v] + (size_t)d->P.nx * d->P.pitch, &g_edge[hi][0][v * rows], rows * sizeof(double));
    }
    g_bar.wait();
    return SPRUCE_OK;
}
static int peer_dt_allgather(spruce_domain *d)
{
    if (g_world == 1) return SPRUCE_OK;
    g_dt[d->cfg.rank] = d->ctl->dtmin_bits;
    g_bar.wait();
    for (int r = 0; r < g_world; r++) d->ctl->dtmin_bits = std::min(d->ctl->dtmin_bits, g_dt[r]);
    g_bar.wait();
    return SPRUCE_OK;
}
'''


# stand-ins only this set's launch code needs (the viscosity terms upload a profile and copy a plane); tests/test_ideal2f_kernels_emulated.py has its own
E2_RUNTIME = r'''
static int h2d_plane(spruce_domain *d, double *dev, const double *host) { std::memcpy(dev, host, (size_t)d->P.nx * d->P.pitch * sizeof(double)); return SPRUCE_OK; }      // pitch == ny here
enum { cudaMemcpyDeviceToDevice = 3 };
static inline int cudaMemcpyAsync(void *dst, const void *src, size_t n, int, int) { std::memcpy(dst, src, n); return 0; }
'''


def assemble():
    mk = (CSRC / "mhd_kernels.cuh").read_text()
    ca = (CSRC / "capi.cu").read_text()
    e2 = (CSRC / "mhd2e_host.cuh").read_text()
    body = cut(e2, "struct E2Args {", "int e2_upload(") + cut(e2, "// after spruce_eqs_setup on every rank of a decomposed run", "\n}\n") + "\n}\n"
    body, n = re.subn(r"(\w+)<<<(.+?), (\w+), 0, d->stream>>>\(", r"launch3(\1, \2, \3, ", body)
    assert n >= 6 and "<<<" not in body
    return "".join([PRELUDE, '#include "mhd2e_cells.cuh"\n#include "mhd2e_step.hpp"\nnamespace spruce {\n',
                    cut(mk, "constexpr int HALO", "enum { KM_NONE", include_end=True), BLOCK_MIN,
                    cut(mk, "struct StepCtl {", "// the rare fallback of the skip test"),
                    cut(ca, "struct HostAxis {", "struct TwoFluid;"), cut(ca, "void build_axis(", "int upload_tables("),
                    "}  // namespace spruce\n", DOMAIN, E2_RUNTIME, body, (ROOT / "tests" / "hostcheck" / "kernel_emu_2e.inc").read_text()])


@pytest.fixture(scope="module")
def emu():
    BUILD.mkdir(exist_ok=True)
    src = BUILD / "kernel_emu_2e.cpp"
    text = assemble()
    if not LIB.exists() or not src.exists() or src.read_text() != text:
        LIB.unlink(missing_ok=True)                # a failed compile must not leave the previous library behind
        src.write_text(text)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-I", str(CSRC), "-I", str(ROOT / "include"), "-o", str(LIB), str(src)], check=True)
    L = C.CDLL(str(LIB))
    L.emu2e_run.restype = C.c_int
    L.emu2e_set_eic(0)
    return L


@pytest.mark.parametrize("world", [1, 2, 3])                          # 2, 3: slabs as host threads, the peer transport a staging copy between them
@pytest.mark.parametrize("k,xb,yb,integrator,nx,ny,loop,nmin", e2_cases())
def test_product_mhd2e_launch_code_runs_whole_steps_equal_to_oracle(emu, k, xb, yb, integrator, nx, ny, loop, nmin, world):
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
    out = np.zeros((7, nx, ny)); dt = np.zeros((nx, ny)); steps = np.zeros(nsteps)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = emu.emu2e_run(C.c_int(world), arr, vp(dx), vp(dy), C.c_int(nx), C.c_int(ny), bc, C.c_int(TI[integrator]), C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]), C.c_double(0.2),
                       C.c_double(floors["density_min"]), C.c_double(floors["temp_min"]), C.c_double(floors["thermal_energy_min"]), C.c_double(1.0), C.c_double(0.5), C.c_int(nsteps),
                       vp(out), vp(dt), vp(steps))
    assert rc == 0
    assert [float(x).hex() for x in steps] == [x.hex() for x in ref_steps]
    for v, nm in enumerate(EVOLVED_2E):
        assert same_bits(out[v], o.get(nm)), "case %d %s: %s" % (k, nm, mismatch(out[v], o.get(nm)))
    assert same_bits(dt, o.get("dt")), "case %d dt: %s" % (k, mismatch(dt, o.get("dt")))
    o.close()


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("name,xb,yb,integrator,nx,ny,drift,bfield", [
    ("ucnp_sides_rk2", ("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp"), "rk2", 27, 25, 20.0, 0.01),
    ("mixed_walls_rk4", ("fixed", "reflect"), ("open_ucnp", "fixed"), "rk4", 24, 29, 10.0, 0.02),
    ("periodic_euler", ("periodic", "periodic"), ("periodic", "periodic"), "euler", 22, 21, 0.0, 0.01)])
def test_product_mhd2e_launch_code_with_eic_thermalization(emu, name, xb, yb, integrator, nx, ny, drift, bfield, world):
    """ideal_mhd_2E + eic_thermalization through the product's kernels and launch code (Geo::eic as spruce_module_eic_thermalization sets it), one rank and slabs:
    within the module's 1e-9 of the restatement that live reference runs pin, and visibly different from the run without the module."""
    from spruce_b200 import synthetic
    s = synthetic.ucnp_cloud_2e(nx, ny, drift=drift, bfield=bfield)
    floors = dict(density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1.0e-30)
    o = Oracle2E(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator=integrator, eic=True, **floors)
    plain = Oracle2E(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator=integrator, **floors)
    nsteps = 5
    ref_steps = np.array([o.step() for _ in range(nsteps)])
    for _ in range(nsteps):
        plain.step()
    names = ["rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "be_x", "be_y", "grav_x", "grav_y"]
    planes = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in names]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    arr = (C.c_void_p * 11)(*[p.ctypes.data for p in planes])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    out = np.zeros((7, nx, ny)); dt = np.zeros((nx, ny)); steps = np.zeros(nsteps)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.emu2e_set_eic(1)
    try:
        rc = emu.emu2e_run(C.c_int(world), arr, vp(dx), vp(dy), C.c_int(nx), C.c_int(ny), bc, C.c_int(TI[integrator]), C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]), C.c_double(0.2),
                           C.c_double(floors["density_min"]), C.c_double(floors["temp_min"]), C.c_double(floors["thermal_energy_min"]), C.c_double(1.0), C.c_double(0.5), C.c_int(nsteps),
                           vp(out), vp(dt), vp(steps))
    finally:
        emu.emu2e_set_eic(0)
    assert rc == 0
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))
    assert np.max(np.abs(steps - ref_steps) / ref_steps) <= 1e-9
    for v, nm in enumerate(EVOLVED_2E):
        assert rel(out[v], o.get(nm)) <= 1e-9, "%s %s: %.3e" % (name, nm, rel(out[v], o.get(nm)))
    assert rel(dt, o.get("dt")) <= 1e-9
    for nm in ("i_thermal_energy", "e_thermal_energy"):
        v = EVOLVED_2E.index(nm)
        assert rel(out[v], plain.get(nm)) > 100 * max(rel(out[v], o.get(nm)), 1e-15), "%s: the exchange term is not visible in %s" % (name, nm)
    o.close(); plain.close()


E2_VISC_EMU_CASES = [
    ("rhs_terms_rk2_eic", ("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp"), "rk2", 27, 25, True,
     [("local", 0.5, "v_x", "mom_x", 0.0, "i"), ("global", 0.3, "v_y", "mom_y", 0.0, "i"), ("local", 0.4, "i_temp", "i_thermal_energy", 0.0, "i"), ("global", 0.2, "e_temp", "e_thermal_energy", 0.0, "e")], "euler", False),
    ("boundary_gc_hv_rk2", ("fixed", "reflect"), ("open_ucnp", "fixed"), "rk4", 24, 29, False,
     [("boundary", 0.8, "v_x", "mom_x", 0.3, "i"), ("global", 3.0, "v_y", "mom_y", 0.0, "i"), ("boundary_global", 0.6, "e_temp", "e_thermal_energy", 0.2, "e")], "rk2", True),
    ("hv_rk4_periodic", ("periodic", "periodic"), ("periodic", "periodic"), "euler", 22, 21, False,
     [("local", 2.5, "i_temp", "i_thermal_energy", 0.0, "i"), ("local", 0.4, "rho", "rho", 0.0, "i")], "rk4", False),
    ("hv_euler_walls", ("reflect", "open"), ("fixed", "open"), "rk2", 26, 23, False,
     [("global", 2.0, "v_x", "mom_x", 0.0, "i"), ("local", 0.7, "e_temp", "e_thermal_energy", 0.0, "e")], "euler", True),
]


@pytest.mark.parametrize("name,xb,yb,integrator,nx,ny,eic,terms,hv_integ,gc", E2_VISC_EMU_CASES, ids=[c[0] for c in E2_VISC_EMU_CASES])
def test_product_mhd2e_launch_code_with_artificial_viscosity(emu, name, xb, yb, integrator, nx, ny, eic, terms, hv_integ, gc):
    """artificial_viscosity on ideal_mhd_2E (the fourth module of the UCNP set) through the product's kernels and launch code -- visc_cell / k_2e_visc_term, the right-hand-side
    terms added by k_2e_cells, the hyper-viscous sub-steps of e2_av_iterate with a propagate and a refreshed primary dt plane after each -- whole steps bit-equal to the
    restatement that live reference runs pin (test_ideal_mhd_2e_with_artificial_viscosity_oracle_equals_live_reference); 1e-9 with eic_thermalization in the mix"""
    from spruce_b200 import synthetic
    from golden_util import boundary_viscosity_profile
    s = synthetic.ucnp_cloud_2e(nx, ny, drift=20.0, bfield=0.01)
    floors = dict(density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1.0e-30)
    o = Oracle2E(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator=integrator, eic=eic, **floors)
    full = [dict(opt=t[0], strength=t[1], var_diff=t[2], var_evol=t[3], species=t[5],
                 strength_grid=boundary_viscosity_profile(s["planes"]["pos_x"], s["planes"]["pos_y"], t[1], t[4]) if t[0].startswith("boundary") else None) for t in terms]
    o.set_viscosity(full, hv_integrator=hv_integ, hv_epsilon=1.0, gradient_correction=gc)
    plain = Oracle2E(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator=integrator, eic=eic, **floors)
    nsteps = 4
    ref_steps = np.array([o.step() for _ in range(nsteps)])
    for _ in range(nsteps):
        plain.step()
    names = ["rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "be_x", "be_y", "grav_x", "grav_y"]
    planes = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in names]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    arr = (C.c_void_p * 11)(*[p.ctypes.data for p in planes])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    out = np.zeros((7, nx, ny)); dt = np.zeros((nx, ny)); steps = np.zeros(nsteps)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.emu2e_set_eic(int(eic))
    emu.emu2e_viscosity_begin(C.c_int(TI[hv_integ]), C.c_int(int(gc)))
    keep = []
    for t in full:
        prof = np.ascontiguousarray(t["strength_grid"], dtype=np.float64) if t["strength_grid"] is not None else None
        keep.append(prof)
        emu.emu2e_viscosity_term(t["opt"].encode(), C.c_double(t["strength"]), t["var_diff"].encode(), t["var_evol"].encode(), t["species"].encode(),
                                 vp(prof) if prof is not None else None, C.c_int(nx * ny))
    try:
        rc = emu.emu2e_run(C.c_int(1), arr, vp(dx), vp(dy), C.c_int(nx), C.c_int(ny), bc, C.c_int(TI[integrator]), C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]), C.c_double(0.2),
                           C.c_double(floors["density_min"]), C.c_double(floors["temp_min"]), C.c_double(floors["thermal_energy_min"]), C.c_double(1.0), C.c_double(0.5), C.c_int(nsteps),
                           vp(out), vp(dt), vp(steps))
    finally:
        emu.emu2e_set_eic(0)
        emu.emu2e_viscosity_begin(C.c_int(0), C.c_int(0))
    assert rc == 0
    if eic:
        rel = lambda a, b: float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))
        assert np.max(np.abs(steps - ref_steps) / ref_steps) <= 1e-9
        for v, nm in enumerate(EVOLVED_2E):
            assert rel(out[v], o.get(nm)) <= 1e-9, "%s %s: %.3e" % (name, nm, rel(out[v], o.get(nm)))
    else:
        assert [float(x).hex() for x in steps] == [float(x).hex() for x in ref_steps]
        for v, nm in enumerate(EVOLVED_2E):
            assert same_bits(out[v], o.get(nm)), "%s %s: %s" % (name, nm, mismatch(out[v], o.get(nm)))
        assert same_bits(dt, o.get("dt"))
    assert any(not same_bits(out[v], plain.get(nm)) for v, nm in enumerate(EVOLVED_2E)), "the viscosity never acted"
    o.close(); plain.close()
