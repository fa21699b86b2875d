"""Module hooks through capi.cu's OWN functions on the host: the real `struct spruce_domain`, launch_propagate / launch_ghosts / derive_to, dc_post, fh_pre + fh_iterate
and anomres_host.cuh (ArExec, ar_setup_run, ar_iterate) are cut from the product source, kernel launches rewritten to block / thread loops, the few CUDA runtime
calls replaced by memcpy-style stand-ins.  One module iteration INCLUDING its propagateChanges (floors, boundary passes of open / reflect sides, dt) must equal the
oracle's hook bit for bit.  tests/test_module_kernels_emulated.py runs the same kernels from a transcription of the launch order; here the order is the product's."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import mismatch, same_bits
from oracle.oracle import Oracle, anomalous_params
from spruce_b200 import synthetic
from test_module_kernels_emulated import BLOCK_MIN, EV, FLOORS, PRELUDE, ST, cut
from test_oracle_vs_live_reference import AR_CASES, ar_kwargs

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "spruce_b200" / "csrc"
BUILD = ROOT / "tests" / "hostcheck" / "_build"
LIB = BUILD / "libkernel_emu_capi.so"
BC = {"periodic": 0, "open": 1, "fixed": 2, "reflect": 3, "open_moc": 4}

RUNTIME = r'''
#include "spruce_b200.h"
#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <utility>
#include "moc_kernels.cuh"
#include "anomres_cells.hpp"
#include "solar_templates.hpp"
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
template <class K, class... A>
static void launch3(K k, dim3 g, unsigned bx, A... a)
{
    gridDim = {g.x, g.y, 1}; blockDim = {bx, 1, 1};
    for (unsigned y = 0; y < g.y; y++) for (unsigned x = 0; x < g.x; x++) for (unsigned t = 0; t < bx; t++) { blockIdx = {x, y, 0}; threadIdx = {t, 0, 0}; k(a...); }
}
// the CUDA runtime calls the cut functions make
typedef void *cudaStream_t; typedef void *cudaEvent_t; typedef int cudaError_t;
enum { cudaSuccess = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline int cudaGetLastError() { return 0; }
static inline const char *cudaGetErrorString(int) { return ""; }
static inline int cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline int cudaMemcpyAsync(void *dst, const void *src, size_t n, int, cudaStream_t) { std::memcpy(dst, src, n); return 0; }
static inline int cudaMemsetAsync(void *dst, int v, size_t n, cudaStream_t) { std::memset(dst, v, n); return 0; }
template <class T> static inline int cudaMalloc(T **p, size_t n) { *p = (T *)std::calloc(n, 1); return *p ? 0 : 2; }
static int fail(int code, const char *, ...) { return code; }
#define CUDA_TRY(x) do { if ((x) != cudaSuccess) return fail(SPRUCE_ERR_CUDA, "cuda"); } while (0)
using namespace spruce;
'''
STUBS = r'''
// the peer transport between slabs, for ranks that are host threads of this process: edge rows through a staging area, reductions through shared slots
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
static unsigned long long g_red[8][4], g_dt[8];
static int peer_exchange(spruce_domain *d, double *const *U, cudaStream_t)
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
static int peer_red_allgather(spruce_domain *d)
{
    if (g_world == 1) return SPRUCE_OK;
    for (int k = 0; k < 4; k++) g_red[d->cfg.rank][k] = d->red[k];
    g_bar.wait();
    for (int r = 0; r < g_world; r++) for (int k = 0; k < 4; k++) d->red[k] = (k & 1) ? std::max(d->red[k], g_red[r][k]) : std::min(d->red[k], g_red[r][k]);
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
int launch_moc(spruce_domain *d, const PlaneSet &S, const PlaneSet &B, const PlaneSet &D, double coef, int primary, int kmode, int dt_only);
int launch_propagate(spruce_domain *d, int from_state);
int reset_reductions(spruce_domain *d);
'''


def cut_fn(text, signature):
    """one top-level function of capi.cu: from its signature to the closing brace in column 0"""
    m = re.search(r"^" + re.escape(signature.split("\n")[0]) + r"[^;{]*\n\{", text, re.M)             # the definition, not a forward declaration
    assert m, signature
    return text[m.start():text.index("\n}\n", m.start()) + 3]


def rewrite_launches(body):
    body, n = re.subn(r"(\w+(?:<\w+>)?)<<<(.+?), (\w+), 0, d->stream>>>\(", r"launch3(\1, \2, \3, ", body)
    assert "<<<" not in body, body[body.index("<<<") - 80:body.index("<<<") + 80]
    return body


def assemble():
    mk = (CSRC / "mhd_kernels.cuh").read_text()
    mo = (CSRC / "module_kernels.cuh").read_text()
    ca = (CSRC / "capi.cu").read_text()
    ah = (CSRC / "anomres_host.cuh").read_text()
    ms = (CSRC / "moc_stage.cuh").read_text()
    fns = ["int alloc_plane(", "void build_axis(", "void build_ghost_proto(", "void fill_moc(", "int launch_moc_save(", "int launch_moc(", "int moc_limit(", "int launch_ghosts(", "int launch_propagate(spruce_domain *d, int from_state)\n{", "int derive_to(",
           "int read_reductions(", "int reset_reductions(", "double bits_to_double(", "int exchange_plane(", "int after_module_propagate(", "int ms_feed(", "int ah_post(", "int launch_op(", "int dc_post(", "int fh_pre(", "int fh_iterate(", "int src_post(", "int bo_post(", "int tc_derive(", "int tc_iterate(", "int rl_launch(", "int rl_iterate(", "bool dev_subcycles(",
           "int exchange_planes4(", "PvArgs pv_args(", "int pv_substeps(",
           "int evolved_slot(int var)\n{", "const double *materialise_var(", "int visc_needs_dt_plane(", "int visc_refresh_dt(", "int visc_term(", "int prepare_rhs_modules(", "int av_iterate("]
    one_liners = {"double bits_to_double(", "int visc_needs_dt_plane(", "int visc_refresh_dt("}
    code = []
    for f in fns:
        if f in one_liners:
            i = ca.index("\n" + f) + 1
            code.append(ca[i:ca.index("\n", i) + 1])
        else:
            code.append(cut_fn(ca, f))
    body = rewrite_launches("".join(code) + ah[ah.index("#pragma once") + len("#pragma once"):])
    domain = cut(ca, "struct spruce_domain {", "\n};\n") + "\n};\n"
    return "".join([PRELUDE, RUNTIME, "namespace spruce {\n",
                    cut(mk, "constexpr int HALO", "enum { KM_NONE", include_end=True),
                    cut(mk, "__device__ __forceinline__ FaceGeom load_face_geom", "// is global row g / column j inside"),
                    BLOCK_MIN,
                    cut(mk, "// is global row g / column j inside", "// block-wide NaN-ignoring minimum"),
                    cut(mk, "struct PropArgs {", "// Slab decomposition: pack the first/last HALO rows"), "\n",
                    "constexpr int MAX_RANKS = 16;\n",
                    cut(mk, "struct StepCtl {", "// dt all-gather over peer memory"),
                    cut(ms, "struct MocArgs {", "}  // namespace spruce"),
                    cut(mo, "constexpr double kKappa0", "// Artificial viscosity (source/modules/viscosity.cpp"), "\n",
                    cut(mo, "struct ViscArgs {", "// PlasmaDomain differential operators on a plane"),
                    cut(mo, "struct PvArgs {", "struct OpArgs"),
                    cut(mo, "struct OpArgs", "}  // namespace spruce"),
                    "}  // namespace spruce\n",
                    "struct PlaneSet { double *p[NEV] = {nullptr}; };\n", cut(ca, "struct HostAxis {", "struct TwoFluid;"), "struct TwoFluid;\nstruct OneFluid2E;\n",
                    domain, STUBS, body, (ROOT / "tests" / "hostcheck" / "kernel_emu_capi.inc").read_text()])


@pytest.fixture(scope="module")
def emu():
    BUILD.mkdir(exist_ok=True)
    src = BUILD / "kernel_emu_capi.cpp"
    text = assemble()
    if not LIB.exists() or not src.exists() or src.read_text() != text:
        LIB.unlink(missing_ok=True)                # a failed compile must not leave the previous library behind
        src.write_text(text)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-I", str(CSRC), "-I", str(ROOT / "include"), "-o", str(LIB), str(src)], check=True)
    L = C.CDLL(str(LIB))
    L.cemu_create.restype = C.c_void_p
    L.cemu_create_slab.restype = C.c_void_p
    L.cemu_dtmin.restype = C.c_double
    return L


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def make_pair(emu, xb, yb, nx, ny, integrator="rk2", warm=2):
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator=integrator, **FLOORS)
    o.run(warm)
    ev = [np.ascontiguousarray(o.get(v)).copy() for v in EV]
    st = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in ST]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    h = emu.cemu_create(C.c_int(nx), C.c_int(ny), bc, C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]), C.c_double(FLOORS["density_min"]), C.c_double(FLOORS["temp_min"]),
                        C.c_double(FLOORS["thermal_energy_min"]), C.c_double(0.2), vp(dx), vp(dy), (C.c_void_p * 8)(*[a.ctypes.data for a in ev]), (C.c_void_p * 5)(*[a.ctypes.data for a in st]))
    xl, xu = (0, nx - 1) if xb[0] == "periodic" else (2, nx - 3)
    yl, yu = (0, ny - 1) if yb[0] == "periodic" else (2, ny - 3)
    step = 0.2 * float(np.min(o.get("dt")[xl:xu + 1, yl:yu + 1]))
    return s, o, C.c_void_p(h), step, (xl, xu, yl, yu)


def compare(emu, h, o, nx, ny, what, bounds):
    for k, v in enumerate(EV):
        got = np.zeros((nx, ny))
        emu.cemu_get(h, C.c_int(k), vp(got))
        assert same_bits(got, o.get(v)), "%s: %s differs: %s" % (what, v, mismatch(got, o.get(v)))
    xl, xu, yl, yu = bounds
    assert float(emu.cemu_dtmin(h)).hex() == float(np.min(o.get("dt")[xl:xu + 1, yl:yu + 1])).hex(), "dt minimum after the module's propagate"


BOUNDS = [(("fixed", "open"), ("reflect", "open")), (("reflect", "reflect"), ("open", "fixed")), (("periodic", "periodic"), ("fixed", "open")), (("open", "open"), ("periodic", "periodic"))]


@pytest.mark.parametrize("xb,yb", BOUNDS)
def test_propagate_with_boundary_passes_equals_oracle_propagate(emu, xb, yb):
    """launch_propagate (floors, pointwise zeroing, the open / reflect passes in the reference's order, dt) on an oracle state against the oracle's propagateChanges.
    (Not idempotent by construction: an open pass re-reads first-interior momenta that a later reflect side has zeroed meanwhile -- in the reference as well.)"""
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    assert emu.cemu_propagate(h) == 0
    o.propagate()
    compare(emu, h, o, nx, ny, "propagate", b)
    o.close()


@pytest.mark.parametrize("xb,yb", BOUNDS)
def test_div_cleaning_through_dc_post(emu, xb, yb):
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    ts = 0.37 * step
    o.add_small_module("div_cleaning", epsilon=0.3, time_scale=ts)
    ns = emu.cemu_div_cleaning(h, C.c_double(0.3), C.c_double(ts), C.c_double(step))
    assert ns == int(step / (0.3 * ts)) + 1
    o.small_module_hooks(2, step)
    compare(emu, h, o, nx, ny, "div_cleaning", b)
    o.close()


@pytest.mark.parametrize("xb,yb", BOUNDS)
def test_field_heating_through_fh_pre_and_fh_iterate(emu, xb, yb):
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    kw = dict(coeff=1.0e-7, current_pow=0.5, b_pow=1.0, n_pow=0.2, roc_pow=0.3)
    o.add_small_module("field_heating", **kw)
    H = np.zeros((nx, ny))
    assert emu.cemu_field_heating(h, *[C.c_double(kw[k]) for k in ("coeff", "current_pow", "b_pow", "n_pow", "roc_pow")], C.c_int(0), C.c_double(step), vp(H)) == 0
    o.small_module_hooks(0, step)
    o.small_module_hooks(1, step)
    assert same_bits(H, o.small_module_plane(0, 0)), "heating plane (derived planes through k_mhd_derive): " + mismatch(H, o.small_module_plane(0, 0))
    compare(emu, h, o, nx, ny, "field_heating", b)
    o.close()


@pytest.mark.parametrize("name,kv,xb,yb,integrator", [c for c in AR_CASES if "frobenius" not in c[0] or True], ids=[c[0] for c in AR_CASES])
def test_anomalous_resistivity_through_ar_iterate(emu, name, kv, xb, yb, integrator):
    """ArExec (cells / reduce_min through the reduction slot), ar_setup_run and ar_iterate incl. the dt plane from derive_to, the write-back and the propagate"""
    nx, ny = 23, 23
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny, integrator=integrator)
    a = ar_kwargs(kv)
    o.set_anomalous_resistivity(**a)
    p = anomalous_params(**a)
    px = np.ascontiguousarray(s["planes"]["pos_x"], dtype=np.float64); py = np.ascontiguousarray(s["planes"]["pos_y"], dtype=np.float64)
    ij = (C.c_int * 2)(); nsub = C.c_int(); tmpl = np.zeros((nx, ny))
    iters = 2
    assert emu.cemu_anomalous_resistivity(h, vp(px), vp(py), vp(p), C.c_double(step), C.c_int(iters), ij, C.byref(nsub), vp(tmpl)) == 0
    for k in range(iters):
        if k == iters - 1:                                   # the last iteration's raw result, for the joule_heating plane (anomalousresistivity.cpp:167)
            e_before = o.get("thermal_energy")
            e_after = o.anomalous_core(step)[3]
            o.close()                                        # anomalous_core advanced the tracked null point: rebuild the oracle up to this iteration
            o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator=integrator, **FLOORS)
            o.run(2)
            o.set_anomalous_resistivity(**a)
            for _ in range(iters - 1):
                o.anomalous_iterate(step)
        o.anomalous_iterate(step)
    joule = np.zeros((nx, ny)); prod = np.zeros((nx, ny))
    assert emu.cemu_anomalous_outputs(h, vp(joule), vp(prod)) == 0
    assert same_bits(joule, (e_after - e_before) / step), "joule_heating plane: " + mismatch(joule, (e_after - e_before) / step)
    assert same_bits(prod, o.anomalous_state()[1] * o.anomalous_diffusivity()), "anomalous_diffusivity plane"
    (ri, rj), rt = o.anomalous_state()
    assert (ij[0], ij[1]) == (ri, rj) and nsub.value == o.anomalous_subcycles()
    assert same_bits(tmpl, rt), "template: " + mismatch(tmpl, rt)
    compare(emu, h, o, nx, ny, name, b)
    o.close()


@pytest.mark.parametrize("xb,yb", BOUNDS)
def test_source_terms_through_src_post(emu, xb, yb):
    """the four pointwise source terms through src_post itself: its time windows, the ramp of localized_heating, the oscillation factor of momentum_injection"""
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    for t in (3.0, 9.5, 30.0):                            # inside every window (rising ramp) / falling ramp, mass injection over / only momentum injection still active
        o.set_time(t)
        if t == 3.0:
            o.add_small_module("ambient_heating_sink", heating_rate=2.0e-4)
            o.add_small_module("localized_heating", start_time=1.0, duration=10.0, max_heating_rate=0.5, stddev_x=3.0, stddev_y=2.0, center_x=8.0, center_y=7.0, ramp_time=4.0)
            o.add_small_module("mass_injection", start_time=2.0, duration=5.0, max_injection_rate=1.0e6, stddev_x=2.0, stddev_y=2.5, center_x=9.0, center_y=6.0)
            o.add_small_module("momentum_injection", start_time=0.0, duration=50.0, max_accel=1.0e4, stddev_x=2.0, stddev_y=2.0, center_x=10.0, center_y=9.0, dir_x=0.6, dir_y=-0.8,
                               template_angle=20.0, oscillatory=1.0, oscillation_period=7.0)
            o.small_module_hooks(0, step)
            planes = [[o.small_module_plane(m, w) for w in (0, 1)] for m in range(4)]
        par = [(0, 0.0, 0.0, 0.0, 0.0, 0, 1.0), (1, 1.0, 10.0, 4.0, 0.0, 0, 1.0), (2, 2.0, 5.0, 0.0, 0.0, 0, 1.0), (3, 0.0, 50.0, 0.0, 1.0e4, 1, 7.0)]
        for m, (kind, start, dur, ramp, acc, osc, per) in enumerate(par):
            p0 = np.ascontiguousarray(planes[m][0]); p1 = np.ascontiguousarray(planes[m][1]) if planes[m][1] is not None else None
            rc = emu.cemu_source_term(h, C.c_int(kind), C.c_double(start), C.c_double(dur), C.c_double(ramp), C.c_double(acc), C.c_int(osc), C.c_double(per), vp(p0),
                                      vp(p1) if p1 is not None else None, C.c_double(t), C.c_double(step))
            assert rc == 0
        o.small_module_hooks(2, step)
        compare(emu, h, o, nx, ny, "source terms at t = %g" % t, b)
    o.close()


@pytest.mark.parametrize("boundary,shape,fa,dyn", [("y_bound_2", "exp", False, False), ("x_bound_1", "gaussian", True, True), ("y_bound_1", "flat", True, False), ("x_bound_2", "exp", False, True)])
@pytest.mark.parametrize("xb,yb", BOUNDS[:2])
def test_boundary_outflow_through_bo_post(emu, xb, yb, boundary, shape, fa, dyn):
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    codes = {"x_bound_1": 0, "x_bound_2": 1, "y_bound_1": 2, "y_bound_2": 3, "exp": 0, "gaussian": 1, "flat": 2}
    o.add_small_module("boundary_outflow", max_accel=3.0e4, falloff_length=6.0e8, boundary=float(codes[boundary]), falloff_shape=float(codes[shape]), feather_length=2.0e8,
                       field_aligned_mode=float(fa), dynamic_mode=float(dyn), dynamic_time=20.0, dynamic_target_speed=1.0e5)
    ref_mean = o.outflow_mean(0)
    px = np.ascontiguousarray(s["planes"]["pos_x"], dtype=np.float64); py = np.ascontiguousarray(s["planes"]["pos_y"], dtype=np.float64)
    mean, accel = C.c_double(), C.c_double()
    rc = emu.cemu_boundary_outflow(h, vp(px), vp(py), C.c_double(3.0e4), C.c_double(6.0e8), C.c_int(codes[boundary]), C.c_int(codes[shape]), C.c_double(2.0e8), C.c_int(int(fa)),
                                   C.c_int(int(dyn)), C.c_double(20.0), C.c_double(1.0e5), C.c_double(step), C.byref(mean), C.byref(accel))
    assert rc == 0
    assert mean.value.hex() == float(ref_mean).hex(), (mean.value, ref_mean)
    o.small_module_hooks(2, step)
    compare(emu, h, o, nx, ny, "boundary_outflow", b)
    o.close()


@pytest.mark.parametrize("limit", [False, True])
@pytest.mark.parametrize("xb,yb", [(("open_moc", "open_moc"), ("fixed", "open_moc")), (("periodic", "periodic"), ("open_moc", "open_moc")), (("open_moc", "reflect"), ("open_moc", "open"))])
def test_propagate_on_open_moc_sides_through_launch_moc_and_moc_limit(emu, xb, yb, limit):
    """launch_propagate with open_moc sides: the dt-only pass of launch_moc over the evolved ghost cells, and with moc_b_limiting / moc_mom_limiting the clamps of
    moc_limit followed by the rebuilt dt minimum -- against the oracle's propagateChanges on a state whose ghost-zone fields were pushed out of the limits"""
    nx, ny = 25, 22
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    floors = dict(density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", moc_limiting=dict(b_limiting=True, b_lower=0.5, b_upper=1.5, mom_limiting=True, mom_lower=0.5, mom_upper=1.5) if limit else None, **floors)
    o.run(2)
    rng = np.random.default_rng(5)
    for v in ("bi_x", "bi_y", "mom_x", "mom_y"):                       # push boundary values around so that the clamps act
        a = o.view(v)                                                  # the oracle's own plane, modified in place
        a[:2, :] *= rng.uniform(0.2, 3.0, size=a[:2, :].shape); a[-2:, :] *= rng.uniform(0.2, 3.0, size=a[-2:, :].shape)
        a[:, :2] *= rng.uniform(0.2, 3.0, size=a[:, :2].shape); a[:, -2:] *= rng.uniform(0.2, 3.0, size=a[:, -2:].shape)
    ev = [np.ascontiguousarray(o.get(v)).copy() for v in EV]
    st = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in ST]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    h = C.c_void_p(emu.cemu_create(C.c_int(nx), C.c_int(ny), bc, C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]), C.c_double(floors["density_min"]), C.c_double(floors["temp_min"]),
                                   C.c_double(floors["thermal_energy_min"]), C.c_double(0.2), vp(dx), vp(dy), (C.c_void_p * 8)(*[a.ctypes.data for a in ev]), (C.c_void_p * 5)(*[a.ctypes.data for a in st])))
    emu.cemu_set_moc(h, C.c_int(int(limit)), C.c_double(0.5), C.c_double(1.5), C.c_int(int(limit)), C.c_double(0.5), C.c_double(1.5))
    assert emu.cemu_propagate(h) == 0
    o.propagate()
    for k, v in enumerate(EV):
        got = np.zeros((nx, ny))
        emu.cemu_get(h, C.c_int(k), vp(got))
        assert same_bits(got, o.get(v)), "%s differs: %s" % (v, mismatch(got, o.get(v)))
    lo = lambda bnd: 0 if bnd in ("periodic", "open_moc") else 2
    hi = lambda bnd, n: n - 1 if bnd in ("periodic", "open_moc") else n - 3
    ref = float(np.min(o.get("dt")[lo(xb[0]):hi(xb[1], nx) + 1, lo(yb[0]):hi(yb[1], ny) + 1]))
    assert float(emu.cemu_dtmin(h)).hex() == ref.hex(), "dt minimum over the bounds widened by the open_moc ghost zones"
    o.close()


@pytest.mark.parametrize("which", ["propagate", "div_cleaning", "field_heating"])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("fixed", "open")), (("fixed", "open"), ("reflect", "open")), (("reflect", "reflect"), ("periodic", "periodic"))])
def test_module_hooks_on_slabs_equal_the_whole_domain(emu, xb, yb, world, which):
    """the slab form of the hooks: every rank is a host thread running capi.cu's own functions on its rows (+ halo rows); the peer transport is a staging copy between
    the threads.  exchange_plane / derive_to's halo rows / after_module_propagate / the dt all-gather must give the oracle's whole-domain result, bit for bit."""
    nx, ny = 26, 19
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
    o.run(2)
    ev = [np.ascontiguousarray(o.get(v)).copy() for v in EV]
    st = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in ST]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    cuts = [round(k * nx / world) for k in range(world + 1)]
    hs = []
    for r in range(world):
        hs.append(emu.cemu_create_slab(C.c_int(r), C.c_int(world), C.c_int(cuts[r]), C.c_int(cuts[r + 1] - cuts[r]), C.c_int(nx), C.c_int(ny), bc, C.c_double(s["ion_mass"]),
                                       C.c_double(s["adiabatic_index"]), C.c_double(FLOORS["density_min"]), C.c_double(FLOORS["temp_min"]), C.c_double(FLOORS["thermal_energy_min"]),
                                       C.c_double(0.2), vp(dx), vp(dy), (C.c_void_p * 8)(*[a.ctypes.data for a in ev]), (C.c_void_p * 5)(*[a.ctypes.data for a in st])))
    xl, xu = (0, nx - 1) if xb[0] == "periodic" else (2, nx - 3)
    yl, yu = (0, ny - 1) if yb[0] == "periodic" else (2, ny - 3)
    step = 0.2 * float(np.min(o.get("dt")[xl:xu + 1, yl:yu + 1]))
    if which == "propagate":
        p = np.zeros(1); code = 0
        o.propagate()
    elif which == "div_cleaning":
        ts = 0.37 * step
        p = np.array([0.3, ts, step]); code = 1
        o.add_small_module("div_cleaning", epsilon=0.3, time_scale=ts)
        o.small_module_hooks(2, step)
    else:
        kw = dict(coeff=1.0e-7, current_pow=0.5, b_pow=1.0, n_pow=0.2, roc_pow=0.3)
        p = np.array([kw["coeff"], kw["current_pow"], kw["b_pow"], kw["n_pow"], kw["roc_pow"], step]); code = 2
        o.add_small_module("field_heating", **kw)
        o.small_module_hooks(0, step); o.small_module_hooks(1, step)
    assert emu.cemu_run_slabs((C.c_void_p * world)(*hs), C.c_int(world), C.c_int(code), vp(p)) == 0
    for k, v in enumerate(EV):
        parts = []
        for r in range(world):
            a = np.zeros((cuts[r + 1] - cuts[r], ny))
            emu.cemu_get_slab(C.c_void_p(hs[r]), C.c_int(k), vp(a))
            parts.append(a)
        got = np.concatenate(parts, axis=0)
        assert same_bits(got, o.get(v)), "%s on %d slabs: %s differs: %s" % (which, world, v, mismatch(got, o.get(v)))
    ref = float(np.min(o.get("dt")[xl:xu + 1, yl:yu + 1])).hex()
    for r in range(world):
        assert float(emu.cemu_dtmin(C.c_void_p(hs[r]))).hex() == ref, "rank %d holds the global dt minimum" % r
    o.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("integ,sat", [("euler", True), ("rk4", False)])
@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("fixed", "open")), (("fixed", "open"), ("reflect", "open")), (("reflect", "reflect"), ("periodic", "periodic"))])
def test_thermal_conduction_on_slabs_equals_the_whole_domain_with_and_without_fast_instances(emu, xb, yb, integ, sat, world):
    """tc_iterate on 2 and 3 slabs (host threads, halo rows of the temperature plane exchanged after every sub-cycle stage): the thermal energy equals the single-rank
    run's bit for bit, with the FAST stencil instances on (on a slab every row of a periodic x axis qualifies: the halo rows hold the neighbours' cells) and off."""
    nx, ny = 27, 22
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
    o.run(2)
    xl, xu = (0, nx - 1) if xb[0] == "periodic" else (2, nx - 3)
    yl, yu = (0, ny - 1) if yb[0] == "periodic" else (2, ny - 3)
    step = 0.2 * float(np.min(o.get("dt")[xl:xu + 1, yl:yu + 1]))
    code = {"euler": 0, "rk2": 1, "rk4": 2}[integ]
    # single rank, general instances: the reference point (itself held to the oracle by the tests above)
    s1, o1, h1, _, _ = make_pair(emu, xb, yb, nx, ny)
    o1.close()
    emu.cemu_set_fast_interior(h1, C.c_int(0))
    dummy = np.zeros((nx, ny))
    assert emu.cemu_thermal_conduction(h1, C.c_int(int(sat)), C.c_double(1.0), C.c_double(1.0e-4), C.c_int(code), C.c_int(2), C.c_double(step), vp(dummy), vp(dummy.copy())) == 0
    whole = np.zeros((nx, ny))
    emu.cemu_get(h1, C.c_int(4), vp(whole))
    for fast in (1.0, 0.0):
        hs, cuts, keep = make_slabs(emu, s, o, xb, yb, nx, ny, world)
        p = np.array([float(sat), 1.0, 1.0e-4, float(code), 2.0, step, fast])
        assert emu.cemu_run_slabs((C.c_void_p * world)(*hs), C.c_int(world), C.c_int(3), vp(p)) == 0
        got = gather(emu, hs, cuts, ny, 4)
        assert same_bits(got, whole), "thermal energy on %d slabs (fast %d): %s" % (world, int(fast), mismatch(got, whole))
    # the plan-driven form (SPRUCE_DEVICE_SUBCYCLES): 5 sub-cycles enqueued on every slab, 2 planned -- the skipped stages still take part in the halo exchanges
    hs, cuts, keep = make_slabs(emu, s, o, xb, yb, nx, ny, world)
    p = np.array([float(sat), 1.0, 1.0e-4, float(code), 2.0, step, 1.0, 5.0])
    assert emu.cemu_run_slabs((C.c_void_p * world)(*hs), C.c_int(world), C.c_int(4), vp(p)) == 0
    got = gather(emu, hs, cuts, ny, 4)
    assert same_bits(got, whole), "thermal energy on %d slabs (device-resident plan): %s" % (world, mismatch(got, whole))
    assert not np.array_equal(whole, o.get("thermal_energy"))
    o.close()


def make_slabs(emu, s, o, xb, yb, nx, ny, world):
    ev = [np.ascontiguousarray(o.get(v)).copy() for v in EV]
    st = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in ST]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    cuts = [round(k * nx / world) for k in range(world + 1)]
    hs = [emu.cemu_create_slab(C.c_int(r), C.c_int(world), C.c_int(cuts[r]), C.c_int(cuts[r + 1] - cuts[r]), C.c_int(nx), C.c_int(ny), bc, C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]),
                               C.c_double(FLOORS["density_min"]), C.c_double(FLOORS["temp_min"]), C.c_double(FLOORS["thermal_energy_min"]), C.c_double(0.2), vp(dx), vp(dy),
                               (C.c_void_p * 8)(*[a.ctypes.data for a in ev]), (C.c_void_p * 5)(*[a.ctypes.data for a in st])) for r in range(world)]
    return hs, cuts, (ev, st, dx, dy)       # the arrays must outlive the calls


def gather(emu, hs, cuts, ny, k):
    parts = []
    for r, h in enumerate(hs):
        a = np.zeros((cuts[r + 1] - cuts[r], ny))
        emu.cemu_get_slab(C.c_void_p(h), C.c_int(k), vp(a))
        parts.append(a)
    return np.concatenate(parts, axis=0)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("boundary,shape,fa,dyn", [("y_bound_2", "exp", False, True), ("x_bound_1", "gaussian", True, True), ("x_bound_2", "flat", True, False)])
def test_boundary_outflow_on_slabs(emu, boundary, shape, fa, dyn, world):
    """the keyed maximum of k_bo_mean travels through the all-gathered module reductions: every rank must find the whole domain's mean outflow"""
    nx, ny = 26, 19
    xb, yb = ("fixed", "open"), ("reflect", "open")
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
    o.run(2)
    hs, cuts, keep = make_slabs(emu, s, o, xb, yb, nx, ny, world)
    step = 0.2 * float(np.min(o.get("dt")[2:-2, 2:-2]))
    codes = {"x_bound_1": 0, "x_bound_2": 1, "y_bound_1": 2, "y_bound_2": 3, "exp": 0, "gaussian": 1, "flat": 2}
    o.add_small_module("boundary_outflow", max_accel=3.0e4, falloff_length=6.0e8, boundary=float(codes[boundary]), falloff_shape=float(codes[shape]), feather_length=2.0e8,
                       field_aligned_mode=float(fa), dynamic_mode=float(dyn), dynamic_time=20.0, dynamic_target_speed=1.0e5)
    ref_mean = o.outflow_mean(0)
    px = np.ascontiguousarray(s["planes"]["pos_x"], dtype=np.float64); py = np.ascontiguousarray(s["planes"]["pos_y"], dtype=np.float64)
    means = np.zeros(world)
    rc = emu.cemu_run_slabs_outflow((C.c_void_p * world)(*hs), C.c_int(world), vp(px), vp(py), C.c_double(3.0e4), C.c_double(6.0e8), C.c_int(codes[boundary]), C.c_int(codes[shape]),
                                    C.c_double(2.0e8), C.c_int(int(fa)), C.c_int(int(dyn)), C.c_double(20.0), C.c_double(1.0e5), C.c_double(step), vp(means))
    assert rc == 0
    assert [float(m).hex() for m in means] == [float(ref_mean).hex()] * world
    o.small_module_hooks(2, step)
    for k, v in enumerate(EV):
        got = gather(emu, hs, cuts, ny, k)
        assert same_bits(got, o.get(v)), "%s differs: %s" % (v, mismatch(got, o.get(v)))
    o.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("limit", [False, True])
@pytest.mark.parametrize("xb,yb", [(("open_moc", "open_moc"), ("fixed", "open_moc")), (("periodic", "periodic"), ("open_moc", "open_moc"))])
def test_open_moc_propagate_on_slabs(emu, xb, yb, limit, world):
    """launch_moc's dt-only pass and moc_limit on slabs: strip cells addressed by global row, the x sides owned by the first / last slab"""
    nx, ny = 26, 19
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2",
               moc_limiting=dict(b_limiting=True, b_lower=0.5, b_upper=1.5, mom_limiting=True, mom_lower=0.5, mom_upper=1.5) if limit else None, **FLOORS)
    o.run(2)
    rng = np.random.default_rng(5)
    for v in ("bi_x", "bi_y", "mom_x", "mom_y"):
        a = o.view(v)
        a[:2, :] *= rng.uniform(0.2, 3.0, size=a[:2, :].shape); a[-2:, :] *= rng.uniform(0.2, 3.0, size=a[-2:, :].shape)
        a[:, :2] *= rng.uniform(0.2, 3.0, size=a[:, :2].shape); a[:, -2:] *= rng.uniform(0.2, 3.0, size=a[:, -2:].shape)
    hs, cuts, keep = make_slabs(emu, s, o, xb, yb, nx, ny, world)
    for h in hs:
        emu.cemu_set_moc(C.c_void_p(h), C.c_int(int(limit)), C.c_double(0.5), C.c_double(1.5), C.c_int(int(limit)), C.c_double(0.5), C.c_double(1.5))
    assert emu.cemu_run_slabs((C.c_void_p * world)(*hs), C.c_int(world), C.c_int(0), vp(np.zeros(1))) == 0
    o.propagate()
    for k, v in enumerate(EV):
        got = gather(emu, hs, cuts, ny, k)
        assert same_bits(got, o.get(v)), "%s differs: %s" % (v, mismatch(got, o.get(v)))
    lo = lambda bnd: 0 if bnd in ("periodic", "open_moc") else 2
    hi = lambda bnd, n: n - 1 if bnd in ("periodic", "open_moc") else n - 3
    ref = float(np.min(o.get("dt")[lo(xb[0]):hi(xb[1], nx) + 1, lo(yb[0]):hi(yb[1], ny) + 1])).hex()
    for h in hs:
        assert float(emu.cemu_dtmin(C.c_void_p(h))).hex() == ref
    o.close()


@pytest.mark.parametrize("integ,sat", [("euler", True), ("rk2", True), ("rk4", False), ("rk4", True)])
@pytest.mark.parametrize("xb,yb", BOUNDS[:2])
def test_thermal_conduction_through_tc_iterate_with_output_planes(emu, xb, yb, integ, sat):
    """tc_iterate itself, diagnostic hooks on; 1e-9: T^2.5 is (T*T)*sqrt(T) in the kernel and pow in the reference"""
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    o.set_thermal_conduction(flux_saturation=sat, integrator=integ, epsilon=0.1, dt_subcycle_min=1.0e-4, weakening_factor=1.0)
    o.step()
    ns = o.subcycles("thermal_conduction")
    avg = np.zeros((nx, ny)); satp = np.zeros((nx, ny))
    assert emu.cemu_thermal_conduction(h, C.c_int(int(sat)), C.c_double(1.0), C.c_double(1.0e-4), C.c_int({"euler": 0, "rk2": 1, "rk4": 2}[integ]), C.c_int(ns), C.c_double(step), vp(avg), vp(satp)) == 0
    ref_avg, ref_sat = o.module_output("thermal_conduction"), o.module_output("flux_saturation")
    assert ns >= 1 and np.count_nonzero(ref_avg) > 0
    assert np.max(np.abs(avg - ref_avg)) <= 1e-9 * np.max(np.abs(ref_avg))
    if sat:
        assert np.max(np.abs(satp - ref_sat)) <= 1e-9 * np.max(np.abs(ref_sat))
    o.close()


@pytest.mark.parametrize("integ,sat", [("euler", False), ("rk2", True), ("rk4", True)])
@pytest.mark.parametrize("xb,yb", BOUNDS)
def test_thermal_conduction_fast_interior_instance_equals_the_general_one(emu, xb, yb, integ, sat):
    """k_tc_stage serves deep-interior cells with the FAST instances of the stencil helpers (no index wrap / clamp, no range tests: module_kernels.cuh
    deep_interior): the thermal energy after tc_iterate and both diagnostic planes must equal the all-general run bit for bit, for every boundary set."""
    nx, ny = 27, 24
    res = []
    for fast in (1, 0):
        s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
        o.close()
        emu.cemu_set_fast_interior(h, C.c_int(fast))
        avg = np.zeros((nx, ny)); satp = np.zeros((nx, ny))
        assert emu.cemu_thermal_conduction(h, C.c_int(int(sat)), C.c_double(1.0), C.c_double(1.0e-4), C.c_int({"euler": 0, "rk2": 1, "rk4": 2}[integ]), C.c_int(3), C.c_double(step), vp(avg), vp(satp)) == 0
        e = np.zeros((nx, ny))
        emu.cemu_get(h, C.c_int(4), vp(e))
        res.append((e, avg, satp))
    for a, b2 in zip(res[0], res[1]):
        assert same_bits(a, b2)
    assert np.count_nonzero(res[0][1]) > 0


@pytest.mark.parametrize("integ", ["euler", "rk2", "rk4"])
@pytest.mark.parametrize("xb,yb", BOUNDS[:2])
def test_radiative_losses_through_rl_iterate_with_output_plane(emu, xb, yb, integ):
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    o.set_radiative_losses(integrator=integ, cutoff_ramp=1.0e3, cutoff_temp=3.0e4, epsilon=0.1, prevent_subcycling=False)
    o.step()
    ns = o.subcycles("radiative_losses")
    avg = np.zeros((nx, ny))
    assert emu.cemu_radiative_losses(h, C.c_int({"euler": 0, "rk2": 1, "rk4": 2}[integ]), C.c_double(1.0e3), C.c_double(3.0e4), C.c_double(0.1), C.c_int(0), C.c_int(ns), C.c_double(step), vp(avg)) == 0
    ref = o.module_output("rad")
    assert ns >= 1 and np.count_nonzero(ref) > 0
    assert np.max(np.abs(avg - ref)) <= 1e-9 * np.max(np.abs(ref))
    o.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("fixed", "open")), (("fixed", "open"), ("reflect", "open"))])
def test_source_terms_on_slabs(emu, xb, yb, world):
    nx, ny = 26, 19
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
    o.run(2)
    hs, cuts, keep = make_slabs(emu, s, o, xb, yb, nx, ny, world)
    xl, xu = (0, nx - 1) if xb[0] == "periodic" else (2, nx - 3)
    step = 0.2 * float(np.min(o.get("dt")[xl:xu + 1, 2:-2]))
    t = 3.0
    o.set_time(t)
    o.add_small_module("mass_injection", start_time=2.0, duration=5.0, max_injection_rate=1.0e6, stddev_x=2.0, stddev_y=2.5, center_x=9.0, center_y=6.0)
    o.add_small_module("momentum_injection", start_time=0.0, duration=50.0, max_accel=1.0e4, stddev_x=2.0, stddev_y=2.0, center_x=10.0, center_y=9.0, dir_x=0.6, dir_y=-0.8, template_angle=20.0,
                       oscillatory=1.0, oscillation_period=7.0)
    o.small_module_hooks(0, step)
    planes = [[o.small_module_plane(m, w) for w in (0, 1)] for m in range(2)]
    for m, (kind, start, dur, acc, osc, per) in enumerate([(2, 2.0, 5.0, 0.0, 0, 1.0), (3, 0.0, 50.0, 1.0e4, 1, 7.0)]):
        p0 = np.ascontiguousarray(planes[m][0]); p1 = np.ascontiguousarray(planes[m][1]) if planes[m][1] is not None else None
        rc = emu.cemu_run_slabs_source((C.c_void_p * world)(*hs), C.c_int(world), C.c_int(kind), C.c_double(start), C.c_double(dur), C.c_double(0.0), C.c_double(acc), C.c_int(osc), C.c_double(per),
                                       vp(p0), vp(p1) if p1 is not None else None, C.c_double(t), C.c_double(step))
        assert rc == 0
    o.small_module_hooks(2, step)
    for k, v in enumerate(EV):
        got = gather(emu, hs, cuts, ny, k)
        assert same_bits(got, o.get(v)), "%s differs: %s" % (v, mismatch(got, o.get(v)))
    o.close()


PV_CASES = [("euler_both", "euler", True, True, False, False), ("rk2_gc_both", "rk2", True, True, True, False), ("rk2_force_only", "rk2", False, True, False, False),
            ("euler_heat_only", "euler", True, False, False, False), ("euler_inactive", "euler", True, True, False, True), ("rk2_inactive", "rk2", True, True, True, True)]


@pytest.mark.parametrize("fast", [1, 0])
@pytest.mark.parametrize("name,integ,heating,force,gc,inactive", PV_CASES, ids=[c[0] for c in PV_CASES])
@pytest.mark.parametrize("xb,yb", BOUNDS[:3])
def test_physical_viscosity_through_pv_substeps_with_output_planes(emu, xb, yb, name, integ, heating, force, gc, inactive, fast):
    """pv_substeps itself (capi.cu: the sub-cycles of PhysicalViscosity::iterateModule, k_pv_stage from its own source), output_to_file planes on: the evolved planes,
    the dt minimum and viscous_heating / viscous_force_x/y/z -- the averages over the sub-cycles (physicalviscosity.cpp:151-170, 218-222), also in inactive_mode --
    within 1e-9 of the oracle, whose planes two reference fixtures pin bit for bit (loop_pv_diag_rk2, ot_pv_diag_inactive).  T^2.5 is (T*T)*sqrt(T) in the kernel."""
    from golden_util import physical_viscosity_coefficient
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    emu.cemu_set_fast_interior(h, C.c_int(fast))
    coeff = 1.0e-14
    cg = np.ascontiguousarray(physical_viscosity_coefficient(s["planes"], coeff, 6.0e8))
    o.set_physical_viscosity(cg, coeff=coeff, epsilon=0.1, heating_on=heating, force_on=force, gradient_correction=gc, integrator=integ, inactive_mode=inactive)
    before = o.get("thermal_energy").copy()
    o.step()
    ns = o.subcycles("physical_viscosity")
    assert ns >= 2
    out = np.zeros((4, nx, ny))
    assert emu.cemu_physical_viscosity(h, vp(cg), C.c_double(coeff), C.c_int(int(heating)), C.c_int(int(force)), C.c_int(int(gc)), C.c_int({"euler": 0, "rk2": 1}[integ]),
                                       C.c_int(int(inactive)), C.c_int(ns), C.c_double(step), vp(out)) == 0
    names = (["viscous_heating"] if heating else []) + (["viscous_force_x", "viscous_force_y", "viscous_force_z"] if force else [])
    for k, nm in enumerate(["viscous_heating", "viscous_force_x", "viscous_force_y", "viscous_force_z"]):
        ref = o.module_output(nm)
        if nm in names:
            assert np.count_nonzero(ref) > 0
            assert np.max(np.abs(out[k] - ref)) <= 1e-9 * np.max(np.abs(ref)), "%s %s: %.3e" % (name, nm, np.max(np.abs(out[k] - ref)) / np.max(np.abs(ref)))
        else:
            assert not out[k].any() and not ref.any()
    # the state after the hook's closing propagate: the oracle ran a whole step (hook, then the MHD stages), so compare with a second oracle that runs the hook alone
    o2 = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
    o2.run(2)
    o2.set_physical_viscosity(cg, coeff=coeff, epsilon=0.1, heating_on=heating, force_on=force, gradient_correction=gc, integrator=integ, inactive_mode=inactive)
    o2.physical_viscosity_iterate(step)
    for k, v in enumerate(EV):
        got = np.zeros((nx, ny))
        emu.cemu_get(h, C.c_int(k), vp(got))
        ref = o2.get(v)
        assert np.max(np.abs(got - ref)) <= 1e-9 * max(np.max(np.abs(ref)), 1e-300), "%s %s" % (name, v)
        if inactive:
            assert same_bits(got, ref)
    if not inactive and heating:
        assert not same_bits(o2.get("thermal_energy"), before)
    o.close(); o2.close()


AV_CASES = [
    ("rhs_and_hv_rk2_gc", [("boundary", 0.8, "v_x", "mom_x", 5.0e8), ("global", 3.0, "v_y", "mom_y", 0.0), ("boundary_global", 0.6, "mom_z", "mom_z", 8.0e8), ("local", 0.5, "temp", "thermal_energy", 0.0)], "rk2", True),
    ("hv_rk4", [("local", 2.5, "v_x", "mom_x", 0.0), ("global", 0.4, "temp", "thermal_energy", 0.0)], "rk4", False),
    ("hv_euler", [("global", 2.5, "v_y", "mom_y", 0.0), ("local", 0.4, "temp", "thermal_energy", 0.0)], "euler", False),
]


@pytest.mark.parametrize("name,terms,hv_integ,gc", AV_CASES, ids=[c[0] for c in AV_CASES])
@pytest.mark.parametrize("xb,yb", BOUNDS[:3])
def test_artificial_viscosity_output_planes_through_av_iterate_and_the_rhs_evaluation(emu, xb, yb, name, terms, hv_integ, gc):
    """The planes Viscosity::fileOutput appends per term (viscosity.cpp:351-376: dqdt, lap, str, dt -- each the leftover of the term's LAST evaluation), kept by
    k_visc_term / k_visc_apply when spruce_module_output_to_file("artificial_viscosity") is on.  One euler step of the reference is: the hyper-viscous sub-cycles
    (av_iterate), then ONE right-hand-side evaluation on the primary state in which every term is evaluated again (prepare_rhs_modules).  The product's own functions run
    that sequence on the host; every plane of every term, bit for bit against the oracle, whose planes three reference fixtures pin (*_visc_diag_*)."""
    from golden_util import boundary_viscosity_profile
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny, integrator="euler")
    full = []
    for opt, strength, vd, ve, length in terms:
        sg = np.ascontiguousarray(boundary_viscosity_profile(s["planes"]["pos_x"], s["planes"]["pos_y"], strength, length)) if opt.startswith("boundary") else None
        full.append(dict(opt=opt, strength=strength, var_diff=vd, var_evol=ve, species="i", strength_grid=sg))
    o.set_viscosity(full, hv_integrator=hv_integ, hv_epsilon=1.0, gradient_correction=gc)
    emu.cemu_viscosity_begin(h, C.c_int({"euler": 0, "rk2": 1, "rk4": 2}[hv_integ]), C.c_double(1.0), C.c_int(int(gc)))
    from oracle.oracle import VARS
    for t in full:
        sgp = vp(t["strength_grid"]) if t["strength_grid"] is not None else None
        assert emu.cemu_viscosity_term(h, C.c_int({"local": 0, "global": 1, "boundary": 2, "boundary_global": 3}[t["opt"]]), C.c_double(t["strength"]), C.c_int(VARS.index(t["var_diff"])),
                                       C.c_int(VARS.index(t["var_evol"])), sgp) == 0
    out = np.zeros((len(full), 4, nx, ny))
    before = np.zeros((len(full), 4, nx, ny))
    xl, xu, yl, yu = b
    dtmin_now = float(np.min(o.get("dt")[xl:xu + 1, yl:yu + 1]))
    assert emu.cemu_viscosity_run(h, C.c_double(step), C.c_double(dtmin_now), vp(before), vp(out)) == 0
    assert not before.any()                                       # zero planes before the first evaluation (:99-102)
    o.step()
    for i in range(len(full)):
        for w, which in enumerate(("dqdt", "lap", "str", "dt")):
            ref = o.viscosity_output(which, i)
            assert same_bits(out[i, w], ref), "%s term %d %s: %s" % (name, i, which, mismatch(out[i, w], ref))
        assert np.count_nonzero(out[i, 1]) > 0
    o.close()


def ms_planes(emu, h, nx, ny):
    out = np.zeros((3, nx, ny))
    emu.cemu_ms_planes(h, vp(out))
    return dict(zip(("cumulative_electron_heating", "cumulative_ion_heating", "cumulative_joule_heating"), out))


@pytest.mark.parametrize("xb,yb", BOUNDS[:3])
def test_multispecies_cumulative_planes_of_the_pointwise_modules_bit_for_bit(emu, xb, yb):
    """multispecies_mode (plasmadomain.hpp:134-135): ambient_heating_sink (default fraction 0.5, subtracting), localized_heating (0.2, inside its window, without the ramp
    factor as in the reference), ambient_heating (0.3) and anomalous_resistivity (joule plane) through src_post / ah_post / ar_iterate and k_ms_feed: the three cumulative
    planes bit-equal to the oracle's, which two reference fixtures pin (loop_ms_solar_rk2, ar_ms_joule_sources_rk2)"""
    nx, ny = 23, 23
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    o.set_multispecies(True, localized_heating=0.2, ambient_heating=0.3)
    assert emu.cemu_multispecies(h, C.c_double(1.0), C.c_double(1.0), C.c_double(0.3), C.c_double(0.0)) == 0
    emu.cemu_set_source_ms_fractions(C.c_double(0.5), C.c_double(0.2))
    o.set_time(3.0)
    o.add_small_module("ambient_heating_sink", heating_rate=2.0e-4)
    o.add_small_module("localized_heating", start_time=1.0, duration=10.0, max_heating_rate=0.5, stddev_x=3.0, stddev_y=2.0, center_x=8.0, center_y=7.0, ramp_time=4.0)
    o.small_module_hooks(0, step)
    planes = [o.small_module_plane(m, 0) for m in range(2)]
    for m, (kind, start, dur, ramp) in enumerate([(0, 0.0, 0.0, 0.0), (1, 1.0, 10.0, 4.0)]):
        p0 = np.ascontiguousarray(planes[m])
        assert emu.cemu_source_term(h, C.c_int(kind), C.c_double(start), C.c_double(dur), C.c_double(ramp), C.c_double(0.0), C.c_int(0), C.c_double(1.0), vp(p0), None, C.c_double(3.0), C.c_double(step)) == 0
    o.small_module_hooks(2, step)
    compare(emu, h, o, nx, ny, "sink + localized heating", b)
    a = ar_kwargs(AR_CASES[0][1])
    o.set_anomalous_resistivity(**a)
    px = np.ascontiguousarray(s["planes"]["pos_x"], dtype=np.float64); py = np.ascontiguousarray(s["planes"]["pos_y"], dtype=np.float64)
    ij = (C.c_int * 2)(); nsub = C.c_int(); tmpl = np.zeros((nx, ny))
    assert emu.cemu_anomalous_resistivity(h, vp(px), vp(py), vp(anomalous_params(**a)), C.c_double(step), C.c_int(1), ij, C.byref(nsub), vp(tmpl)) == 0
    o.anomalous_iterate(step)
    got = ms_planes(emu, h, nx, ny)
    for name in got:
        ref = o.ms_plane(name)
        assert np.count_nonzero(ref) > 0 and same_bits(got[name], ref), "%s: %s" % (name, mismatch(got[name], ref))
    # ambient_heating is a post-iterate hook of a whole step in the oracle; its feed ((mask fr) heating) dt does not depend on the state: compare the increment
    heating = np.ascontiguousarray(1.0e-4 * (1.0 + 0.1 * np.cos(np.arange(nx * ny).reshape(nx, ny))))
    before = ms_planes(emu, h, nx, ny)
    assert emu.cemu_ambient_heating(h, vp(heating), C.c_double(step)) == 0
    after = ms_planes(emu, h, nx, ny)
    xl, xu, yl, yu = b
    mask = np.zeros((nx, ny)); mask[xl:xu + 1, yl:yu + 1] = 1.0
    assert same_bits(after["cumulative_ion_heating"], before["cumulative_ion_heating"] + ((mask * (1.0 - 0.3)) * heating) * step)
    assert same_bits(after["cumulative_electron_heating"], before["cumulative_electron_heating"] + ((mask * 0.3) * heating) * step)
    assert same_bits(after["cumulative_joule_heating"], before["cumulative_joule_heating"])
    o.close()


@pytest.mark.parametrize("xb,yb", BOUNDS[:2])
def test_multispecies_cumulative_planes_of_the_subcycled_modules(emu, xb, yb):
    """thermal_conduction (fraction 0.7), radiative_losses (0.4) and physical_viscosity (0.6, rk2 sub-cycles) with multispecies_mode on: the electron / ion planes after
    tc_iterate + rl_iterate, then after pv_substeps, within the modules' 1e-9 of the oracle's (libm powers); the joule plane stays zero"""
    from golden_util import physical_viscosity_coefficient
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    o.set_multispecies(True, thermal_conduction=0.7, radiative_losses=0.4, physical_viscosity=0.6)
    assert emu.cemu_multispecies(h, C.c_double(0.7), C.c_double(0.4), C.c_double(0.5), C.c_double(0.6)) == 0
    o.set_thermal_conduction(flux_saturation=True, integrator="rk2", epsilon=0.1, dt_subcycle_min=1.0e-4, weakening_factor=1.0)
    o.set_radiative_losses(integrator="rk2", cutoff_ramp=1.0e3, cutoff_temp=3.0e4, epsilon=0.1, prevent_subcycling=False)
    o.step()
    avg = np.zeros((nx, ny)); satp = np.zeros((nx, ny))
    assert emu.cemu_thermal_conduction(h, C.c_int(1), C.c_double(1.0), C.c_double(1.0e-4), C.c_int(1), C.c_int(o.subcycles("thermal_conduction")), C.c_double(step), vp(avg), vp(satp)) == 0
    assert emu.cemu_radiative_losses(h, C.c_int(1), C.c_double(1.0e3), C.c_double(3.0e4), C.c_double(0.1), C.c_int(0), C.c_int(o.subcycles("radiative_losses")), C.c_double(step), vp(avg)) == 0
    got = ms_planes(emu, h, nx, ny)
    for name in ("cumulative_electron_heating", "cumulative_ion_heating"):
        ref = o.ms_plane(name)
        assert np.count_nonzero(ref) > 0 and np.max(np.abs(got[name] - ref)) <= 1e-9 * np.max(np.abs(ref)), name
    assert not got["cumulative_joule_heating"].any()
    o.close()
    # physical viscosity on a fresh pair (the oracle above has gone through a whole step)
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    o.set_multispecies(True, physical_viscosity=0.6)
    assert emu.cemu_multispecies(h, C.c_double(1.0), C.c_double(1.0), C.c_double(0.5), C.c_double(0.6)) == 0
    coeff = 1.0e-14
    cg = np.ascontiguousarray(physical_viscosity_coefficient(s["planes"], coeff, 6.0e8))
    o.set_physical_viscosity(cg, coeff=coeff, epsilon=0.1, heating_on=True, force_on=True, gradient_correction=False, integrator="rk2", inactive_mode=False)
    o.physical_viscosity_iterate(step)
    out = np.zeros((4, nx, ny))
    assert emu.cemu_physical_viscosity(h, vp(cg), C.c_double(coeff), C.c_int(1), C.c_int(1), C.c_int(0), C.c_int(1), C.c_int(0), C.c_int(o.subcycles("physical_viscosity")), C.c_double(step), vp(out)) == 0
    got = ms_planes(emu, h, nx, ny)
    for name in ("cumulative_electron_heating", "cumulative_ion_heating"):
        ref = o.ms_plane(name)
        assert np.count_nonzero(ref) > 0 and np.max(np.abs(got[name] - ref)) <= 1e-9 * np.max(np.abs(ref)), name
    o.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("integ,gc", [("euler", False), ("rk2", True)])
@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("fixed", "open")), (("fixed", "open"), ("reflect", "open"))])
def test_physical_viscosity_on_slabs_equals_the_whole_domain_with_and_without_fast_instances(emu, xb, yb, integ, gc, world):
    """pv_substeps on 2 and 3 slabs (host threads; velocity and temperature halo rows exchanged after every sub-cycle stage, the coefficient plane's once): momenta, thermal
    energy and the four output planes equal the single-rank run's bit for bit, FAST stencil instances on and off (the combination no device test has seen yet)"""
    from golden_util import physical_viscosity_coefficient
    nx, ny = 27, 22
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
    o.run(2)
    xl, xu = (0, nx - 1) if xb[0] == "periodic" else (2, nx - 3)
    yl, yu = (0, ny - 1) if yb[0] == "periodic" else (2, ny - 3)
    step = 0.2 * float(np.min(o.get("dt")[xl:xu + 1, yl:yu + 1]))
    coeff, ns, code = 1.0e-14, 3, {"euler": 0, "rk2": 1}[integ]
    cg = np.ascontiguousarray(physical_viscosity_coefficient(s["planes"], coeff, 6.0e8))
    s1, o1, h1, _, _ = make_pair(emu, xb, yb, nx, ny)
    o1.close()
    emu.cemu_set_fast_interior(h1, C.c_int(0))
    whole_avg = np.zeros((4, nx, ny))
    assert emu.cemu_physical_viscosity(h1, vp(cg), C.c_double(coeff), C.c_int(1), C.c_int(1), C.c_int(int(gc)), C.c_int(code), C.c_int(0), C.c_int(ns), C.c_double(step), vp(whole_avg)) == 0
    whole = [np.zeros((nx, ny)) for _ in EV]
    for k in range(len(EV)):
        emu.cemu_get(h1, C.c_int(k), vp(whole[k]))
    for fast in (1.0, 0.0):
        hs, cuts, keep = make_slabs(emu, s, o, xb, yb, nx, ny, world)
        p = np.array([coeff, 1.0, 1.0, float(gc), float(code), 0.0, float(ns), step, fast])
        avg = np.zeros(4 * nx * ny)
        assert emu.cemu_run_slabs_pv((C.c_void_p * world)(*hs), C.c_int(world), vp(cg), vp(p), vp(avg)) == 0
        for k, v in enumerate(EV):
            got = gather(emu, hs, cuts, ny, k)
            assert same_bits(got, whole[k]), "%s on %d slabs (fast %d): %s" % (v, world, int(fast), mismatch(got, whole[k]))
        off = 0
        for r in range(world):
            rows = cuts[r + 1] - cuts[r]
            blk = avg[off:off + 4 * rows * ny].reshape(4, rows, ny); off += 4 * rows * ny
            for q in range(4):
                assert same_bits(blk[q], whole_avg[q][cuts[r]:cuts[r + 1]]), "output plane %d of rank %d (fast %d)" % (q, r, int(fast))
    assert np.count_nonzero(whole_avg[0]) > 0 and not np.array_equal(whole[4], o.get("thermal_energy"))
    o.close()


@pytest.mark.parametrize("xb,yb", BOUNDS[:2])
def test_inactive_mode_of_conduction_and_losses_forms_the_planes_and_leaves_the_state(emu, xb, yb):
    """inactive_mode = true (thermalconduction.cpp:109, radiativelosses.cpp:98) through tc_iterate / rl_iterate: the sub-cycles run on a copy, the output_to_file and cumulative
    planes equal the active run's within 1e-9 of the oracle's (pinned by the fixture loop_inactive_tc_rl_rk2), every evolved plane keeps its bits"""
    nx, ny = 25, 22
    s, o, h, step, b = make_pair(emu, xb, yb, nx, ny)
    o.set_multispecies(True, thermal_conduction=0.7)
    o.set_thermal_conduction(flux_saturation=True, integrator="rk2", epsilon=0.1, dt_subcycle_min=1.0e-4, weakening_factor=1.0)
    o.set_radiative_losses(integrator="euler", cutoff_ramp=1.0e3, cutoff_temp=3.0e4, epsilon=0.1, prevent_subcycling=False)
    o.set_module_inactive("thermal_conduction"); o.set_module_inactive("radiative_losses")
    before = [np.zeros((nx, ny)) for _ in EV]
    for k in range(len(EV)):
        emu.cemu_get(h, C.c_int(k), vp(before[k]))
    assert emu.cemu_multispecies(h, C.c_double(0.7), C.c_double(1.0), C.c_double(0.5), C.c_double(0.0)) == 0
    emu.cemu_set_inactive(h, C.c_int(1), C.c_int(1))
    o.step()
    avg = np.zeros((nx, ny)); satp = np.zeros((nx, ny)); rad = np.zeros((nx, ny))
    assert emu.cemu_thermal_conduction(h, C.c_int(1), C.c_double(1.0), C.c_double(1.0e-4), C.c_int(1), C.c_int(o.subcycles("thermal_conduction")), C.c_double(step), vp(avg), vp(satp)) == 0
    assert emu.cemu_radiative_losses(h, C.c_int(0), C.c_double(1.0e3), C.c_double(3.0e4), C.c_double(0.1), C.c_int(0), C.c_int(o.subcycles("radiative_losses")), C.c_double(step), vp(rad)) == 0
    for k in range(len(EV)):
        now = np.zeros((nx, ny))
        emu.cemu_get(h, C.c_int(k), vp(now))
        assert same_bits(now, before[k]), EV[k]
    close = lambda a, r: np.count_nonzero(r) > 0 and np.max(np.abs(a - r)) <= 1e-9 * np.max(np.abs(r))
    assert close(avg, o.module_output("thermal_conduction")) and close(satp, o.module_output("flux_saturation")) and close(rad, o.module_output("rad"))
    got = ms_planes(emu, h, nx, ny)
    for name in ("cumulative_electron_heating", "cumulative_ion_heating"):
        assert close(got[name], o.ms_plane(name)), name
    o.close()


# ---- the device-resident sub-cycle plan (SPRUCE_DEVICE_SUBCYCLES=1): k_sub_plan + the plan-driven tc_iterate / rl_iterate against the host-driven form of the same step
def _bits(x):
    return int(np.array([x], dtype=np.float64).view(np.uint64)[0])


def _run_planned(emu, xb, yb, tc, rl, red, step, budget=None, stopped=0, nx=22, ny=19):
    """(host-driven run, plan-driven run) of one module step on twin domains: evolved planes, output planes, counts"""
    out = []
    for planned in (False, True):
        s, o, h, _, bounds = make_pair(emu, xb, yb, nx, ny)
        n = nx * ny
        avg, sat, rad = np.zeros(n), np.zeros(n), np.zeros(n)
        redv = (C.c_ulonglong * 4)(*red)
        common = [C.c_int(int(tc is not None)), C.c_int(int(bool(tc and tc["sat"]))), C.c_double(1.0e-4), C.c_int(tc["integ"] if tc else 0), C.c_double(tc["eps"] if tc else 0.1),
                  C.c_int(int(rl is not None)), C.c_int(rl["integ"] if rl else 0), C.c_double(rl["eps"] if rl else 0.1), redv, C.c_double(step)]
        if planned:
            info = (C.c_int * 7)()
            assert emu.cemu_planned_modules(h, *common, C.c_int(budget), C.c_int(stopped), vp(avg), vp(sat), vp(rad), info) == 0
        else:
            info = (C.c_int * 2)()
            assert emu.cemu_host_modules(h, *common, vp(avg), vp(sat), vp(rad), info) == 0
        planes = []
        for k in range(len(EV)):
            got = np.zeros((nx, ny))
            emu.cemu_get(h, C.c_int(k), vp(got))
            planes.append(got)
        out.append(dict(planes=planes, avg=avg, sat=sat, rad=rad, info=list(info), before=[np.ascontiguousarray(o.get(v)).copy() for v in EV], dtmin=emu.cemu_dtmin(h)))
    return out


@pytest.mark.parametrize("integ", [0, 1, 2])
@pytest.mark.parametrize("sat", [False, True])
@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("periodic", "periodic")), (("periodic", "periodic"), ("fixed", "open")), (("reflect", "open"), ("fixed", "fixed"))])
def test_planned_subcycles_equal_the_host_driven_step(emu, xb, yb, integ, sat):
    """thermal_conduction + radiative_losses: counts planned by k_sub_plan equal the host's numberSubcycles; with more sub-cycles enqueued than planned, the evolved planes,
    the dt minimum and the three output planes equal the host-driven step bit for bit"""
    _, _, _, step, _ = make_pair(emu, xb, yb, 22, 19)
    # reduction words that make thermal_conduction plan 3 and radiative_losses 2 sub-cycles (epsilon = 0.1); word 1 / 3: non-zero maxima
    red = [_bits(step / 2.5 / 0.1), _bits(1.0), _bits(step / 1.5 / 0.1), _bits(1.0)]
    host, dev = _run_planned(emu, xb, yb, dict(sat=sat, integ=integ, eps=0.1), dict(integ=integ, eps=0.1), red, step, budget=7)
    assert host["info"] == [3, 2]
    assert dev["info"] == [3, 2, 0, 3, 3, 2, 0]
    for k, v in enumerate(EV):
        assert same_bits(dev["planes"][k], host["planes"][k]), "%s: %s" % (v, mismatch(dev["planes"][k], host["planes"][k]))
    assert not same_bits(dev["planes"][EV.index("thermal_energy")], dev["before"][EV.index("thermal_energy")])
    assert dev["dtmin"] == host["dtmin"]
    for name in ("avg", "sat", "rad"):
        assert same_bits(dev[name], host[name]), name


def test_planned_subcycles_zero_counts_and_single_modules(emu):
    """saturated conduction with a vanishing field-aligned gradient plans no sub-cycle (thermalconduction.cpp:143), a vanishing loss maximum none for radiative_losses
    (radiativelosses.cpp:162); each module alone"""
    xb, yb = ("periodic", "periodic"), ("fixed", "open")
    _, _, _, step, _ = make_pair(emu, xb, yb, 22, 19)
    red = [_bits(step / 2.5 / 0.1), 0, _bits(step / 1.5 / 0.1), 0]
    host, dev = _run_planned(emu, xb, yb, dict(sat=True, integ=1, eps=0.1), dict(integ=1, eps=0.1), red, step, budget=4)
    assert host["info"] == [0, 0] and dev["info"][:2] == [0, 0] and dev["info"][4:] == [0, 0, 0]
    for k in range(len(EV)):
        assert same_bits(dev["planes"][k], host["planes"][k])
    red = [_bits(step / 1.5 / 0.1), _bits(1.0), _bits(step / 3.5 / 0.1), _bits(1.0)]
    for tc, rl, want in ((dict(sat=False, integ=0, eps=0.1), None, [2, 0]), (None, dict(integ=2, eps=0.1), [0, 4])):
        host, dev = _run_planned(emu, xb, yb, tc, rl, red, step, budget=4)
        assert host["info"] == want and dev["info"][:2] == want
        for k in range(len(EV)):
            assert same_bits(dev["planes"][k], host["planes"][k])
        assert same_bits(dev["avg"], host["avg"]) and same_bits(dev["rad"], host["rad"])


@pytest.mark.parametrize("stopped", [0, 1])
def test_planned_subcycles_stop_before_anything_changes(emu, stopped):
    """a count above the enqueued budget stops the run (StepCtl::done = 3) and reports what it needs; a run that had already stopped (max_time) stays stopped: in both cases
    no evolved plane, no output plane and not the dt minimum may change"""
    xb, yb = ("periodic", "periodic"), ("fixed", "open")
    _, _, _, step, _ = make_pair(emu, xb, yb, 22, 19)
    red = [_bits(step / 4.5 / 0.1), _bits(1.0), _bits(step / 1.5 / 0.1), _bits(1.0)]          # 5 conduction sub-cycles wanted
    host, dev = _run_planned(emu, xb, yb, dict(sat=True, integ=1, eps=0.1), dict(integ=0, eps=0.1), red, step, budget=4, stopped=stopped)
    assert host["info"] == [5, 2]
    if stopped:
        assert dev["info"] == [0, 0, 0, 0, -1, -1, 1]
    else:
        assert dev["info"] == [0, 0, 5, 5, -1, -1, 3]
    for k, v in enumerate(EV):
        assert same_bits(dev["planes"][k], dev["before"][k]), v
    assert not dev["avg"].any() and not dev["sat"].any() and not dev["rad"].any()


def test_fused_conduction_planes_equal_the_three_derive_passes(emu):
    """k_tc_derive (temp, b_hat_x, b_hat_y in one pass) against three derive_to passes, bit for bit, ghost cells included"""
    for xb, yb in ((("periodic", "periodic"), ("fixed", "open")), (("reflect", "open"), ("fixed", "fixed"))):
        _, _, h, _, _ = make_pair(emu, xb, yb, 22, 19)
        assert emu.cemu_tc_derive_mismatches(h) == 0


@pytest.mark.parametrize("integ", ["euler", "rk2", "rk4"])
@pytest.mark.parametrize("fast", [1, 0])
@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("periodic", "periodic")), (("periodic", "periodic"), ("fixed", "open")), (("reflect", "open"), ("fixed", "fixed")),
                                   (("fixed", "open"), ("reflect", "open"))])
def test_two_pass_saturated_conduction_equals_the_five_point_form(emu, xb, yb, integ, fast):
    """k_tc_coef + the plane-differentiating saturated branch (default) against the form that evaluates the coefficient at five points per cell (SPRUCE_TC_TWO_PASS=0): thermal
    energy and both output planes after tc_iterate bit for bit, FAST instances on and off"""
    res = []
    for two_pass in (1, 0):
        s, o, h, step, _ = make_pair(emu, xb, yb, 22, 19)
        emu.cemu_set_fast_interior(h, C.c_int(fast))
        emu.cemu_set_two_pass(h, C.c_int(two_pass))
        avg, satp = np.zeros(22 * 19), np.zeros(22 * 19)
        assert emu.cemu_thermal_conduction(h, C.c_int(1), C.c_double(1.0), C.c_double(1.0e-4), C.c_int({"euler": 0, "rk2": 1, "rk4": 2}[integ]), C.c_int(3), C.c_double(step), vp(avg), vp(satp)) == 0
        e = np.zeros((22, 19))
        emu.cemu_get(h, C.c_int(4), vp(e))
        res.append((e, avg, satp, np.ascontiguousarray(o.get("thermal_energy")).copy()))
        o.close()
    for a, b, name in zip(res[0][:3], res[1][:3], ("thermal energy", "thermal_conduction plane", "flux_saturation plane")):
        assert same_bits(a, b), "%s: %s" % (name, mismatch(a.reshape(22, 19), b.reshape(22, 19)))
    assert not same_bits(res[0][0], res[0][3])
