"""The late solar-module kernels (k_source_term, k_dc_update / k_plane_sum with the k_operator passes, k_fh_compute / k_fh_apply, k_bo_mean / k_bo_apply) and
the pointwise propagate kernel, EXECUTED ON THE HOST: their source text is cut from spruce_b200/csrc/*.cuh between fixed markers, compiled with g++ over
stand-ins for the CUDA keywords, and "launched" by looping over blocks and threads in the order of capi.cu's src_post / dc_post / fh_pre + fh_iterate /
bo_post.  Results must equal the CPU restatement's module hooks (pinned to the reference) bit for bit -- planes, sub-cycle count, mean outflow.
What this cannot see: launch configuration and memory-space mistakes; those are the GPU tests' part (tests/test_zz_gpu_unvalidated.py)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import mismatch, same_bits
from oracle.oracle import Oracle
from spruce_b200 import synthetic

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "spruce_b200" / "csrc"
BUILD = ROOT / "tests" / "hostcheck" / "_build"
LIB = BUILD / "libkernel_emu.so"
BC = {"periodic": 0, "open": 1, "fixed": 2, "reflect": 3, "open_moc": 4}

PRELUDE = r'''
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __noinline__
using std::max; using std::min;
static inline long long __double_as_longlong(double x) { long long b; std::memcpy(&b, &x, 8); return b; }
static inline double __longlong_as_double(long long b) { double x; std::memcpy(&x, &b, 8); return x; }
static inline int __double2hiint(double x) { return (int)(__double_as_longlong(x) >> 32); }
static inline int __double2loint(double x) { return (int)(__double_as_longlong(x) & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) { return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo)); }
struct dim3e { unsigned x, y, z; };
static thread_local dim3e threadIdx, blockIdx, blockDim, gridDim;      // thread_local: the slab tests run one host thread per rank
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v < o) *p = v; return o; }
// the warp reductions of the sub-cycle count kernels are compiled, never run here
static inline unsigned long long __shfl_xor_sync(unsigned, unsigned long long v, int) { return v; }
#define SPRUCE_EXACT_MATH_HOST_CHECK 1
#include "cell_math.cuh"
#include "solar_templates.hpp"
'''
BLOCK_MIN = r'''
// sequential stand-in for the block minimum: same NaN / sign handling and encoding as block_min_impl
static inline void block_min_to_global(double v, unsigned long long *target)
{
    unsigned long long b = (v == v) ? (unsigned long long)__double_as_longlong(v) : 0x7FF0000000000000ULL;
    if (v < 0.0) b = 0ULL;
    if (b < *target) *target = b;
}
'''


def cut(text, start, end, include_end=False):
    i = text.index(start)
    j = text.index(end, i + len(start))
    if include_end:
        j = text.index("\n", j) + 1
    return text[i:j]


def assemble():
    mk = (CSRC / "mhd_kernels.cuh").read_text()
    mo = (CSRC / "module_kernels.cuh").read_text()
    ca = (CSRC / "capi.cu").read_text()
    ms = (CSRC / "moc_stage.cuh").read_text()
    parts = [PRELUDE, "#include <vector>\nnamespace spruce {\n",
             cut(mk, "constexpr int HALO", "enum { KM_NONE", include_end=True),
             cut(mk, "__device__ __forceinline__ FaceGeom load_face_geom", "// is global row g / column j inside"),
             BLOCK_MIN,
             cut(mk, "// is global row g / column j inside", "// block-wide NaN-ignoring minimum"),
             cut(mk, "struct PropArgs {", "// Ghost cells of the non-periodic sides"), "\n",
             cut(mk, "struct StepCtl {", "// begin: step = epsilon"),
             cut(mo, "constexpr double kKappa0", "// Artificial viscosity (source/modules/viscosity.cpp"), "\n",
             cut(mo, "struct OpArgs", "}  // namespace spruce"),
             "}  // namespace spruce\n#include \"moc_kernels.cuh\"\nnamespace spruce {\n",
             cut(ms, "struct MocArgs {", "}  // namespace spruce"),
             cut(ca, "struct HostAxis {", "struct TwoFluid;"),
             cut(ca, "void build_axis(", "int upload_tables("),
             "}  // namespace spruce\n",
             (ROOT / "tests" / "hostcheck" / "kernel_emu_main.inc").read_text()]
    return "".join(parts)


@pytest.fixture(scope="module")
def emu():
    BUILD.mkdir(exist_ok=True)
    src = BUILD / "kernel_emu.cpp"
    text = assemble()
    if not LIB.exists() or not src.exists() or src.read_text() != text:
        LIB.unlink(missing_ok=True)                # a failed compile must not leave the previous library behind
        src.write_text(text)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(CSRC), "-o", str(LIB), str(src)], check=True)
    L = C.CDLL(str(LIB))
    L.emu_create.restype = C.c_void_p
    L.emu_div_cleaning.restype = C.c_int
    return L


EV = ["n", "mom_x", "mom_y", "mom_z", "thermal_energy", "bi_x", "bi_y", "bi_z"]
ST = ["be_x", "be_y", "be_z", "grav_x", "grav_y"]
FLOORS = dict(density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def make_pair(emu, xb, yb, nx=24, ny=21, warm=2):
    """an oracle a few steps into a run, and the emulator loaded with its planes"""
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
    o.run(warm)
    ev = [np.ascontiguousarray(o.get(v)).copy() for v in EV]
    st = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in ST]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    h = emu.emu_create(C.c_int(nx), C.c_int(ny), bc, C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]), C.c_double(FLOORS["density_min"]), C.c_double(FLOORS["temp_min"]),
                       C.c_double(FLOORS["thermal_energy_min"]), C.c_double(0.2), vp(dx), vp(dy), (C.c_void_p * 8)(*[a.ctypes.data for a in ev]), (C.c_void_p * 5)(*[a.ctypes.data for a in st]))
    step = 0.2 * float(np.min(o.get("dt")[2:-2, 2:-2]))
    return s, o, C.c_void_p(h), step


def compare(emu, h, o, nx, ny, what):
    for k, v in enumerate(EV):
        got = np.zeros((nx, ny))
        emu.emu_get(h, C.c_int(k), vp(got))
        assert same_bits(got, o.get(v)), "%s: %s differs: %s" % (what, v, mismatch(got, o.get(v)))


BOUNDS = [(("fixed", "fixed"), ("fixed", "fixed")), (("periodic", "periodic"), ("fixed", "fixed")), (("fixed", "fixed"), ("periodic", "periodic"))]


@pytest.mark.parametrize("xb,yb", BOUNDS)
def test_source_term_kernel_equals_oracle_hooks(emu, xb, yb):
    nx, ny = 24, 21
    s, o, h, step = make_pair(emu, xb, yb, nx, ny)
    t = 3.0
    o.set_time(t)
    o.add_small_module("ambient_heating_sink", heating_rate=2.0e-4)
    o.add_small_module("localized_heating", start_time=1.0, duration=10.0, max_heating_rate=0.5, stddev_x=3.0, stddev_y=2.0, center_x=8.0, center_y=7.0, ramp_time=4.0)
    o.add_small_module("mass_injection", start_time=2.0, duration=5.0, max_injection_rate=1.0e6, stddev_x=2.0, stddev_y=2.5, center_x=9.0, center_y=6.0)
    o.add_small_module("momentum_injection", start_time=0.0, duration=50.0, max_accel=1.0e4, stddev_x=2.0, stddev_y=2.0, center_x=10.0, center_y=9.0, dir_x=0.6, dir_y=-0.8, template_angle=20.0,
                       oscillatory=1.0, oscillation_period=7.0)
    o.small_module_hooks(0, step)                                      # preIterate builds the time-windowed templates
    planes = [[o.small_module_plane(m, w) for w in (0, 1)] for m in range(4)]
    # the scalars src_post forms on the host (capi.cu), from the step's start time
    ramp = (t - 1.0) / 4.0                                             # localizedheating.cpp:53-59: t - start = 2 < ramp_time and <= duration/2
    osc = np.sin(2.0 * np.pi * (t - 0.0) / 7.0)
    fs = [step, step * ramp, step, step * (osc * 1.0e4)]
    for m, (kind, f) in enumerate(zip((0, 1, 2, 3), fs)):
        p0 = np.ascontiguousarray(planes[m][0]); p1 = np.ascontiguousarray(planes[m][1]) if planes[m][1] is not None else p0
        emu.emu_source_term(h, C.c_int(kind), C.c_double(f), vp(p0), vp(p1))
    o.small_module_hooks(2, step)
    compare(emu, h, o, nx, ny, "source terms")
    emu.emu_destroy(h); o.close()


@pytest.mark.parametrize("xb,yb", BOUNDS)
def test_div_cleaning_launch_sequence_equals_oracle(emu, xb, yb):
    nx, ny = 24, 21
    s, o, h, step = make_pair(emu, xb, yb, nx, ny)
    ts = 0.37 * step
    o.add_small_module("div_cleaning", epsilon=0.3, time_scale=ts)
    ns = emu.emu_div_cleaning(h, C.c_double(step), C.c_double(0.3), C.c_double(ts))
    assert ns == int(step / (0.3 * ts)) + 1 and ns >= 3
    before = o.get("bi_x").copy()
    o.small_module_hooks(2, step)
    assert not same_bits(before, o.get("bi_x")), "the case cleans nothing"
    compare(emu, h, o, nx, ny, "div_cleaning")
    emu.emu_destroy(h); o.close()


@pytest.mark.parametrize("xb,yb", BOUNDS)
def test_field_heating_kernels_equal_oracle(emu, xb, yb):
    nx, ny = 24, 21
    s, o, h, step = make_pair(emu, xb, yb, nx, ny)
    kw = dict(coeff=1.0e-7, current_pow=0.5, b_pow=1.0, n_pow=0.2, roc_pow=0.3)
    o.add_small_module("field_heating", **kw)
    d = [np.ascontiguousarray(o.get(v)).copy() for v in ("b_x", "b_y", "b_hat_x", "b_hat_y", "b_mag")]
    H = np.zeros((nx, ny))
    emu.emu_field_heating(h, *[vp(a) for a in d], *[C.c_double(kw[k]) for k in ("coeff", "current_pow", "b_pow", "n_pow", "roc_pow")], C.c_int(0), C.c_double(step), vp(H))
    o.small_module_hooks(0, step)
    o.small_module_hooks(1, step)
    ref = o.small_module_plane(0, 0)
    assert np.count_nonzero(ref) > 0
    assert same_bits(H, ref), "heating plane: " + mismatch(H, ref)     # pow comes from the same libm on both sides here
    compare(emu, h, o, nx, ny, "field_heating")
    emu.emu_destroy(h); o.close()


@pytest.mark.parametrize("boundary,shape,fa,dyn", [("y_bound_2", "exp", False, False), ("x_bound_1", "gaussian", True, True), ("y_bound_1", "flat", True, False), ("x_bound_2", "exp", False, True)])
def test_boundary_outflow_kernels_equal_oracle(emu, boundary, shape, fa, dyn):
    nx, ny = 24, 21
    xb, yb = ("fixed", "fixed"), ("fixed", "fixed")
    s, o, h, step = make_pair(emu, xb, yb, nx, ny)
    codes = {"x_bound_1": 0.0, "x_bound_2": 1.0, "y_bound_1": 2.0, "y_bound_2": 3.0, "exp": 0.0, "gaussian": 1.0, "flat": 2.0}
    o.add_small_module("boundary_outflow", max_accel=3.0e4, falloff_length=6.0e8, boundary=codes[boundary], falloff_shape=codes[shape], feather_length=2.0e8, field_aligned_mode=float(fa),
                       dynamic_mode=float(dyn), dynamic_time=20.0, dynamic_target_speed=1.0e5)
    px = np.ascontiguousarray(s["planes"]["pos_x"], dtype=np.float64); py = np.ascontiguousarray(s["planes"]["pos_y"], dtype=np.float64)
    tmpl = np.zeros((nx, ny)); win = (C.c_int * 4)()
    emu.emu_outflow_setup(h, vp(px), vp(py), C.c_double(6.0e8), C.c_double(2.0e8), C.c_int(int(codes[boundary])), C.c_int(int(codes[shape])), vp(tmpl), win)
    assert same_bits(tmpl, o.small_module_plane(0, 0)), "acceleration template: " + mismatch(tmpl, o.small_module_plane(0, 0))
    ref_mean = o.outflow_mean(0)
    mean, accel = C.c_double(), C.c_double()
    emu.emu_boundary_outflow(h, vp(tmpl), win, C.c_int(int(codes[boundary])), C.c_int(int(fa)), C.c_int(int(dyn)), C.c_double(3.0e4), C.c_double(20.0), C.c_double(1.0e5),
                             C.c_double(step), C.byref(mean), C.byref(accel))
    assert mean.value.hex() == float(ref_mean).hex(), (mean.value, ref_mean)
    o.small_module_hooks(2, step)
    compare(emu, h, o, nx, ny, "boundary_outflow")
    emu.emu_destroy(h); o.close()



TC_CASES = [("euler", True), ("rk2", True), ("rk4", False), ("rk4", True)]


@pytest.mark.parametrize("integ,sat", TC_CASES)
def test_thermal_conduction_with_diagnostic_planes_equals_oracle(emu, integ, sat):
    """tc_iterate with output_to_file: energy, the avg-change plane and the saturation plane.  1e-9: the kernel forms T^2.5 as (T*T)*sqrt(T), the reference calls pow."""
    nx, ny = 24, 21
    xb, yb = ("periodic", "periodic"), ("fixed", "fixed")
    s, o, h, step = make_pair(emu, xb, yb, nx, ny)
    o.set_thermal_conduction(flux_saturation=sat, integrator=integ, epsilon=0.1, dt_subcycle_min=1.0e-4, weakening_factor=1.0)
    d = [np.ascontiguousarray(o.get(v)).copy() for v in ("temp", "b_hat_x", "b_hat_y")]
    # run the oracle's own step far enough to learn the sub-cycle count and the planes: pre + iterate happen at the start of step()
    e_before = o.get("thermal_energy").copy()
    o.step()
    ns = o.subcycles("thermal_conduction")
    assert ns >= 1
    avg = np.zeros((nx, ny)); satp = np.zeros((nx, ny))
    emu.emu_thermal_conduction(h, *[vp(a) for a in d], C.c_int(int(sat)), C.c_double(1.0), C.c_double(1.0e-4), C.c_int({"euler": 0, "rk2": 1, "rk4": 2}[integ]), C.c_int(ns), C.c_double(step),
                               vp(avg), vp(satp))
    ref_avg, ref_sat = o.module_output("thermal_conduction"), o.module_output("flux_saturation")
    assert np.count_nonzero(ref_avg) > 0
    assert np.max(np.abs(avg - ref_avg)) <= 1e-9 * np.max(np.abs(ref_avg)), np.max(np.abs(avg - ref_avg)) / np.max(np.abs(ref_avg))
    if sat:
        assert np.max(np.abs(satp - ref_sat)) <= 1e-9 * np.max(np.abs(ref_sat))
    emu.emu_destroy(h); o.close()


@pytest.mark.parametrize("integ", ["euler", "rk2", "rk4"])
def test_radiative_losses_with_diagnostic_plane_equals_oracle(emu, integ):
    nx, ny = 24, 21
    xb, yb = ("periodic", "periodic"), ("fixed", "fixed")
    s, o, h, step = make_pair(emu, xb, yb, nx, ny)
    o.set_radiative_losses(integrator=integ, cutoff_ramp=1.0e3, cutoff_temp=3.0e4, epsilon=0.1, prevent_subcycling=False)
    o.step()
    ns = o.subcycles("radiative_losses")
    avg = np.zeros((nx, ny))
    emu.emu_radiative_losses(h, C.c_int({"euler": 0, "rk2": 1, "rk4": 2}[integ]), C.c_double(1.0e3), C.c_double(3.0e4), C.c_double(0.1), C.c_int(0), C.c_int(ns), C.c_double(step), vp(avg))
    ref = o.module_output("rad")
    assert ns >= 1 and np.count_nonzero(ref) > 0
    assert np.max(np.abs(avg - ref)) <= 1e-9 * np.max(np.abs(ref)), np.max(np.abs(avg - ref)) / np.max(np.abs(ref))
    emu.emu_destroy(h); o.close()


@pytest.mark.parametrize("gvisc", [0.0, 0.3])
@pytest.mark.parametrize("xb,yb", [(("open_moc", "open_moc"), ("open_moc", "open_moc")), (("periodic", "periodic"), ("open_moc", "open_moc")), (("open_moc", "open_moc"), ("periodic", "periodic"))])
def test_open_moc_stage_kernels_equal_oracle(emu, xb, yb, gvisc):
    """k_moc_save / k_moc_visc_min / k_moc_stage (moc_stage.cuh) as launch_moc runs them: the exported right-hand side of the evolved ghost cells equals the oracle's,
    an euler stage leaves the oracle's new ghost-cell state, and the dt the kernel folds in is the minimum over those cells."""
    nx, ny = 22, 19
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    floors = dict(density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="euler", **floors)
    o.set_global_viscosity(gvisc)
    o.run(2)
    ev = [np.ascontiguousarray(o.get(v)).copy() for v in EV]
    st = [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in ST]
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    bc = (C.c_int * 4)(BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]])
    h = C.c_void_p(emu.emu_create(C.c_int(nx), C.c_int(ny), bc, C.c_double(s["ion_mass"]), C.c_double(s["adiabatic_index"]), C.c_double(floors["density_min"]), C.c_double(floors["temp_min"]),
                                  C.c_double(floors["thermal_energy_min"]), C.c_double(0.2), vp(dx), vp(dy), (C.c_void_p * 8)(*[a.ctypes.data for a in ev]), (C.c_void_p * 5)(*[a.ctypes.data for a in st])))
    ghost = np.ones((nx, ny), dtype=bool)
    ghost[(0 if xb[0] == "periodic" else 2):(nx if xb[0] == "periodic" else nx - 2), (0 if yb[0] == "periodic" else 2):(ny if yb[0] == "periodic" else ny - 2)] = False
    D = np.zeros((8, nx, ny)); K1 = np.zeros((8, nx, ny)); dtm = C.c_double()
    k_ref = o.rhs()
    emu.emu_moc_stage(h, C.c_int(5), C.c_double(1.0), C.c_double(0.0), C.c_int(0), C.c_double(gvisc), vp(D), vp(K1), C.byref(dtm))            # KM_EXPORT
    for v in range(8):
        assert np.array_equal(~np.isnan(K1[v]), ghost), "cells written by the export: %d of %d" % (np.count_nonzero(~np.isnan(K1[v])), np.count_nonzero(ghost))
        assert same_bits(K1[v][ghost], k_ref[v][ghost]), "d(%s)/dt on the evolved ghost cells: %s" % (EV[v], mismatch(K1[v][ghost], k_ref[v][ghost]))
    step = o.step()
    emu.emu_moc_stage(h, C.c_int(0), C.c_double(1.0), C.c_double(step), C.c_int(1), C.c_double(gvisc), vp(D), vp(K1), C.byref(dtm))         # KM_NONE, primary
    for v, name in enumerate(EV):
        assert same_bits(D[v][ghost], o.get(name)[ghost]), "%s on the evolved ghost cells after an euler stage: %s" % (name, mismatch(D[v][ghost], o.get(name)[ghost]))
    assert dtm.value.hex() == float(np.min(o.get("dt")[ghost])).hex()
    emu.emu_destroy(h); o.close()
