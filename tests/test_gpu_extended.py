"""GPU parity tests of the SURVEY 8f rows and of the stage-kernel build options: open_moc boundaries and their limiters, the small solar
modules, div_cleaning / field_heating, boundary_outflow, anomalous_resistivity, IdealMHD2E, the two-fluid set with open_ucnp next to wall
sides, the compile-time stage variants and the relaxed-arithmetic stage kernel.  All of them passed on a B200 in round 1 (GPUTEST_r01.json)
and are ordinary, strict tests now.  Each runs in its own interpreter (environment switches are read at domain creation; a faulting kernel
cannot poison the CUDA context of the other tests).  Every check is bit-for-bit against the CPU oracle or a reference fixture, like
tests/test_gpu_parity.py, except where a libm function or relaxed arithmetic sets the north star's 1e-9 tolerance."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


TIMEOUTS = []          # circuit breaker: a hung kernel costs `timeout` seconds of box time per test; after three, the rest of this file is skipped


def run_isolated(code: str, env: dict, timeout: int = 90):
    if len(TIMEOUTS) >= 3:
        pytest.skip("three isolated runs timed out (%s); the remaining unvalidated tests are skipped to bound the box time" % ", ".join(TIMEOUTS))
    e = dict(os.environ); e.update(env)
    e["PYTHONPATH"] = os.pathsep.join([str(ROOT), str(ROOT / "tests"), e.get("PYTHONPATH", "")])
    try:
        r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=str(ROOT), env=e, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        TIMEOUTS.append(os.environ.get("PYTEST_CURRENT_TEST", "?").split("::")[-1].split(" ")[0])
        raise
    assert r.returncode == 0, "exit %d\n%s\n%s" % (r.returncode, r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


STAGE_VARIANT_CODE = """
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    integ, zfull, nx, ny, xb, yb = {integ!r}, {zfull!r}, {nx}, {ny}, {xb!r}, {yb!r}
    s = synthetic.orszag_tang(nx, ny, zfull=zfull) if xb[0] == "periodic" else synthetic.stratified_loop(nx, ny)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
    if xb[0] != "periodic":
        kw = dict(xb=xb, yb=yb, integrator=integ, density_min=3.0e8)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    ref = o.run(7)
    dts = d.advance(7)
    assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
    for v in PlasmaDomain.EVOLVED + ["dt", "temp"]:
        assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    print("ok")
"""


@pytest.mark.parametrize("switches", [{"SPRUCE_STAGE_VARIANTS": "0"}, {"SPRUCE_VEC_ROWS": "0"}, {"SPRUCE_STAGE_VARIANTS": "0", "SPRUCE_VEC_ROWS": "0"}])
@pytest.mark.parametrize("integ,zfull,nx,ny,xb,yb", [
    ("rk2", False, 150, 203, ("periodic", "periodic"), ("periodic", "periodic")),      # 2-D list: k_mhd_stage_xy<6, ., 1> then <6, ., 2>
    ("rk2", True, 150, 203, ("periodic", "periodic"), ("periodic", "periodic")),       # full list: <12, ., 1> / <12, ., 2>
    ("euler", False, 131, 96, ("periodic", "periodic"), ("periodic", "periodic")),     # <6, ., 3>
    ("euler", True, 131, 96, ("periodic", "periodic"), ("periodic", "periodic")),      # <12, ., 3>
    ("rk2", False, 97, 140, ("open", "fixed"), ("reflect", "open")),                   # primary-stage strips + ghost passes
    ("rk4", False, 99, 77, ("periodic", "periodic"), ("periodic", "periodic")),        # K planes: the run-time-stage instance in every setting
])
def test_stage_kernel_build_options_vs_oracle(integ, zfull, nx, ny, xb, yb, switches):
    """The stage kernel's two switches, both ON by default (the default path is what tests/test_gpu_parity.py exercises):
    SPRUCE_STAGE_VARIANTS=0 runs the instances of k_mhd_stage_xy that take the integrator stage from the launch arguments instead of the
    template constant VAR; SPRUCE_VEC_ROWS=0 copies every ring row through the 8-byte per-column path instead of 16-byte chunks.  Same
    arithmetic: bit for bit against the oracle in every setting."""
    out = run_isolated(STAGE_VARIANT_CODE.format(integ=integ, zfull=zfull, nx=nx, ny=ny, xb=xb, yb=yb), switches)
    assert "ok" in out


MOC_CODE = """
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    xb, yb, integ, gvisc, nx, ny = {xb!r}, {yb!r}, {integ!r}, {gvisc!r}, {nx}, {ny}
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    o.set_global_viscosity(gvisc)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], eqs_options=dict(global_viscosity=gvisc), **kw)
    assert same_bits(d.grid("dt"), o.get("dt")), "dt after setup: " + mismatch(d.grid("dt"), o.get("dt"))
    k_dev, k_ref = d.computeTimeDerivatives(), o.rhs()
    for i, nm in enumerate(PlasmaDomain.EVOLVED):
        assert same_bits(k_dev[i], k_ref[i]), "d(%s)/dt: %s" % (nm, mismatch(k_dev[i], k_ref[i]))
    ref = o.run(6)
    dts = d.advance(6)
    assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
    for v in PlasmaDomain.EVOLVED + ["dt", "temp", "v_x"]:
        assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    print("ok")
"""


@pytest.mark.parametrize("xb,yb,integ,gvisc,nx,ny", [
    (("periodic", "periodic"), ("fixed", "open_moc"), "euler", 0.0, 126, 93),
    (("periodic", "periodic"), ("open_moc", "fixed"), "rk2", 0.3, 64, 125),
    (("open_moc", "reflect"), ("fixed", "open"), "rk4", 0.0, 97, 72),
    (("open_moc", "open_moc"), ("open_moc", "open_moc"), "rk2", 0.1, 85, 134),
    (("fixed", "open_moc"), ("open", "open_moc"), "euler", 0.0, 73, 66),
    (("open_moc", "open_moc"), ("periodic", "periodic"), "rk4", 0.2, 70, 130),
])
def test_open_moc_vs_oracle(xb, yb, integ, gvisc, nx, ny):
    """The method-of-characteristics open boundary on the device (moc_stage.cuh: k_moc_stage, k_moc_visc_min) against the CPU
    restatement that is pinned to live reference runs: right-hand side incl. the evolved ghost cells, step-size history with the
    widened bounds, every plane.  The per-cell arithmetic is already proven on the host (tests/test_moc_host_check.py); this is the
    launch side: thread-to-cell mapping, corner ownership, K planes of rk4, order against the stage kernel and the ghost passes."""
    out = run_isolated(MOC_CODE.format(xb=xb, yb=yb, integ=integ, gvisc=gvisc, nx=nx, ny=ny), {})
    assert "ok" in out


SOURCE_CODE = """
    import math
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    mods, xb, yb, integ, nx, ny = {mods!r}, {xb!r}, {yb!r}, {integ!r}, {nx}, {ny}
    s = synthetic.stratified_loop(nx, ny)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    for name, a in mods:
        o.add_small_module(name, **a)
        if name == "ambient_heating_sink":            # AmbientHeatingSink::setupModule (ambientheatingsink.cpp:27-33), host libm
            X, Y = s["planes"]["pos_x"], s["planes"]["pos_y"]
            mask = np.zeros((nx, ny))
            il = lambda b: 0 if b == "periodic" else 2
            ih = lambda b, n: n if b == "periodic" else n - 2
            mask[il(xb[0]):ih(xb[1], nx), il(yb[0]):ih(yb[1], ny)] = 1.0
            if a.get("exp_mode"):
                ex = np.vectorize(math.exp)((-1.0 * Y) / a["exp_scale_height"])
                q = (X - a["center_x"]) / a["half_width"]
                red = ((mask * a["exp_base_heating_rate"]) * ex) * np.maximum(1.0 - q * q, 0.0)
            else:
                red = mask * a["heating_rate"]
            d.set_ambient_heating_sink_plane(red)
        else:
            b = dict(a)
            if "oscillatory" in b: b["oscillatory"] = bool(b["oscillatory"])
            getattr(d, "set_" + name)(**b)
    ref = o.run(8)
    dts = d.advance(8)
    assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
    for v in PlasmaDomain.EVOLVED + ["dt", "temp"]:
        assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    print("ok")
"""


@pytest.mark.parametrize("mods,xb,yb,integ,nx,ny", [
    ([("ambient_heating_sink", dict(heating_rate=1.0e-5)),
      ("localized_heating", dict(start_time=0.0, duration=5.0, max_heating_rate=1.0e-3, stddev_x=3.0, stddev_y=4.0, center_x=10.0, center_y=8.0, ramp_time=1.0))],
     ("periodic", "periodic"), ("fixed", "fixed"), "rk2", 96, 70),
    ([("ambient_heating_sink", dict(exp_mode=1.0, exp_base_heating_rate=2.0e-5, exp_scale_height=8.0e8, center_x=2.0e9, half_width=1.5e9)),
      ("mass_injection", dict(start_time=0.5, duration=10.0, max_injection_rate=1.0e6, stddev_x=3.0, stddev_y=3.0, center_x=12.0, center_y=10.0))],
     ("reflect", "open"), ("fixed", "open"), "euler", 75, 88),
    ([("momentum_injection", dict(start_time=0.0, duration=50.0, max_accel=1.0e3, stddev_x=4.0, stddev_y=3.0, center_x=13.0, center_y=9.0, dir_x=1.0, dir_y=0.5,
                                  template_angle=20.0, oscillatory=1.0, oscillation_period=3.0))],
     ("fixed", "open"), ("reflect", "fixed"), "rk2", 81, 64),
    ([("momentum_injection", dict(start_time=0.0, duration=50.0, max_accel=1.0e3, stddev_x=4.0, stddev_y=3.0, center_x=13.0, center_y=9.0, dir_x=-1.0, dir_y=0.5,
                                  template_angle=0.0, oscillatory=0.0, oscillation_period=1.0)),
      ("localized_heating", dict(start_time=0.3, duration=2.0, max_heating_rate=5.0e-4, stddev_x=5.0, stddev_y=2.0, center_x=1.0, center_y=60.0, ramp_time=0.4))],
     ("periodic", "periodic"), ("periodic", "periodic"), "rk4", 66, 67),
])
def test_pointwise_solar_source_terms_vs_oracle(mods, xb, yb, integ, nx, ny):
    """ambient_heating_sink, localized_heating, mass_injection, momentum_injection on the device (k_source_term + the host-built templates
    of solar_templates.hpp, which tests/test_solar_templates_host_check.py proves bit-identical on the host) against the pinned oracle."""
    out = run_isolated(SOURCE_CODE.format(mods=mods, xb=xb, yb=yb, integ=integ, nx=nx, ny=ny), {})
    assert "ok" in out


DC_FH_CODE = """
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    mods, xb, yb, integ, nx, ny, exact = {mods!r}, {xb!r}, {yb!r}, {integ!r}, {nx}, {ny}, {exact!r}
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    for name, a in mods:
        o.add_small_module(name, **a)
        b = dict(a)
        if "inactive_mode" in b: b["inactive_mode"] = bool(b["inactive_mode"])
        getattr(d, "set_" + name)(**b)
    rel = lambda x, y: float(np.max(np.abs(x - y)) / max(np.max(np.abs(y)), 1e-300))
    ref = o.run(6)
    dts = d.advance(6)
    if exact:
        assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
    else:
        assert np.max(np.abs(np.array(dts) - np.array(ref)) / np.array(ref)) <= 1e-9
    for v in PlasmaDomain.EVOLVED + ["dt", "temp"]:
        if exact:
            assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
        else:
            assert rel(d.grid(v), o.get(v)) <= 1e-9, "%s: rel Linf %.3e" % (v, rel(d.grid(v), o.get(v)))
    print("ok")
"""


@pytest.mark.parametrize("mods,xb,yb,integ,nx,ny,exact", [
    ([("div_cleaning", dict(epsilon=0.1, time_scale=5.0))], ("fixed", "open"), ("reflect", "fixed"), "rk2", 81, 64, True),
    ([("div_cleaning", dict(epsilon=0.05, time_scale=2.0))], ("periodic", "periodic"), ("fixed", "fixed"), "euler", 70, 93, True),
    ([("field_heating", dict(coeff=1.0, current_pow=1.0, b_pow=0.5, n_pow=0.25, roc_pow=0.5, inactive_mode=0.0))], ("periodic", "periodic"), ("fixed", "open"), "rk2", 77, 66, False),
    ([("field_heating", dict(coeff=3.0e-3, current_pow=0.0, b_pow=2.0, n_pow=0.0, roc_pow=0.0, inactive_mode=0.0)), ("div_cleaning", dict(epsilon=0.1, time_scale=5.0))],
     ("reflect", "reflect"), ("open", "fixed"), "rk4", 65, 72, False),
])
def test_div_cleaning_and_field_heating_vs_oracle(mods, xb, yb, integ, nx, ny, exact):
    """div_cleaning (whole-plane operator passes + k_dc_update; bit for bit) and field_heating (k_fh_compute / k_fh_apply; pow with run-time
    exponents, CUDA vs glibc: <= 1e-9) against the pinned oracle."""
    out = run_isolated(DC_FH_CODE.format(mods=mods, xb=xb, yb=yb, integ=integ, nx=nx, ny=ny, exact=exact), {})
    assert "ok" in out


OUTFLOW_CODE = """
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    a, xb, yb, integ, nx, ny = {a!r}, {xb!r}, {yb!r}, {integ!r}, {nx}, {ny}
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    B = dict(x_bound_1=0, x_bound_2=1, y_bound_1=2, y_bound_2=3); S = dict(exp=0, gaussian=1, flat=2)
    o.add_small_module("boundary_outflow", max_accel=a["max_accel"], falloff_length=a["falloff_length"], boundary=B[a["boundary"]], falloff_shape=S[a["falloff_shape"]],
                       feather_length=a["feather_length"], field_aligned_mode=float(a["field_aligned_mode"]), dynamic_mode=float(a["dynamic_mode"]),
                       dynamic_time=a["dynamic_time"], dynamic_target_speed=a["dynamic_target_speed"])
    d.set_boundary_outflow(s["planes"]["pos_x"], s["planes"]["pos_y"], **a)
    ref = o.run(7)
    dts = d.advance(7)
    assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
    for v in PlasmaDomain.EVOLVED + ["dt", "temp"]:
        assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    print("ok")
"""


@pytest.mark.parametrize("a,xb,yb,integ,nx,ny,moc", [
    (dict(max_accel=2.0e3, falloff_length=6.0e8, boundary="y_bound_2", falloff_shape="exp", feather_length=3.0e8, field_aligned_mode=True, dynamic_mode=True,
          dynamic_time=10.0, dynamic_target_speed=2.0e6), ("periodic", "periodic"), ("fixed", "open"), "rk2", 88, 71, False),
    (dict(max_accel=1.0e3, falloff_length=4.0e8, boundary="x_bound_1", falloff_shape="flat", feather_length=2.0e8, field_aligned_mode=False, dynamic_mode=False,
          dynamic_time=1.0, dynamic_target_speed=0.0), ("open", "fixed"), ("reflect", "reflect"), "euler", 67, 90, False),
    (dict(max_accel=2.0e3, falloff_length=5.0e8, boundary="y_bound_2", falloff_shape="gaussian", feather_length=0.0, field_aligned_mode=True, dynamic_mode=True,
          dynamic_time=5.0, dynamic_target_speed=1.0e6), ("periodic", "periodic"), ("fixed", "open_moc"), "rk2", 72, 69, True),
])
def test_boundary_outflow_vs_oracle(a, xb, yb, integ, nx, ny, moc):
    """boundary_outflow on the device (k_bo_mean, k_bo_apply; template and window built by solar_templates.hpp, host-checked) against the pinned oracle;
    the last case combines it with an open_moc side, whose ghost zone the template reaches into."""
    out = run_isolated(OUTFLOW_CODE.format(a=a, xb=xb, yb=yb, integ=integ, nx=nx, ny=ny), {})
    assert "ok" in out


RELAXED_CODE = """
    import numpy as np
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    integ, zfull, nx, ny, nsteps = {integ!r}, {zfull!r}, {nx}, {ny}, {nsteps}
    s = synthetic.orszag_tang(nx, ny, zfull=zfull)
    kw = dict(xb=("periodic", "periodic"), yb=("periodic", "periodic"), integrator=integ, density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    ref = np.array(o.run(nsteps)); dts = np.array(d.advance(nsteps))
    assert len(dts) == len(ref)                                           # identical step count
    TOL = 1.0e-9                                                          # the north star's stated FP64 tolerance after 100 steps
    assert np.max(np.abs(dts - ref) / ref) <= TOL, float(np.max(np.abs(dts - ref) / ref))
    worst = 0.0
    for v in PlasmaDomain.EVOLVED + ["dt", "temp"]:
        a, b = d.grid(v), o.get(v)
        r = float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))
        assert r <= TOL, "%s: rel Linf %.3e" % (v, r)
        worst = max(worst, r)
    assert np.any(dts != ref) or worst > 0.0, "relaxed arithmetic reproduced the reference bit for bit: the relaxed kernel was probably not the one that ran"
    print("ok worst rel Linf %.3e, dt rel %.3e" % (worst, float(np.max(np.abs(dts - ref) / ref))))
"""


@pytest.mark.parametrize("integ,zfull,nx,ny,nsteps,variants", [
    ("rk2", False, 256, 256, 100, "0"),
    ("rk2", False, 256, 256, 100, "1"),
    ("rk2", True, 150, 203, 100, "1"),
    ("rk4", False, 131, 96, 60, "0"),
    ("euler", True, 96, 131, 60, "1"),
])
def test_relaxed_arithmetic_within_north_star_tolerance(integ, zfull, nx, ny, nsteps, variants):
    """SPRUCE_ARITH=relaxed (stage_relaxed.cu: FMA contraction, one-multiplication table divisions): identical step count, step sizes and every
    field within 1e-9 relative L-infinity of the oracle after 100 steps -- the north star's stated floating-point bar; the default build stays
    bit-identical.  Also checks that the result is NOT bit-identical, i.e. that the relaxed kernel really ran."""
    out = run_isolated(RELAXED_CODE.format(integ=integ, zfull=zfull, nx=nx, ny=ny, nsteps=nsteps), {"SPRUCE_ARITH": "relaxed", "SPRUCE_STAGE_VARIANTS": variants}, timeout=240)
    assert "ok" in out


@pytest.mark.parametrize("args", [("96", "80", "4", "rk2", "periodic", "p2p", "moc"), ("90", "70", "4", "rk4", "open_moc", "p2p", "mocv"),
                                  ("96", "80", "3", "euler", "open_moc", "nccl", "moc"), ("128", "96", "4", "rk2", "periodic", "p2p", "src"),
                                  ("120", "90", "3", "rk2", "periodic", "p2p", "dc,fh"), ("96", "80", "4", "rk2", "reflect", "p2p", "bo,dc"),
                                  ("110", "84", "4", "rk4", "periodic", "p2p", "2e"), ("96", "80", "4", "rk2", "reflect", "p2p", "2e")])
def test_slab_decomposition_of_the_unvalidated_paths_equals_single_gpu(args):
    """open_moc sides and the pointwise solar source terms on 2 slabs == 1 GPU, bit for bit (the slab form of the open_moc evaluation is proven on the
    host by tests/test_moc_host_check.py; this is the launch side).  Needs >= 2 visible GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    e = dict(os.environ)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29519", str(ROOT / "scripts" / "mgpu_check.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300, env=e)
    out = r.stdout.decode()
    assert r.returncode == 0 and "IDENTICAL" in out, out[-3000:]


GOLDEN_CODE = """
    import numpy as np
    from golden_util import Golden, OUT_VARS, same_bits, mismatch, small_module_kwargs, sink_reduction_plane
    from spruce_b200.domain import PlasmaDomain
    g = Golden({name!r})
    exact = {exact!r}
    opts = dict(global_viscosity=float(g.eqs_raw["global_viscosity"])) if "global_viscosity" in g.eqs_raw else None
    d = PlasmaDomain(g.planes, g.ion_mass, g.adiabatic_index, equation_set=g.equation_set, eqs_options=opts, **g.kw)
    for mname, kv in g.modules:
        prod = small_module_kwargs(mname, kv)[1]
        if mname == "ambient_heating_sink":
            d.set_ambient_heating_sink_plane(sink_reduction_plane(g.planes, prod, g.kw["xb"], g.kw["yb"]))
        elif mname == "boundary_outflow":
            d.set_boundary_outflow(g.planes["pos_x"], g.planes["pos_y"], **prod)
        else:
            getattr(d, "set_" + mname)(**prod)
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))
    done = 0
    for it in sorted(g.frames):
        dts = d.advance(it - done)
        ref = g.steps[done:it]
        if exact:
            assert all(a == b for a, b in zip(dts, ref)), ([x.hex() for x in dts], [float(x).hex() for x in ref])
        else:
            assert np.max(np.abs(np.array(dts) - ref) / ref) <= 1e-9
        done = it
        for v in g.out_vars:
            got = d.grid(v)
            if exact:
                assert same_bits(got, g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(got, g.frames[it][v]))
            else:
                assert rel(got, g.frames[it][v]) <= 1e-9, "%s after iteration %d: %.3e" % (v, it, rel(got, g.frames[it][v]))
    print("ok")
"""


@pytest.mark.parametrize("name,exact", [("moc_y2_euler", True), ("moc_all_visc_rk2", True), ("moc_x1_mixed_rk4", True), ("sm_sink_heat_mass_rk2", True),
                                        ("sm_momentum_divclean_rk2", True), ("sm_field_heating_euler", False), ("sm_outflow_dynamic_rk2", True),
                                        ("e2_mixed_rk2", True), ("e2_pp_ucnp_rk4", True)])
def test_extended_golden_reference_outputs(name, exact):
    """The device paths of the SURVEY 8f rows against committed outputs of the UNMODIFIED reference binary (tests/golden/moc_*, sm_*): step-size
    history and every output plane, bit for bit (field_heating: pow with run-time exponents, <= 1e-9)."""
    out = run_isolated(GOLDEN_CODE.format(name=name, exact=exact), {})
    assert "ok" in out


SOLAR = dict(density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
HOST_CASES = {
    "moc_visc": (dict(nx=40, ny=36, bump=0.4), dict(integrator="rk2", xb=("open_moc", "open_moc"), yb=("fixed", "open_moc"), eqs_block=[("global_viscosity", "0.1")], max_iterations=5,
                 iter_output_interval=1, write_precision=17, **SOLAR), True),
    "source_terms": (dict(nx=40, ny=36), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), max_iterations=6, iter_output_interval=2, write_precision=17, modules=[
        ("ambient_heating_sink", [("exp_mode", "true"), ("exp_base_heating_rate", "2.0e-5"), ("exp_scale_height", "8.0e8"), ("center_x", "2.0e9"), ("half_width", "1.5e9")]),
        ("localized_heating", [("start_time", "0.0"), ("duration", "5.0"), ("max_heating_rate", "1.0e-3"), ("stddev_x", "3.0"), ("stddev_y", "4.0"), ("center_x", "2.0"), ("center_y", "8.0"), ("ramp_time", "1.0")]),
        ("mass_injection", [("start_time", "0.5"), ("duration", "10.0"), ("max_injection_rate", "1.0e6"), ("stddev_x", "3.0"), ("stddev_y", "3.0"), ("center_x", "12.0"), ("center_y", "10.0")]),
        ("momentum_injection", [("start_time", "0.0"), ("duration", "50.0"), ("max_accel", "1.0e3"), ("stddev_x", "4.0"), ("stddev_y", "3.0"), ("center_x", "13.0"), ("center_y", "9.0"),
                                ("dir_x", "1.0"), ("dir_y", "0.5"), ("template_angle", "20.0"), ("oscillatory", "true"), ("oscillation_period", "3.0")])], **SOLAR), True),
    "divclean_outflow": (dict(nx=40, ny=36, bump=0.4), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "open"), max_iterations=5, iter_output_interval=1, write_precision=17, modules=[
        ("div_cleaning", [("epsilon", "0.1"), ("time_scale", "5.0")]),
        ("boundary_outflow", [("max_accel", "2.0e3"), ("falloff_length", "6.0e8"), ("boundary", "y_bound_2"), ("falloff_shape", "exp"), ("feather_length", "3.0e8"),
                              ("field_aligned_mode", "true"), ("dynamic_mode", "true"), ("dynamic_time", "10.0"), ("dynamic_target_speed", "2.0e6")])], **SOLAR), True),
    "ideal_mhd_2e": (dict(nx=40, ny=36, two_energy=True), dict(integrator="rk2", xb=("reflect", "open"), yb=("fixed", "open"), eqs="ideal_mhd_2E", max_iterations=6, iter_output_interval=2,
                     write_precision=17, output_flags=("rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "i_thermal_energy", "e_thermal_energy", "press", "n", "v_x", "b_mag", "dt"),
                     **SOLAR), True),
    "field_heating": (dict(nx=40, ny=36, bump=0.5), dict(integrator="euler", xb=("periodic", "periodic"), yb=("fixed", "open"), max_iterations=4, iter_output_interval=1, modules=[
        ("field_heating", [("coeff", "1.0"), ("current_pow", "1.0"), ("b_pow", "0.5"), ("n_pow", "0.25"), ("roc_pow", "0.5")])], **SOLAR), False),
}


@pytest.mark.parametrize("name", list(HOST_CASES))
def test_host_shell_matches_reference_files_on_the_8f_rows(name, tmp_path):
    """The C++ host shell (spruce_b200/bin/run) with open_moc sides / the solar modules of SURVEY 8f against the UNMODIFIED reference binary on the same
    .state / .config: mhd.out and end.state byte-identical (field_heating: within 1e-9)."""
    import numpy as np
    from oracle import refrun
    from spruce_b200 import synthetic
    ours = ROOT / "spruce_b200" / "bin" / "run"
    if not ours.exists():
        subprocess.run(["make", "-C", str(ROOT / "spruce_b200" / "host")], check=True)
    assert refrun.have_reference(), "oracle/_ref/run must travel with the repo"
    gkw, ckw, exact = HOST_CASES[name]
    if gkw.get("two_energy"):
        from test_oracle_vs_live_reference import e2_state
        s = e2_state(gkw["nx"], gkw["ny"], True)
    else:
        s = synthetic.stratified_loop(**gkw)
    state = tmp_path / "in.state"
    refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"], comments=["# drop-in test " + name])
    cfg = refrun.ideal_mhd_config(std_out_interval=1, **ckw)
    refrun.run_reference(state, cfg, tmp_path / "ref", threads=4)
    out_dir = tmp_path / "ours"
    out_dir.mkdir(parents=True, exist_ok=True)
    (out_dir / "run.config").write_text(cfg)
    e = dict(os.environ)
    r = subprocess.run([str(ours), "-m", "input", "-o", str(out_dir), "-s", str(state)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120, env=e)
    assert r.returncode in (-6, 134) and "Simulation successfully reached max simulation time or iterations" in r.stderr.decode(), r.stderr.decode()[-2000:]
    for fname in ("mhd.out", "end.state"):
        a, b = (out_dir / fname).read_bytes(), (tmp_path / "ref" / fname).read_bytes()
        if exact:
            assert a == b, "%s differs from the reference's (%d vs %d bytes)" % (fname, len(a), len(b))
    if not exact:
        ma, pa = refrun.read_state(out_dir / "end.state")
        mb, pb = refrun.read_state(tmp_path / "ref" / "end.state")
        assert ma["t"] == mb["t"] and list(pa) == list(pb)
        for k in pb:
            assert np.max(np.abs(pa[k] - pb[k])) <= 1e-9 * max(np.max(np.abs(pb[k])), 1e-300), k


E2_CODE = """
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import EVOLVED_2E, Oracle2E
    from spruce_b200.domain import PlasmaDomain
    from test_oracle_vs_live_reference import e2_state
    xb, yb, integ, nx, ny, loop, nmin = {xb!r}, {yb!r}, {integ!r}, {nx}, {ny}, {loop!r}, {nmin!r}
    s = e2_state(nx, ny, loop)
    floors = dict(density_min=nmin, temp_min=1.0e4, thermal_energy_min=1.0e-6) if loop else dict(density_min=1.0, temp_min=1.0, thermal_energy_min=1.0e-30)
    kw = dict(xb=xb, yb=yb, integrator=integ, **floors)
    o = Oracle2E(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], equation_set="ideal_mhd_2E", **kw)
    for v in ("i_thermal_energy", "e_thermal_energy", "rho", "dt", "i_temp", "press", "v_x", "b_hat_y", "kinetic_energy"):
        assert same_bits(d.grid(v), o.get(v)), "after setup, %s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    k_dev, k_ref = d.computeTimeDerivatives(), o.rhs()
    for i, nm in enumerate(EVOLVED_2E):
        assert same_bits(k_dev[i], k_ref[i]), "d(%s)/dt: %s" % (nm, mismatch(k_dev[i], k_ref[i]))
    ref = [o.step() for _ in range(6)]
    dts = d.advance(6)
    assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
    for v in EVOLVED_2E + ["dt", "e_temp", "n", "b_mag"]:
        assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    print("ok")
"""


@pytest.mark.parametrize("xb,yb,integ,nx,ny,loop,nmin", [
    (("periodic", "periodic"), ("periodic", "periodic"), "rk2", 150, 133, False, 1.0),
    (("periodic", "periodic"), ("fixed", "fixed"), "rk4", 97, 140, True, 1.0e7),
    (("reflect", "open"), ("fixed", "open"), "rk2", 131, 96, True, 3.0e8),
    (("open_ucnp", "open_ucnp"), ("open", "reflect"), "euler", 66, 131, True, 1.0e7),
    (("open", "open"), ("periodic", "periodic"), "rk4", 90, 77, True, 1.0e7),
])
def test_ideal_mhd_2e_vs_oracle(xb, yb, integ, nx, ny, loop, nmin):
    """The IdealMHD2E equation set on the device (mhd2e_host.cuh) against its pinned CPU restatement: setup, right-hand side, step sizes, every evolved
    and several derived planes bit for bit.  The arithmetic and the stage order are already proven on the host (tests/test_mhd2e_host_check.py)."""
    out = run_isolated(E2_CODE.format(xb=xb, yb=yb, integ=integ, nx=nx, ny=ny, loop=loop, nmin=nmin), {})
    assert "ok" in out


E2_EIC_CODE = """
    import numpy as np
    from oracle.oracle import EVOLVED_2E, Oracle2E
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    xb, yb, integ, nx, ny, drift, bfield = {xb!r}, {yb!r}, {integ!r}, {nx}, {ny}, {drift!r}, {bfield!r}
    s = synthetic.ucnp_cloud_2e(nx, ny, drift=drift, bfield=bfield)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1.0e-30)
    o = Oracle2E(s["planes"], s["ion_mass"], s["adiabatic_index"], eic=True, **kw)
    plain = Oracle2E(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], equation_set="ideal_mhd_2E", **kw)
    d.set_eic_thermalization()
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))
    k_dev, k_ref = d.computeTimeDerivatives(), o.rhs()
    for i, nm in enumerate(EVOLVED_2E):
        assert rel(k_dev[i], k_ref[i]) <= 1e-9, "d(%s)/dt: %.3e" % (nm, rel(k_dev[i], k_ref[i]))
    ref = np.array([o.step() for _ in range(6)])
    for _ in range(6):
        plain.step()
    dts = np.array(d.advance(6))
    assert np.max(np.abs(dts - ref) / ref) <= 1e-9, (dts, ref)
    for v in EVOLVED_2E + ["dt", "e_temp", "i_temp", "n"]:
        assert rel(d.grid(v), o.get(v)) <= 1e-9, "%s: %.3e" % (v, rel(d.grid(v), o.get(v)))
    for v in ("i_thermal_energy", "e_thermal_energy"):
        assert rel(d.grid(v), plain.get(v)) > 100 * max(rel(d.grid(v), o.get(v)), 1e-15), "the exchange term is not visible in " + v
    print("ok")
"""


@pytest.mark.xfail(reason="written after round 2's GPU budget was spent (CPU-checked: tests/test_mhd2e_host_check.py, test_mhd2e_kernels_emulated.py); first executed by the round-end run", strict=False)
@pytest.mark.parametrize("xb,yb,integ,nx,ny,drift,bfield", [
    (("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp"), "rk2", 131, 97, 20.0, 0.01),
    (("fixed", "reflect"), ("open_ucnp", "fixed"), "rk4", 90, 133, 10.0, 0.02),
    (("periodic", "periodic"), ("periodic", "periodic"), "euler", 140, 66, 0.0, 0.01),
])
def test_ideal_mhd_2e_with_eic_thermalization_vs_oracle(xb, yb, integ, nx, ny, drift, bfield):
    """ideal_mhd_2E + eic_thermalization on the device (the UCNP configuration; Geo::eic in mhd2e_cells.cuh) against the restatement that live reference runs pin bit for
    bit: right-hand side, step sizes, evolved and derived planes within the module's 1e-9 (cbrt / x sqrt(x) / CUDA log against glibc pow / log), the term visibly acting."""
    out = run_isolated(E2_EIC_CODE.format(xb=xb, yb=yb, integ=integ, nx=nx, ny=ny, drift=drift, bfield=bfield), {})
    assert "ok" in out


@pytest.mark.xfail(reason="written after round 2's GPU budget was spent; first executed by the round-end run", strict=False)
def test_ideal_mhd_2e_with_eic_golden_reference_outputs():
    """the same configuration against a committed fixture of the UNMODIFIED reference binary (tests/golden/e2_ucnp_eic_rk2.npz), <= 1e-9"""
    out = run_isolated(GOLDEN_CODE.format(name="e2_ucnp_eic_rk2", exact=False), {})
    assert "ok" in out


E2_VISC_CODE = """
    import numpy as np
    from golden_util import Golden, same_bits, mismatch, module_kwargs, viscosity_terms_with_profiles
    from spruce_b200.domain import PlasmaDomain
    g = Golden("e2_ucnp_visc_rk2")
    d = PlasmaDomain(g.planes, g.ion_mass, g.adiabatic_index, equation_set=g.equation_set, **g.kw)
    for mname, kv in g.modules:
        kw = module_kwargs(mname, kv)
        d.set_viscosity(viscosity_terms_with_profiles(g.planes, kw.pop("terms")), **kw)
    done = 0
    for it in sorted(g.frames):
        dts = d.advance(it - done)
        ref = g.steps[done:it]
        assert all(a == b for a, b in zip(dts, ref)), ([x.hex() for x in dts], [float(x).hex() for x in ref])
        done = it
        for v in g.out_vars:
            got = d.grid(v)
            assert same_bits(got, g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(got, g.frames[it][v]))
    print("ok")
"""


@pytest.mark.xfail(reason="written after round 2's GPU budget was spent (CPU-checked: tests/test_mhd2e_kernels_emulated.py); first executed by the round-end run", strict=False)
def test_ideal_mhd_2e_with_artificial_viscosity_golden_reference_outputs():
    """artificial_viscosity on ideal_mhd_2E on the device (visc_cell in mhd2e_cells.cuh; right-hand-side terms and hyper-viscous rk2 sub-cycles, gradient correction) against a
    committed fixture of the UNMODIFIED reference binary (tests/golden/e2_ucnp_visc_rk2.npz): step sizes and every output plane bit for bit"""
    out = run_isolated(E2_VISC_CODE, {})
    assert "ok" in out


MOC_LIMIT_CODE = """
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    xb, yb, integ, gvisc, lim, nx, ny = {xb!r}, {yb!r}, {integ!r}, {gvisc!r}, {lim!r}, {nx}, {ny}
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], moc_limiting=lim, **kw)
    o.set_global_viscosity(gvisc)
    opts = dict(global_viscosity=gvisc, moc_b_limiting=lim.get("b_limiting", False), moc_b_lower_lim=lim.get("b_lower", 0.1), moc_b_upper_lim=lim.get("b_upper", 10.0),
                moc_mom_limiting=lim.get("mom_limiting", False), moc_mom_lower_lim=lim.get("mom_lower", 0.1), moc_mom_upper_lim=lim.get("mom_upper", 10.0))
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], eqs_options=opts, **kw)
    for v in PlasmaDomain.EVOLVED + ["dt"]:
        assert same_bits(d.grid(v), o.get(v)), "after setup, %s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    ref = o.run(6)
    dts = d.advance(6)
    assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
    for v in PlasmaDomain.EVOLVED + ["dt", "temp", "v_x"]:
        assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    print("ok")
"""


@pytest.mark.parametrize("xb,yb,integ,gvisc,lim,nx,ny", [
    (("periodic", "periodic"), ("fixed", "open_moc"), "rk2", 0.0, dict(b_limiting=True, b_lower=0.9, b_upper=1.05, mom_limiting=True, mom_lower=0.5, mom_upper=1.5), 86, 93),
    (("open_moc", "open_moc"), ("open_moc", "open_moc"), "euler", 0.1, dict(mom_limiting=True, mom_lower=0.8, mom_upper=1.1), 85, 134),
    (("open_moc", "reflect"), ("fixed", "open"), "rk4", 0.0, dict(b_limiting=True, b_lower=-0.5, b_upper=1.02), 97, 72),
])
def test_open_moc_limiters_vs_oracle(xb, yb, integ, gvisc, lim, nx, ny):
    """moc_b_limiting / moc_mom_limiting on the device (k_moc_limit after the boundary passes of every propagate, the dt minimum rebuilt in the last stage)
    against the pinned oracle; limit_line itself is proven on the host (tests/test_moc_host_check.py)."""
    out = run_isolated(MOC_LIMIT_CODE.format(xb=xb, yb=yb, integ=integ, gvisc=gvisc, lim=lim, nx=nx, ny=ny), {})
    assert "ok" in out


TF_MIXED_CODE = """
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import EVOLVED_2F, Oracle2F
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    xb, yb, integ, nx, ny = {xb!r}, {yb!r}, {integ!r}, {nx}, {ny}
    s = synthetic.ucnp_cloud(nx, ny, drift=2.0e3, bfield=5.0)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1e-30)
    o = Oracle2F(s["planes"], s["ion_mass"], s["adiabatic_index"], remove_curl_terms=False, eic=False, **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], equation_set="ideal_2F", eqs_options=dict(use_sub_cycling=False), **kw)
    for v in ("i_thermal_energy", "e_thermal_energy", "dt", "i_mom_x", "e_mom_y"):
        assert same_bits(d.grid(v), o.get(v)), "after setup, %s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    ref = [o.step() for _ in range(6)]
    dts = d.advance(6)
    assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
    for v in EVOLVED_2F + ["dt", "dt_i", "e_temp"]:
        assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    print("ok")
"""


@pytest.mark.parametrize("xb,yb,integ,nx,ny", [
    (("open_ucnp", "open_ucnp"), ("reflect", "reflect"), "rk2", 97, 81),      # the case the pointwise form gets wrong: a ucnp pass before a reflect side
    (("reflect", "open_ucnp"), ("fixed", "open_ucnp"), "rk4", 66, 91),
    (("open_ucnp", "fixed"), ("open_ucnp", "reflect"), "euler", 80, 75),
    (("periodic", "periodic"), ("open_ucnp", "reflect"), "rk2", 70, 88),
])
def test_two_fluid_ucnp_next_to_wall_sides_vs_oracle(xb, yb, integ, nx, ny):
    """Ideal2F with an open_ucnp side next to fixed / reflect sides (refused until now): the literal, ordered boundary passes of ideal2f_sides.cuh for the
    primary state -- proven on the host (tests/test_ideal2f_sides_host_check.py) -- against the two-fluid restatement, bit for bit."""
    out = run_isolated(TF_MIXED_CODE.format(xb=xb, yb=yb, integ=integ, nx=nx, ny=ny), {})
    assert "ok" in out


MODULE_PLANES_CODE = """
    import numpy as np
    from golden_util import Golden, module_kwargs
    from ambient import heating_plane
    from spruce_b200.domain import PlasmaDomain
    g = Golden({name!r})
    d = PlasmaDomain(g.planes, g.ion_mass, g.adiabatic_index, **g.kw)
    for mname, kv in g.modules:
        kw = module_kwargs(mname, kv)
        if mname == "ambient_heating":
            d.set_ambient_heating_plane(heating_plane(g, kw))
        else:
            getattr(d, "set_" + mname)(**kw)
        if kv.get("output_to_file") == "true":
            d.set_module_output_to_file(mname)
    done = 0
    for it in sorted(g.frames):
        d.advance(it - done); done = it
        for pname, ref in g.module_planes[it].items():
            got = d.module_output(pname)
            scale = max(float(np.max(np.abs(ref))), 1e-300)
            err = float(np.max(np.abs(got - ref))) / scale
            # a rate of change formed from the difference of two nearly equal energies: the module's 1e-9 agreement on e is amplified by e/|de|
            assert err <= 1.0e-6, "%s after iteration %d: rel Linf %.3e" % (pname, it, err)
            assert np.count_nonzero(ref) == 0 or np.count_nonzero(got) > 0
    print("ok")
"""


@pytest.mark.parametrize("name", ["loop_rl_euler", "loop_rl_rk4", "loop_tc_euler", "loop_tc_sat_rk2", "loop_tc_sat_rk4", "loop_solar_all"])
def test_module_diagnostic_planes_vs_reference_fixtures(name):
    """output_to_file = true of thermal_conduction / radiative_losses: the planes the reference appends to mhd.out ("thermal_conduction", "flux_saturation",
    "rad"), from the device against the committed fixtures of the unmodified reference binary."""
    out = run_isolated(MODULE_PLANES_CODE.format(name=name), {})
    assert "ok" in out


ANOMRES_CODE = """
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    from test_oracle_vs_live_reference import AR_CASES, ar_kwargs
    from test_anomres_host_check import EXTRA
    name, kv, xb, yb, integ = [c for c in AR_CASES + EXTRA if c[0] == {name!r}][0]
    exact = "frobenius" not in name                    # the Frobenius template ends in pow(., 1.5): the libm tolerance class
    s = synthetic.stratified_loop(23, 23, bump=0.5)
    kw = dict(xb=xb, yb=yb, integrator=integ, density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    a = ar_kwargs(kv)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    o.set_anomalous_resistivity(**a)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d.set_anomalous_resistivity(s["planes"]["pos_x"], s["planes"]["pos_y"], **a)
    d.set_module_output_to_file("anomalous_resistivity")
    ref = o.run(4)
    dts = d.advance(4)
    (ri, rj), rt = o.anomalous_state()
    (di, dj), nsub = d.anomalous_resistivity_state()
    assert (di, dj) == (ri, rj), ((di, dj), (ri, rj))
    assert nsub == o.anomalous_subcycles(), (nsub, o.anomalous_subcycles())
    tm = d.module_output("anomalous_template")
    if exact:
        assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
        assert same_bits(tm, rt), "template: " + mismatch(tm, rt)
        for v in PlasmaDomain.EVOLVED + ["dt", "temp"]:
            assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    else:
        assert np.allclose(dts, ref, rtol=1e-9, atol=0.0), (dts, ref)
        assert np.max(np.abs(tm - rt)) <= 1e-9 * max(np.max(np.abs(rt)), 1e-300)
        for v in PlasmaDomain.EVOLVED + ["dt", "temp"]:
            r = o.get(v); g = d.grid(v)
            assert np.max(np.abs(g - r)) <= 1e-9 * max(np.max(np.abs(r)), 1e-300), v
    assert np.count_nonzero(d.module_output("joule_heating")) > 0
    assert d.module_output("anomalous_diffusivity").shape == tm.shape
    print("ok")
"""


@pytest.mark.parametrize("name", ["ar_default_floodfill", "ar_frobenius_rk2_gc", "ar_syntelis_rk4", "ar_ys94_radius", "ar_periodic_x_euler", "ar_moc_bounds_rk4", "ar_frobenius_plain"])
def test_anomalous_resistivity_vs_oracle(name):
    """anomalous_resistivity on the device (anomres_cells.hpp functors as kernels, anomres_host.cuh): whole steps with the module against the CPU
    restatement -- step sizes, every plane, the tracked null point, the template and the sub-cycle count.  The same functors in the same sequence are
    proven bit for bit on the host by tests/test_anomres_host_check.py; this is the launch side."""
    env = {}
    out = run_isolated(ANOMRES_CODE.format(name=name), env)
    assert "ok" in out


ANOMRES_GOLDEN_CODE = """
    import numpy as np
    from golden_util import Golden, same_bits, mismatch, OUT_VARS
    from spruce_b200.domain import PlasmaDomain
    g = Golden("ar_floodfill_rk2")
    d = PlasmaDomain(g.planes, g.ion_mass, g.adiabatic_index, **g.kw)
    for mname, kv in g.modules:
        d.set_anomalous_resistivity(g.planes["pos_x"], g.planes["pos_y"], **{k: float(v) for k, v in kv.items()})
    done = 0
    hist = []
    for it in sorted(g.frames):
        hist += list(d.advance(it - done)); done = it
        for v in OUT_VARS:
            assert same_bits(d.grid(v), g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(d.grid(v), g.frames[it][v]))
    assert [float(x).hex() for x in hist] == [float(x).hex() for x in g.steps[:done]]
    print("ok")
"""


def test_anomalous_resistivity_vs_reference_fixture():
    """the committed fixture of the unmodified reference binary with anomalous_resistivity (tests/golden/ar_floodfill_rk2.npz), bit for bit"""
    out = run_isolated(ANOMRES_GOLDEN_CODE, {})
    assert "ok" in out


ANOMRES_DIAG_CODE = """
    import numpy as np
    from golden_util import Golden, same_bits, mismatch, OUT_VARS
    from spruce_b200.domain import PlasmaDomain
    g = Golden("ar_diag_planes")
    d = PlasmaDomain(g.planes, g.ion_mass, g.adiabatic_index, **g.kw)
    for mname, kv in g.modules:
        kw = {k: float(v) for k, v in kv.items() if k != "output_to_file"}
        if mname == "anomalous_resistivity":
            d.set_anomalous_resistivity(g.planes["pos_x"], g.planes["pos_y"], **kw)
            d.set_module_output_to_file("anomalous_resistivity")
        else:
            d.set_field_heating(**kw)
    done = 0
    for it in sorted(g.frames):
        if it > done:
            d.advance(it - done); done = it
        for v in (OUT_VARS if it else []):
            r = g.frames[it][v]
            assert np.max(np.abs(d.grid(v) - r)) <= 1e-9 * max(np.max(np.abs(r)), 1e-300), "%s after iteration %d" % (v, it)      # field_heating calls pow
        mp = g.module_planes[it]
        assert same_bits(d.module_output("anomalous_template"), mp["anomalous_template"]) or it > 0, "set-up template (frame 0)"
        for name in ("anomalous_template", "anomalous_diffusivity", "joule_heating", "field_heating"):
            got, ref = d.module_output(name), mp[name]
            assert np.max(np.abs(got - ref)) <= 1e-6 * max(np.max(np.abs(ref)), 1e-300), "%s after iteration %d: %s" % (name, it, mismatch(got, ref))
            assert np.count_nonzero(got) == np.count_nonzero(ref) or name == "joule_heating", name
    print("ok")
"""


def test_anomalous_resistivity_and_field_heating_diagnostic_planes_vs_reference_fixture():
    """output_to_file planes of anomalous_resistivity and field_heating against the fixture of the unmodified reference (tests/golden/ar_diag_planes.npz):
    frame 0 carries the set-up template and zero planes, later frames the last evaluation's template, template*diffusivity, (e_after - e_before)/dt and mask*(dt*heating)."""
    out = run_isolated(ANOMRES_DIAG_CODE, {})
    assert "ok" in out


MULTISPECIES_AR_CODE = """
    import numpy as np
    from golden_util import Golden, MS_PLANES, OUT_VARS, same_bits, mismatch, multispecies_fractions, small_module_kwargs, sink_reduction_plane
    from spruce_b200.domain import PlasmaDomain
    g = Golden("ar_ms_joule_sources_rk2")
    d = PlasmaDomain(g.planes, g.ion_mass, g.adiabatic_index, **g.kw)
    for mname, kv in g.modules:
        if mname == "anomalous_resistivity":
            d.set_anomalous_resistivity(g.planes["pos_x"], g.planes["pos_y"], **{k: float(v) for k, v in kv.items()})
        elif mname == "ambient_heating_sink":
            d.set_ambient_heating_sink_plane(sink_reduction_plane(g.planes, small_module_kwargs(mname, kv)[1], g.kw["xb"], g.kw["yb"]))
        else:
            getattr(d, "set_" + mname)(**small_module_kwargs(mname, kv)[1])
    d.set_multispecies(True, **multispecies_fractions(g.modules))
    hist = []
    for it in range(1, g.n_steps + 1):
        hist += list(d.advance(1))
        if it in g.frames:
            for v in OUT_VARS:
                assert same_bits(d.grid(v), g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(d.grid(v), g.frames[it][v]))
            for name in MS_PLANES:
                got, ref = d.module_output(name), g.module_planes[it][name]
                assert np.count_nonzero(ref) > 0 and same_bits(got, ref), "%s after iteration %d: %s" % (name, it, mismatch(got, ref))
        d.multispecies_reset()                  # the fixture stores every iteration (evolution.cpp:36-41)
    assert [float(x).hex() for x in hist] == [float(x).hex() for x in g.steps[:g.n_steps]]
    print("ok")
"""


@pytest.mark.xfail(reason="written after round 2's GPU budget was spent (CPU-checked: tests/test_capi_hooks_emulated.py); first executed by the round-end run", strict=False)
def test_multispecies_joule_and_source_planes_vs_reference_fixture():
    """multispecies_mode on the device: the joule heating of anomalous_resistivity and the electron / ion shares of ambient_heating_sink and localized_heating against the fixture
    of the unmodified reference (tests/golden/ar_ms_joule_sources_rk2.npz), bit for bit"""
    out = run_isolated(MULTISPECIES_AR_CODE, {})
    assert "ok" in out


OPERATOR2_CODE = """
    import numpy as np
    from golden_util import same_bits, mismatch
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    for xb, yb, gen in [(("periodic", "periodic"), ("periodic", "periodic"), lambda: synthetic.orszag_tang(70, 51, zfull=True)),
                        (("fixed", "open"), ("reflect", "fixed"), lambda: synthetic.stratified_loop(45, 66))]:
        s = gen()
        kw = dict(xb=xb, yb=yb)
        o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
        q, ax, ay = o.get("temp"), o.get("v_x") - 0.3 * o.get("v_y"), o.get("b_y") + 2.0 * o.get("v_y")
        # derivs.cpp:407-409, 472-474, 216-220: the oracle's single-direction operators combined with the reference's Grid + / -
        ref = {"divergence2D": o.operator("derivative1D", 0, ax) + o.operator("derivative1D", 1, ay),
               "curl2D": o.operator("derivative1D", 0, ay) - o.operator("derivative1D", 1, ax),
               "transportDivergence2D": o.operator("transportDerivative1D", 0, q, ax) + o.operator("transportDerivative1D", 1, q, ay)}
        assert same_bits(d.operator2("divergence2D", ax, ay), ref["divergence2D"])
        assert same_bits(d.operator2("curl2D", ax, ay), ref["curl2D"])
        assert same_bits(d.operator2("transportDivergence2D", q, ax, ay), ref["transportDivergence2D"])
    print("ok")
"""


@pytest.mark.xfail(reason="written after round 2's GPU budget was spent (two validated spruce_operator passes combined on the host); first executed by the round-end run", strict=False)
def test_two_operand_operators_vs_oracle():
    """spruce_operator2: divergence2D, curl2D, transportDivergence2D of PlasmaDomain (plasmadomain.hpp:201-242), bit for bit."""
    out = run_isolated(OPERATOR2_CODE, {})
    assert "ok" in out
