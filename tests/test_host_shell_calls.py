"""The host shell (spruce_b200/bin/run: the C++ PlasmaDomain / EquationSet / ModuleHandler / Module mirror) run WITHOUT a GPU over a recording
stand-in for the device library (tests/hostcheck/capi_stub.c, preloaded): a .config with every solar module must become the C-ABI calls the
reference's classes imply -- ModuleHandler order, parameter meaning, the reference's defaults, enum codes, diagnostic planes appended to mhd.out --
and the run loop must honour max_iterations / iter_output_interval.  The device side of each call is covered by the GPU parity tests."""
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import refrun
from spruce_b200 import synthetic

ROOT = Path(__file__).resolve().parents[1]
OURS = ROOT / "spruce_b200" / "bin" / "run"
STUB_SRC = ROOT / "tests" / "hostcheck" / "capi_stub.c"
STUB = ROOT / "tests" / "hostcheck" / "_build" / "libcapi_stub.so"


@pytest.fixture(scope="module")
def stub():
    STUB.parent.mkdir(exist_ok=True)
    hdr = ROOT / "include" / "spruce_b200.h"
    if not STUB.exists() or STUB.stat().st_mtime < max(STUB_SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["gcc", "-std=gnu11", "-O1", "-Wall", "-Werror", "-shared", "-fPIC", "-I", str(ROOT / "include"), str(STUB_SRC), "-o", str(STUB)], check=True)
    subprocess.run(["make", "-C", str(ROOT / "spruce_b200" / "host")], check=True, stdout=subprocess.DEVNULL)
    return STUB


def run_shell(stub, tmp_path, s, cfg):
    state = tmp_path / "in.state"
    refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"])
    out = tmp_path / "out"
    out.mkdir()
    (out / "run.config").write_text(cfg)
    log = tmp_path / "calls.log"
    env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(log))
    r = subprocess.run([str(OURS), "-m", "input", "-o", str(out), "-s", str(state)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode in (-6, 134), r.stderr.decode()[-2000:]
    assert "Simulation successfully reached max simulation time or iterations" in r.stderr.decode(), r.stderr.decode()[-2000:]      # not a spruce_die
    return log.read_text().splitlines(), r.stdout.decode(), out


def args_of(line):
    return dict(re.findall(r"(\w+)=(\S+)", line))


def test_every_solar_module_reaches_the_c_abi_in_config_order(stub, tmp_path):
    nx, ny = 20, 18
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    modules = [
        ("thermal_conduction", [("flux_saturation", "true"), ("epsilon", "0.1"), ("dt_subcycle_min", "1.0e-4"), ("time_integrator", "rk4"), ("output_to_file", "true")]),
        ("radiative_losses", [("cutoff_ramp", "1.0e3"), ("cutoff_temp", "3.0e4"), ("epsilon", "0.1"), ("output_to_file", "true")]),
        ("ambient_heating", [("heating_rate", "1.0e-4")]),
        ("ambient_heating_sink", [("heating_rate", "2.0e-4")]),
        ("localized_heating", [("start_time", "1.0"), ("duration", "10.0"), ("max_heating_rate", "0.5"), ("stddev_x", "3.0"), ("stddev_y", "2.0"), ("center_x", "8.0"), ("center_y", "7.0")]),
        ("mass_injection", [("start_time", "2.0"), ("duration", "5.0"), ("max_injection_rate", "1.0e6"), ("stddev_x", "2.0"), ("stddev_y", "2.5"), ("center_x", "9.0"), ("center_y", "6.0")]),
        ("momentum_injection", [("start_time", "0.0"), ("duration", "50.0"), ("max_accel", "1.0e4"), ("stddev_x", "2.0"), ("stddev_y", "2.0"), ("center_x", "10.0"), ("center_y", "9.0"),
                                ("dir_x", "0.0"), ("dir_y", "1.0"), ("oscillatory", "true"), ("oscillation_period", "7.0")]),
        ("div_cleaning", [("epsilon", "0.05"), ("time_scale", "4.0")]),
        ("field_heating", [("coeff", "1.0e-7"), ("current_pow", "0.5"), ("b_pow", "1.0"), ("n_pow", "0.2"), ("roc_pow", "0.3"), ("output_to_file", "true")]),
        ("boundary_outflow", [("max_accel", "3.0e4"), ("falloff_length", "5.0e8"), ("boundary", "x_bound_2"), ("falloff_shape", "gaussian"), ("feather_length", "1.0e8"),
                              ("dynamic_mode", "true"), ("dynamic_time", "20.0"), ("dynamic_target_speed", "1.0e6")]),
        ("anomalous_resistivity", [("time_scale", "0.3"), ("output_to_file", "true"), ("resistivity_model", "ys_94"), ("resistivity_model_params", "0.2,2.0e14,3.0e15"),
                                   ("time_integrator", "rk4"), ("gradient_correction", "true"), ("flood_fill_threshold", "2.5"), ("metric_smoothing", "false")]),
        ("physical_viscosity", [("coeff", "1.0e-14"), ("epsilon", "0.1"), ("time_integrator", "rk2"), ("gradient_correction", "true"), ("output_to_file", "true")]),
    ]
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="rk4", xb=("fixed", "open"), yb=("reflect", "open"), max_iterations=3, iter_output_interval=2, modules=modules)
    log, stdout, out = run_shell(stub, tmp_path, s, cfg)
    calls = [ln.split()[0] for ln in log if ln.startswith("spruce_")]

    create = args_of(log[0])
    assert log[0].startswith("spruce_domain_create") and create["xdim"] == "20" and create["ydim"] == "18" and create["eqs"] == "0"
    assert create["bc"] == "2,1,3,1" and create["ti"] == "2"                       # fixed, open, reflect, open; rk4
    assert float(create["epsilon"]) == 0.2 and create["n_ranks"] == "1"

    # state variables are uploaded before the set-up, modules are configured after it, in config order (modulehandler.cpp:27-65)
    uploads = [ln.split()[1] for ln in log if ln.startswith("spruce_grid_upload")]
    assert uploads[:3] == ["be_x", "be_y", "be_z"] and set(uploads[3:]) == {"rho", "temp", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "grav_x", "grav_y"}
    i_setup = calls.index("spruce_eqs_setup")
    assert all(c != "spruce_grid_upload" for c in calls[i_setup:])
    module_calls = [c for c in calls if c.startswith("spruce_module_") and c not in ("spruce_module_output", "spruce_module_output_to_file", "spruce_module_viscosity_term")]
    assert module_calls == ["spruce_module_thermal_conduction", "spruce_module_radiative_losses", "spruce_module_ambient_heating", "spruce_module_ambient_heating_sink",
                            "spruce_module_localized_heating", "spruce_module_mass_injection", "spruce_module_momentum_injection", "spruce_module_div_cleaning",
                            "spruce_module_field_heating", "spruce_module_boundary_outflow", "spruce_module_anomalous_resistivity", "spruce_module_physical_viscosity"]
    assert calls.index("spruce_module_thermal_conduction") > i_setup

    line = {c: args_of(next(ln for ln in log if ln.startswith(c + " "))) for c in module_calls if c not in ("spruce_module_ambient_heating", "spruce_module_ambient_heating_sink")}
    tc = line["spruce_module_thermal_conduction"]
    assert tc["flux_saturation"] == "1" and tc["ti"] == "2" and float(tc["epsilon"]) == 0.1 and float(tc["dt_subcycle_min"]) == 1.0e-4
    rl = line["spruce_module_radiative_losses"]
    assert float(rl["cutoff_ramp"]) == 1.0e3 and float(rl["cutoff_temp"]) == 3.0e4 and float(rl["epsilon"]) == 0.1
    lh = line["spruce_module_localized_heating"]
    assert float(lh["start"]) == 1.0 and float(lh["duration"]) == 10.0 and float(lh["max"]) == 0.5 and lh["stddev"] == "3,2" and lh["center"] == "8,7"
    mi = line["spruce_module_mass_injection"]
    assert float(mi["start"]) == 2.0 and float(mi["max"]) == 1.0e6 and mi["center"] == "9,6"
    mo = line["spruce_module_momentum_injection"]
    assert float(mo["max_accel"]) == 1.0e4 and mo["dir"] == "0,1" and mo["oscillatory"] == "1" and float(mo["period"]) == 7.0
    dc = line["spruce_module_div_cleaning"]
    assert float(dc["epsilon"]) == 0.05 and float(dc["time_scale"]) == 4.0
    fh = line["spruce_module_field_heating"]
    assert [float(fh[k]) for k in ("coeff", "current_pow", "b_pow", "n_pow", "roc_pow")] == [1.0e-7, 0.5, 1.0, 0.2, 0.3] and fh["inactive"] == "0"
    bo = line["spruce_module_boundary_outflow"]
    assert float(bo["max_accel"]) == 3.0e4 and bo["boundary"] == "1" and bo["shape"] == "1" and float(bo["feather"]) == 1.0e8 and bo["dynamic"] == "1" and bo["field_aligned"] == "0"
    assert float(bo["dynamic_time"]) == 20.0 and float(bo["target"]) == 1.0e6
    pv = line["spruce_module_physical_viscosity"]
    assert float(pv["coeff"]) == 1.0e-14 and pv["ti"] == "1" and pv["gradient_correction"] == "1" and pv["heating_on"] == "1" and pv["force_on"] == "1"

    # anomalous_resistivity: the 17 numbers in the header's order, the reference's defaults where the config is silent (anomalousresistivity.hpp:16-39)
    i_ar = next(i for i, ln in enumerate(log) if ln.startswith("spruce_module_anomalous_resistivity "))
    p = [float(log[i_ar + 1 + k].split("=")[1]) for k in range(17)]
    assert p == [0.3, 1.0e50, 3.0, 1.0, 0.0, 2.0, 1.0, -1.0, 5.0e9, -1.0, 1.0e-5, 2.5, 2.0, 1.0, 0.2, 2.0e14, 3.0e15]
    pos = args_of(log[i_ar + 18])
    assert pos["count"] == str(nx * ny) and float(pos["sum"]) == pytest.approx(float(np.sum(s["planes"]["pos_x"])), rel=1e-12)

    # planes handed over with a module: the sink's reduction carries the ghost-zone mask, the plain heating plane does not have to
    i_sink = next(i for i, ln in enumerate(log) if ln.startswith("spruce_module_ambient_heating_sink"))
    assert args_of(log[i_sink + 1])["count"] == str(nx * ny)

    # output_to_file: enabled per module, every frame (iterations 0, 2 and the final one) appends the diagnostic planes in module order
    enabled = [ln.split()[1] for ln in log if ln.startswith("spruce_module_output_to_file")]
    assert enabled == ["thermal_conduction", "radiative_losses", "anomalous_resistivity", "physical_viscosity"]
    outs = [ln.split()[1] for ln in log if ln.startswith("spruce_module_output ")]
    frame = ["thermal_conduction", "flux_saturation", "rad", "field_heating", "anomalous_diffusivity", "anomalous_template", "joule_heating",
             "viscous_heating", "viscous_force_x", "viscous_force_y", "viscous_force_z"]                                         # physicalviscosity.cpp:292-308
    assert outs[:len(frame)] == frame and len(outs) % len(frame) == 0 and len(outs) // len(frame) >= 2
    _, frames = refrun.read_out(out / "mhd.out")
    for f in frames:
        for name in frame:
            assert name in f and np.all(f[name] == 7.0), name

    # run loop: three iterations, advance() batched up to the next output / message point
    assert sum(int(args_of(ln)["done"]) for ln in log if ln.startswith("spruce_advance")) == 3
    for msg in ("Thermal Subcycles: 9", "Radiative Subcycles: 9", "Anomalous Resistivity Subcycles: 9", "Field Heating On", "x_bound_2 boundary outflow enforced"):
        assert msg in stdout, msg
    assert (out / "end.state").exists() and (out / "init.state").exists()


def test_moc_options_and_unknown_module(stub, tmp_path):
    s = synthetic.stratified_loop(16, 14)
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="euler", xb=("open_moc", "fixed"), yb=("fixed", "open_moc"), max_iterations=2, iter_output_interval=1,
                                  eqs_block=[("global_viscosity", "0.25"), ("moc_b_limiting", "true"), ("moc_b_lower_lim", "0.2"), ("moc_mom_limiting", "true"), ("moc_mom_upper_lim", "5.0")])
    log, _, _ = run_shell(stub, tmp_path, s, cfg)
    assert args_of(log[0])["bc"] == "4,2,2,4" and args_of(log[0])["ti"] == "0"
    gv = args_of(next(ln for ln in log if ln.startswith("spruce_eqs_ideal_mhd_options")))
    assert float(gv["global_viscosity"]) == 0.25
    lim = next(ln for ln in log if ln.startswith("spruce_eqs_ideal_mhd_moc_limiting")).split()
    assert lim[1:] == ["b=1", "0.20000000000000001", "10", "mom=1", "0.10000000000000001", "5"]              # silent bounds keep idealmhd.hpp:59-64's defaults
    calls = [ln.split()[0] for ln in log if ln.startswith("spruce_")]
    assert calls.index("spruce_eqs_ideal_mhd_moc_limiting") < calls.index("spruce_eqs_setup")


def test_continue_mode_resumes_from_end_state_with_its_time(stub, tmp_path):
    """mhdSolve(prev_run_directory, ...) (mhd.cpp:6-19): end.state + the stored config, m_time taken from the file (fileio.cpp:50-51), run for -d more seconds."""
    s = synthetic.stratified_loop(16, 14)
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="rk2", xb=("fixed", "open"), yb=("fixed", "open"), max_iterations=-1, iter_output_interval=2, duration=1.75)
    log1, _, out = run_shell(stub, tmp_path, s, cfg)
    adv = [args_of(ln) for ln in log1 if ln.startswith("spruce_advance")]
    assert sum(int(a["done"]) for a in adv) == 4 and all(float(a["max_time"]) == 1.75 for a in adv)        # 0.5, 0.5, 0.5, then the 0.25 that reaches the duration
    meta, _ = refrun.read_state(out / "end.state")
    assert float(meta["t"]) == 1.75
    log2 = tmp_path / "calls2.log"
    env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(log2))
    r = subprocess.run([str(OURS), "-m", "continue", "-o", str(out), "-d", "1.0"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode in (-6, 134) and "Simulation successfully reached" in r.stderr.decode(), r.stderr.decode()[-2000:]
    lines = log2.read_text().splitlines()
    assert float(args_of(lines[0])["time"]) == 1.75
    adv = [args_of(ln) for ln in lines if ln.startswith("spruce_advance")]
    assert all(float(a["max_time"]) == 2.75 for a in adv) and sum(int(a["done"]) for a in adv) == 2
    meta, _ = refrun.read_state(out / "end.state")
    assert float(meta["t"]) == 2.75


def test_two_fluid_and_two_energy_sets_select_their_equation_set(stub, tmp_path):
    s = synthetic.ucnp_cloud(16, 14, drift=2.0e3, bfield=5.0)
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="rk4", xb=("open_ucnp", "open_ucnp"), yb=("periodic", "periodic"), max_iterations=2, iter_output_interval=1, eqs="ideal_2F",
                                  eqs_block=[("use_sub_cycling", "false")], density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1e-30, output_flags=("i_rho", "e_rho", "E_x", "dt"),
                                  modules=[("eic_thermalization", [])])
    (tmp_path / "a").mkdir()
    log, _, _ = run_shell(stub, tmp_path / "a", s, cfg)
    c = args_of(log[0])
    assert c["eqs"] == "3" and c["bc"] == "5,5,0,0" and c["ti"] == "2"          # index in EquationSet::m_sets (equationset.hpp:22)
    opt = args_of(next(ln for ln in log if ln.startswith("spruce_eqs_ideal2f_options")))
    assert opt["use_sub_cycling"] == "0" and opt["remove_curl_terms"] == "0"
    calls = [ln.split()[0] for ln in log if ln.startswith("spruce_")]
    assert "spruce_module_eic_thermalization" in calls and calls.index("spruce_module_eic_thermalization") > calls.index("spruce_eqs_setup")
    uploads = {ln.split()[1] for ln in log if ln.startswith("spruce_grid_upload")}
    assert {"i_rho", "e_rho", "i_mom_x", "e_mom_x", "i_temp", "e_temp", "E_x", "E_y", "E_z", "bi_x", "bi_y", "bi_z"} <= uploads

    s = synthetic.two_energy(16, 14)
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="euler", xb=("reflect", "open"), yb=("fixed", "open"), max_iterations=2, iter_output_interval=1, eqs="ideal_mhd_2E",
                                  output_flags=("rho", "i_temp", "e_temp", "dt"), modules=[("eic_thermalization", [])])
    (tmp_path / "b").mkdir()
    log, _, _ = run_shell(stub, tmp_path / "b", s, cfg)
    c = args_of(log[0])
    assert c["eqs"] == "2" and c["bc"] == "3,1,2,1" and c["ti"] == "0"
    calls = [ln.split()[0] for ln in log if ln.startswith("spruce_")]          # the UCNP configuration: ideal_mhd_2E has every grid eic_thermalization looks up
    assert "spruce_module_eic_thermalization" in calls and calls.index("spruce_module_eic_thermalization") > calls.index("spruce_eqs_setup")
    uploads = {ln.split()[1] for ln in log if ln.startswith("spruce_grid_upload")}
    assert {"rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "grav_x", "grav_y"} <= uploads


def test_host_resident_module_hooks_run_between_single_device_steps(stub, tmp_path):
    """sg_filtering stays on the host (host/module.cpp: SGFilter): the run loop advances ONE step per device call and the post-iterate hook stages
    rho and thermal_energy through the C ABI (download, filter, upload) and propagates -- after the steps whose index is a non-zero multiple of
    filter_interval (the hook sees m_iter before its increment, evolution.cpp:74-81), next to a device-resident module in the same config."""
    s = synthetic.stratified_loop(20, 18)
    modules = [("ambient_heating", [("heating_rate", "1.0e-4")]), ("sg_filtering", [("filter_interval", "2")])]
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "open"), max_iterations=5, iter_output_interval=5, modules=modules)
    log, stdout, _ = run_shell(stub, tmp_path, s, cfg)
    assert "SG Filtering On" in stdout                                        # commandLineMessage, sgfilter.cpp:25-28
    calls = [ln.split()[0] + (" " + ln.split()[1] if ln.startswith("spruce_grid_upload") else "") for ln in log if ln.startswith("spruce_")]
    run = calls[calls.index("spruce_advance"):]
    advances = [ln for ln in log if ln.startswith("spruce_advance")]
    assert len(advances) == 5 and all(args_of(ln)["n"] == "1" for ln in advances)
    hook = ["spruce_grid_upload rho", "spruce_grid_upload thermal_energy", "spruce_eqs_propagate_changes"]
    expected = []
    for it in range(5):                                                       # m_iter of the step just integrated
        expected.append("spruce_advance")
        if it != 0 and it % 2 == 0:
            expected += hook
    assert [c for c in run if c == "spruce_advance" or c in hook] == expected


def test_ucnp_host_modules_stage_their_planes_after_every_step(stub, tmp_path):
    """coulomb_explosion and global_temperature stay on the host (host/module.cpp, host/ucnp_modules.hpp): after every single device step, in config order,
    global_temperature runs int(gt_strength) midpoint sub-steps (two laplacian operator passes each), uploads thermal_energy and propagates;
    coulomb_explosion differentiates the pressure for its output planes, uploads mom_x / mom_y and propagates -- until 3 * timescale has passed."""
    s = synthetic.ucnp_cloud_mhd(83, 81)
    modules = [("global_temperature", [("gt_species", "i"), ("gt_strength", "2.5"), ("gt_use_diffusion", "true")]),
               ("coulomb_explosion", [("timescale", "0.4"), ("lengthscale", "0.2"), ("strength", "1.0e-3"), ("output_to_file", "true")])]
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="euler", xb=("open_ucnp", "open_ucnp"), yb=("open_ucnp", "open_ucnp"), max_iterations=4, iter_output_interval=4, modules=modules,
                                  density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1.0e-30)
    log, stdout, out = run_shell(stub, tmp_path, s, cfg)
    assert "Coulomb Explosion On" in stdout
    calls = []
    for ln in log:
        w = ln.split()
        if w[0] in ("spruce_grid_upload", "spruce_operator"):
            calls.append(" ".join(w[:3] if w[0] == "spruce_operator" else w[:2]))
        elif w[0] in ("spruce_advance", "spruce_eqs_propagate_changes"):
            calls.append(w[0])
    run = calls[calls.index("spruce_advance"):]
    gt = ["spruce_operator laplacian 0"] * 4 + ["spruce_grid_upload thermal_energy", "spruce_eqs_propagate_changes"]
    ce = ["spruce_operator derivative1D 0", "spruce_operator derivative1D 1", "spruce_grid_upload mom_x", "spruce_grid_upload mom_y", "spruce_eqs_propagate_changes"]
    # the stand-in's steps are 0.5 each: the hook sees t = 0, 0.5, 1.0, 1.5 and 3 * timescale = 1.2
    expected = []
    for it in range(4):
        expected += ["spruce_advance"] + gt + (ce if 0.5 * it < 1.2 else [])
    assert run == expected
    frames = (out / "mhd.out").read_text()
    assert all(("\n%s\n" % v) in frames for v in ("F_x", "F_y", "dP_x", "dP_y"))            # fileOutput, coulomb_explosion.cpp:89-97


def test_unported_names_are_refused(stub, tmp_path):
    s = synthetic.stratified_loop(16, 14)
    for block, msg in (([("no_such_module", [])], "no_such_module"), ([("artificial_viscosity", [("visc_opt", "boundary"), ("visc_strength", "0.5"), ("visc_vars_to_diff", "v_x"), ("visc_vars_to_evol", "mom_x"),
                                                                                                  ("visc_length", "1.0e8"), ("visc_species", "i"), ("boundary_falloff_shape", "hexagonal")])], "falloff shape")):
        cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="euler", xb=("fixed", "fixed"), yb=("fixed", "fixed"), max_iterations=1, iter_output_interval=1, modules=block)
        state = tmp_path / ("in_%s.state" % msg.replace(" ", "_"))
        refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"])
        out = tmp_path / ("out_" + msg.replace(" ", "_"))
        out.mkdir()
        (out / "run.config").write_text(cfg)
        env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(tmp_path / "calls.log"))
        r = subprocess.run([str(OURS), "-m", "input", "-o", str(out), "-s", str(state)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
        err = r.stderr.decode()
        assert r.returncode in (-6, 134) and msg in err and "successfully reached" not in err, err[-1500:]
    state = tmp_path / "in_ms.state"
    refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"])
    out = tmp_path / "out_ms"
    out.mkdir()
    env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(tmp_path / "calls.log"))
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="euler", xb=("fixed", "fixed"), yb=("fixed", "fixed"), max_iterations=1, iter_output_interval=1,
                                  modules=[("global_temperature", [("gt_species", "i"), ("gt_strength", "2.0"), ("gt_use_global_temp", "true")])])
    (out / "run.config").write_text(cfg)
    r = subprocess.run([str(OURS), "-m", "input", "-o", str(out), "-s", str(state)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode in (-6, 134) and "gt_use_global_temp" in r.stderr.decode() and "successfully reached" not in r.stderr.decode()


def test_artificial_viscosity_output_planes_are_requested_per_term_and_named_like_the_reference(stub, tmp_path):
    """visc_output_visc / _lap / _strength / _timescale (viscosity.cpp:9-12, 351-376): enabled after the terms are configured; every frame appends, per flag, one plane per
    term in config order under the reference's names <evolved>_dqdt / _lap / _str / _dt"""
    s = synthetic.stratified_loop(20, 18, bump=0.5)
    modules = [("artificial_viscosity", [("visc_opt", "boundary,global"), ("visc_strength", "0.8,3.0"), ("visc_vars_to_diff", "v_x,temp"), ("visc_vars_to_evol", "mom_x,thermal_energy"),
                                         ("visc_length", "5.0e8,0"), ("visc_species", "i,i"), ("hv_time_integrator", "rk2"), ("visc_output_visc", "true"), ("visc_output_lap", "false"),
                                         ("visc_output_strength", "true"), ("visc_output_timescale", "true")])]
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="rk2", xb=("fixed", "open"), yb=("reflect", "open"), max_iterations=2, iter_output_interval=1, modules=modules)
    log, stdout, out = run_shell(stub, tmp_path, s, cfg)
    calls = [ln.split()[0] for ln in log if ln.startswith("spruce_")]
    i_en = next(i for i, ln in enumerate(log) if ln.startswith("spruce_module_output_to_file artificial_viscosity 1"))
    assert all(i < i_en for i, ln in enumerate(log) if ln.startswith("spruce_module_viscosity_term"))
    assert calls.count("spruce_module_viscosity_term") == 2
    outs = [ln.split()[1] for ln in log if ln.startswith("spruce_module_output ")]
    frame = ["visc_dqdt:0", "visc_dqdt:1", "visc_str:0", "visc_str:1", "visc_dt:0", "visc_dt:1"]
    assert outs[:len(frame)] == frame and len(outs) == 3 * len(frame)
    names, frames = refrun.read_out(out / "mhd.out")
    for f in frames:
        got = [k for k in f if k.endswith(("_dqdt", "_lap", "_str", "_dt")) and k != "dt"]
        assert got == ["mom_x_dqdt", "thermal_energy_dqdt", "mom_x_str", "thermal_energy_str", "mom_x_dt", "thermal_energy_dt"], got


def test_a_set_written_by_the_problem_generator_runs_through_the_shell(stub, tmp_path):
    """spruce_b200/bin/gengrids (the reference's execs/gengrids.cpp; tests/test_host_gengrids.py holds its files to the reference generator's byte for byte) -> `run`: the
    generated ucnp.config / init.state of a set are the inputs of the drop-in binary as they stand -- ideal_mhd_2E with eic_thermalization on the generator's odd, non-uniform
    grid, the step count taken from a settings row named like a config key"""
    from test_host_gengrids import BASE, CASES, CONFIG
    settings, config, _ = CASES["sweep_2e_nonuniform_runs"]
    (tmp_path / "sweep.settings").write_text(BASE.format(**dict(settings, n="1e9", n_dist="gaussian", extra="max_iterations = cgs = 2\nstd_out_interval = cgs = 1\n")))
    (tmp_path / "template.config").write_text(CONFIG.format(**config))
    gen = ROOT / "spruce_b200" / "bin" / "gengrids"
    r = subprocess.run([str(gen), "-p", str(tmp_path / "sets"), "-s", str(tmp_path / "sweep.settings"), "-c", str(tmp_path / "template.config")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    set_dir = tmp_path / "sets" / "set_0"
    out = tmp_path / "out"
    out.mkdir()
    (out / "run.config").write_text((set_dir / "ucnp.config").read_text())
    log = tmp_path / "calls.log"
    env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(log))
    r = subprocess.run([str(OURS), "-m", "input", "-o", str(out), "-s", str(set_dir / "init.state")], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode in (-6, 134) and "Simulation successfully reached max simulation time or iterations" in r.stderr.decode(), r.stderr.decode()[-2000:]
    lines = log.read_text().splitlines()
    c = args_of(lines[0])
    assert c["eqs"] == "2" and c["xdim"] == "21" and c["ydim"] == "17" and c["bc"] == "5,5,5,5" and float(c["epsilon"]) == 0.15      # epsilon: the sweep's row replaced the template's line
    calls = [ln.split()[0] for ln in lines if ln.startswith("spruce_")]
    assert "spruce_module_eic_thermalization" in calls
    assert 1 <= sum(int(args_of(ln)["done"]) for ln in lines if ln.startswith("spruce_advance")) <= 2       # max_iterations = 2 from the sweep, or its duration (0.3 tau) first
    uploads = {ln.split()[1] for ln in lines if ln.startswith("spruce_grid_upload")}
    assert {"be_x", "be_y", "rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y"} <= uploads


def test_multispecies_mode_enables_stores_and_resets_the_cumulative_planes(stub, tmp_path):
    """multispecies_mode = true (plasmadomain.hpp:134-135): enabled on the device before the modules are set up, each module's ms_electron_heating_fraction handed over when its
    block sets one (and range-checked like the reference's asserts), the three cumulative planes appended to every stored frame between the equation set's variables and the
    modules' planes (fileio.cpp:164-183), reset after every store inside the time loop (evolution.cpp:36-41) -- not after the frame the constructor writes"""
    s = synthetic.stratified_loop(20, 18, bump=0.5)
    modules = [("thermal_conduction", [("flux_saturation", "true"), ("epsilon", "0.1"), ("dt_subcycle_min", "1.0e-4"), ("ms_electron_heating_fraction", "0.7"), ("output_to_file", "true")]),
               ("radiative_losses", [("cutoff_ramp", "1.0e3"), ("cutoff_temp", "3.0e4"), ("epsilon", "0.1"), ("inactive_mode", "true")]),
               ("ambient_heating", [("heating_rate", "1.0e-4"), ("ms_electron_heating_fraction", "0.3")]),
               ("localized_heating", [("start_time", "0.0"), ("duration", "10.0"), ("max_heating_rate", "0.5"), ("stddev_x", "3.0"), ("stddev_y", "2.0"), ("center_x", "8.0"), ("center_y", "7.0"),
                                      ("ms_electron_heating_fraction", "0.2")])]
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), max_iterations=4, iter_output_interval=2, modules=modules, multispecies=True)
    log, stdout, out = run_shell(stub, tmp_path, s, cfg)
    calls = [ln.split()[0] for ln in log if ln.startswith("spruce_")]
    assert calls.index("spruce_eqs_setup") < calls.index("spruce_multispecies_mode") < calls.index("spruce_module_thermal_conduction")
    fr = {ln.split()[1]: float(args_of(ln)["fraction"]) for ln in log if ln.startswith("spruce_module_ms_fraction")}
    assert fr == {"thermal_conduction": 0.7, "ambient_heating": 0.3, "localized_heating": 0.2}                # radiative_losses keeps the library's default
    assert calls.count("spruce_multispecies_reset") == 2                                                         # iterations 2 and 4
    assert "spruce_module_inactive_mode radiative_losses 1" in log and "Radiative Subcycles: 9 (Not Applied)" in stdout and "Thermal Subcycles: 9|" in stdout + "|"
    first_store = calls.index("spruce_module_output")
    assert first_store < calls.index("spruce_multispecies_reset") and "spruce_advance" not in calls[:first_store]
    _, frames = refrun.read_out(out / "mhd.out")
    assert len(frames) == 3
    for f in frames:
        keys = list(f)
        assert keys.index("cumulative_electron_heating") + 1 == keys.index("cumulative_ion_heating") and keys.index("cumulative_ion_heating") + 1 == keys.index("cumulative_joule_heating")
        assert keys.index("cumulative_joule_heating") < keys.index("thermal_conduction")
    bad = cfg.replace("ms_electron_heating_fraction = 0.3", "ms_electron_heating_fraction = 1.5")
    (tmp_path / "b").mkdir()
    state = tmp_path / "b" / "in.state"
    refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"])
    (tmp_path / "b" / "out").mkdir()
    (tmp_path / "b" / "out" / "run.config").write_text(bad)
    env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(tmp_path / "b" / "calls.log"))
    r = subprocess.run([str(OURS), "-m", "input", "-o", str(tmp_path / "b" / "out"), "-s", str(state)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode in (-6, 134) and "Ambient Heating MS electron heating fraction must be between 0 and 1" in r.stderr.decode()
