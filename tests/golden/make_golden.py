#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference binary
(oracle/_ref/run, built from /root/reference by oracle/Makefile) on small deterministic inputs.

    python tests/golden/make_golden.py [case ...]

Each fixture <case>.npz holds: the exact input planes, the case description (JSON: boundary
conditions, integrator, floors, module blocks), the per-iteration step sizes recovered from the
lossless (write_precision = 17) `dt` output planes  -- step_k = epsilon * min(dt_k over the interior),
reference source/mhd/evolution.cpp:62 --, and the evolved planes + dt + temp after selected iterations.

Needs /root/reference only through the prebuilt binary; the fixtures themselves travel with the repo.
"""
from __future__ import annotations

import json
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import refrun  # noqa: E402
from spruce_b200 import synthetic  # noqa: E402

HERE = Path(__file__).resolve().parent
OUT_VARS = ["rho", "temp", "thermal_energy", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "dt"]
OUT_VARS_2E = ["rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "i_thermal_energy", "e_thermal_energy", "press", "n", "v_x", "kinetic_energy", "b_mag", "b_hat_y", "dt"]
OUT_VARS_2F = ["i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_thermal_energy", "e_thermal_energy", "E_x", "E_y", "E_z",
               "bi_x", "bi_y", "bi_z", "i_temp", "e_temp", "dt", "dt_i", "j_x", "rho_c", "divE", "divB", "curlE_z", "e_dPdx", "b_hat_x"]
NX, NY = 32, 28

INACTIVE_FLOORS = dict(density_min=1.0, temp_min=1.0, thermal_energy_min=1.0e-30)
SOLAR_FLOORS = dict(density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)

UCNP_FLOORS = dict(density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1.0e-30)
TF = dict(eqs="ideal_2F", eqs_block=[("use_sub_cycling", "false")])
EIC = ("eic_thermalization", [])
UC = ("open_ucnp", "open_ucnp")
PP = ("periodic", "periodic")
CLOUD = dict(nx=NX, ny=NY, drift=2.0e3, bfield=5.0)

TC = lambda **kw: ("thermal_conduction", list(dict(dict(flux_saturation="false", epsilon="0.1", dt_subcycle_min="1.0e-4", output_to_file="true"), **kw).items()))
RL = lambda **kw: ("radiative_losses", list(dict(dict(cutoff_ramp="1.0e3", cutoff_temp="3.0e4", epsilon="0.1", output_to_file="true"), **kw).items()))
AV = lambda **kw: ("artificial_viscosity", list(kw.items()))
PV = lambda **kw: ("physical_viscosity", list(dict(dict(coeff="3.0e-15", epsilon="0.1"), **kw).items()))
AH = lambda **kw: ("ambient_heating", list(dict(dict(heating_rate="1.0e-4"), **kw).items()))

CASES = {
    # name: (generator, gen kwargs, config kwargs, n_steps, frames to keep)
    "ot_pp_rk2":   ("orszag_tang", dict(nx=NX, ny=NY, zfull=True), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("periodic", "periodic"), **INACTIVE_FLOORS), 10, (1, 2, 10)),
    "ot_pp_euler": ("orszag_tang", dict(nx=NX, ny=NY, zfull=True), dict(integrator="euler", xb=("periodic", "periodic"), yb=("periodic", "periodic"), **INACTIVE_FLOORS), 4, (1, 4)),
    "ot_pp_rk4":   ("orszag_tang", dict(nx=NX, ny=NY, zfull=True), dict(integrator="rk4", xb=("periodic", "periodic"), yb=("periodic", "periodic"), **INACTIVE_FLOORS), 4, (1, 4)),
    "ot2d_pp_rk2_100": ("orszag_tang", dict(nx=NX, ny=NY, zfull=False), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("periodic", "periodic"), **INACTIVE_FLOORS), 100, (1, 50, 100)),
    "loop_pf_rk2": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), **SOLAR_FLOORS), 10, (1, 2, 10)),
    "loop_ro_rk2": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=("reflect", "reflect"), yb=("open", "open"), **SOLAR_FLOORS), 10, (1, 2, 10)),
    "loop_oo_rk4": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk4", xb=("open", "open"), yb=("open", "open"), **SOLAR_FLOORS), 4, (1, 4)),
    "loop_ff_euler": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="euler", xb=("fixed", "fixed"), yb=("fixed", "fixed"), **SOLAR_FLOORS), 6, (1, 6)),
    "loop_mixed_rk2": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=("fixed", "reflect"), yb=("open", "fixed"), density_min=3.0e8, temp_min=1.0e4, thermal_energy_min=1.0e-6), 8, (1, 8)),
    "loop_ucnp_rk2": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=("open_ucnp", "open_ucnp"), yb=("open_ucnp", "open_ucnp"), **SOLAR_FLOORS), 6, (1, 6)),
    # physics modules (default.config:57-76 parameter values)
    "loop_tc_euler": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), modules=[TC()], **SOLAR_FLOORS), 4, (1, 4)),
    "loop_tc_sat_rk2": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), modules=[TC(flux_saturation="true", time_integrator="rk2")], **SOLAR_FLOORS), 4, (1, 4)),
    "loop_tc_sat_rk4": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="euler", xb=("reflect", "reflect"), yb=("fixed", "open"), modules=[TC(flux_saturation="true", time_integrator="rk4")], **SOLAR_FLOORS), 3, (1, 3)),
    "loop_rl_euler": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), modules=[RL()], **SOLAR_FLOORS), 4, (1, 4)),
    "loop_rl_rk4": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), modules=[RL(time_integrator="rk4")], **SOLAR_FLOORS), 3, (1, 3)),
    "loop_ah": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), modules=[AH()], **SOLAR_FLOORS), 4, (1, 4)),
    "loop_ah_exp": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), modules=[AH(exp_mode="true", exp_base_heating_rate="3.0e-4", exp_scale_height="6.0e8")], **SOLAR_FLOORS), 3, (1, 3)),
    "loop_solar_all": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"),
                       modules=[TC(flux_saturation="true"), RL(time_integrator="rk2"), AH()], **SOLAR_FLOORS), 6, (1, 6)),
    # artificial viscosity (source/modules/viscosity.cpp): RHS terms (strength <= 1) and sub-cycled hyper-viscosity (> 1)
    "loop_visc_rhs": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"),
                      modules=[AV(visc_opt="local,global,local", visc_strength="0.5,0.3,0.2", visc_vars_to_diff="v_x,v_y,temp", visc_vars_to_evol="mom_x,mom_y,thermal_energy",
                                  visc_length="0,0,0", visc_species="i,i,i")], **SOLAR_FLOORS), 5, (1, 5)),
    "loop_visc_boundary_hv": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk4", xb=("reflect", "open"), yb=("fixed", "open"),
                      modules=[AV(visc_opt="boundary,global,boundary_global", visc_strength="0.8,3.0,0.6", visc_vars_to_diff="v_x,v_y,mom_z", visc_vars_to_evol="mom_x,mom_y,mom_z",
                                  visc_length="5.0e8,0,8.0e8", visc_species="i,i,i", hv_time_integrator="rk2", hv_epsilon="1.0", gradient_correction="true")], **SOLAR_FLOORS), 4, (1, 4)),
    "ot_visc_hv_rk4": ("orszag_tang", dict(nx=NX, ny=NY, zfull=True), dict(integrator="euler", xb=("periodic", "periodic"), yb=("periodic", "periodic"),
                      modules=[AV(visc_opt="local,global", visc_strength="2.5,0.4", visc_vars_to_diff="v_x,temp", visc_vars_to_evol="mom_x,thermal_energy",
                                  visc_length="0,0", visc_species="i,i", hv_time_integrator="rk4", hv_epsilon="1.0")], **INACTIVE_FLOORS), 3, (1, 3)),
    # ... and the planes Viscosity::fileOutput appends per term (viscosity.cpp:351-376: <evolved>_dqdt, _lap, _str, _dt -- whatever the term's LAST evaluation left), all four
    # flags set explicitly (the reference leaves them uninitialised otherwise, viscosity.hpp:60-63)
    "loop_visc_diag_hv_rk2": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=("reflect", "open"), yb=("fixed", "open"),
                      modules=[AV(visc_opt="boundary,global,boundary_global,local", visc_strength="0.8,3.0,0.6,0.5", visc_vars_to_diff="v_x,v_y,mom_z,temp",
                                  visc_vars_to_evol="mom_x,mom_y,mom_z,thermal_energy", visc_length="5.0e8,0,8.0e8,0", visc_species="i,i,i,i", hv_time_integrator="rk2", hv_epsilon="1.0",
                                  gradient_correction="true", visc_output_visc="true", visc_output_lap="true", visc_output_strength="true", visc_output_timescale="true")],
                      **SOLAR_FLOORS), 3, (0, 1, 3)),
    "ot_visc_diag_hv_rk4": ("orszag_tang", dict(nx=NX, ny=NY, zfull=True), dict(integrator="euler", xb=("periodic", "periodic"), yb=("periodic", "periodic"),
                      modules=[AV(visc_opt="local,global", visc_strength="2.5,0.4", visc_vars_to_diff="v_x,temp", visc_vars_to_evol="mom_x,thermal_energy", visc_length="0,0", visc_species="i,i",
                                  hv_time_integrator="rk4", hv_epsilon="1.0", visc_output_visc="true", visc_output_lap="true", visc_output_strength="false", visc_output_timescale="true")],
                      **INACTIVE_FLOORS), 3, (1, 3)),
    "ot_visc_diag_hv_euler": ("orszag_tang", dict(nx=NX, ny=NY, zfull=True), dict(integrator="rk4", xb=("periodic", "periodic"), yb=("periodic", "periodic"),
                      modules=[AV(visc_opt="global,local", visc_strength="2.5,0.4", visc_vars_to_diff="v_y,temp", visc_vars_to_evol="mom_y,thermal_energy", visc_length="0,0", visc_species="i,i",
                                  hv_time_integrator="euler", hv_epsilon="1.0", visc_output_visc="true", visc_output_lap="true", visc_output_strength="true", visc_output_timescale="false")],
                      **INACTIVE_FLOORS), 2, (1, 2)),
    # Braginskii physical viscosity (source/modules/solar/physicalviscosity.cpp): heating + force, euler / rk2 sub-cycles, coefficient ramp
    "loop_pv_euler": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), modules=[PV()], **SOLAR_FLOORS), 4, (1, 4)),
    "loop_pv_rk2_gc": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="euler", xb=("reflect", "open"), yb=("fixed", "open"),
                       modules=[PV(time_integrator="rk2", gradient_correction="true", ramp_length="6.0e8", coeff="1.0e-14")], **SOLAR_FLOORS), 3, (1, 3)),
    "ot_pv_heat_only": ("orszag_tang", dict(nx=NX, ny=NY, zfull=True), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("periodic", "periodic"),
                        modules=[PV(force_on="false", coeff="4.0e-16", epsilon="0.2")], **INACTIVE_FLOORS), 3, (1, 3)),
    "ot_pv_force_only_rk2": ("orszag_tang", dict(nx=NX, ny=NY, zfull=True), dict(integrator="rk4", xb=("periodic", "periodic"), yb=("periodic", "periodic"),
                        modules=[PV(heating_on="false", time_integrator="rk2", coeff="4.0e-16", epsilon="0.2")], **INACTIVE_FLOORS), 3, (1, 3)),
    # ... and its output_to_file planes (viscous_heating, viscous_force_x/y/z: the sub-cycle averages of physicalviscosity.cpp:151-170,218-222), also in inactive mode
    "loop_pv_diag_rk2": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("reflect", "open"), yb=("fixed", "open"),
                         modules=[PV(time_integrator="rk2", gradient_correction="true", ramp_length="6.0e8", coeff="1.0e-14", output_to_file="true")], **SOLAR_FLOORS), 3, (1, 3)),
    "ot_pv_diag_inactive": ("orszag_tang", dict(nx=NX, ny=NY, zfull=True), dict(integrator="euler", xb=("periodic", "periodic"), yb=("periodic", "periodic"),
                            modules=[PV(coeff="4.0e-16", epsilon="0.2", inactive_mode="true", output_to_file="true")], **INACTIVE_FLOORS), 3, (1, 3)),
    # multispecies_mode (plasmadomain.hpp:134-135, evolution.cpp:36-41, fileio.cpp:164-183): cumulative electron / ion / joule heating between outputs, fed by the modules
    # with their ms_electron_heating_fraction
    "loop_ms_solar_rk2": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), multispecies=True, modules=[
        TC(flux_saturation="true", ms_electron_heating_fraction="0.7"), RL(), AH(ms_electron_heating_fraction="0.3"),
        PV(coeff="1.0e-14", time_integrator="rk2", ms_electron_heating_fraction="0.4")], **SOLAR_FLOORS), 3, (0, 1, 3)),
    "ar_ms_joule_sources_rk2": ("stratified_loop", dict(nx=26, ny=26, bump=0.5), dict(integrator="rk2", xb=("fixed", "open"), yb=("fixed", "open"), multispecies=True, modules=[
        ("anomalous_resistivity", [("time_scale", "0.3"), ("safety_factor", "0.5"), ("flood_fill_threshold", "1.5"), ("smoothing_sigma", "1.0")]),
        ("ambient_heating_sink", [("heating_rate", "2.0e-5")]),
        ("localized_heating", [("start_time", "0.0"), ("duration", "5.0"), ("max_heating_rate", "1.0e-3"), ("stddev_x", "3.0"), ("stddev_y", "4.0"), ("center_x", "2.0"), ("center_y", "8.0"),
                               ("ramp_time", "1.0"), ("ms_electron_heating_fraction", "0.2")])], **SOLAR_FLOORS), 3, (1, 3)),
    # inactive_mode of thermal_conduction / radiative_losses (thermalconduction.cpp:109, radiativelosses.cpp:98): output and cumulative planes are formed, the state is not touched
    "loop_inactive_tc_rl_rk2": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), multispecies=True, modules=[
        TC(flux_saturation="true", inactive_mode="true", ms_electron_heating_fraction="0.7"), RL(inactive_mode="true", time_integrator="rk2"), AH()], **SOLAR_FLOORS), 3, (1, 3)),
    # two-fluid equation set (source/equationsets/ideal2F.cpp, non-sub-cycled Maxwell update) + EIC thermalization (BASELINE.json configs[2])
    "tf_ucnp_rk2": ("ucnp_cloud", CLOUD, dict(integrator="rk2", xb=UC, yb=UC, **TF, **UCNP_FLOORS), 8, (1, 8)),
    "tf_ucnp_eic_rk2": ("ucnp_cloud", CLOUD, dict(integrator="rk2", xb=UC, yb=UC, modules=[EIC], **TF, **UCNP_FLOORS), 8, (1, 8)),
    "tf_pp_euler": ("ucnp_cloud", CLOUD, dict(integrator="euler", xb=PP, yb=PP, **TF, **UCNP_FLOORS), 5, (1, 5)),
    "tf_pu_eic_rk4": ("ucnp_cloud", CLOUD, dict(integrator="rk4", xb=PP, yb=UC, modules=[EIC], **TF, **UCNP_FLOORS), 4, (1, 4)),
    "tf_fr_rk2": ("ucnp_cloud", CLOUD, dict(integrator="rk2", xb=("fixed", "reflect"), yb=("reflect", "fixed"), **TF, **UCNP_FLOORS), 6, (1, 6)),
    "tf_floors_nocurl_rk2": ("ucnp_cloud", CLOUD, dict(integrator="rk2", xb=UC, yb=PP, eqs="ideal_2F", eqs_block=[("use_sub_cycling", "false"), ("remove_curl_terms", "true")],
                             density_min=3.0e6, temp_min=1.0e-3, thermal_energy_min=2.0e-10), 5, (1, 5)),
    # ---- "next" rows of SURVEY 8f: the method-of-characteristics open boundary and the small solar modules (prefixes moc_ / sm_ / ar_:
    #      pinned for the oracle by tests/test_oracle_golden.py, and for the not-yet-GPU-validated device paths by tests/test_zz_gpu_unvalidated.py)
    "moc_y2_euler": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.4), dict(integrator="euler", xb=PP, yb=("fixed", "open_moc"), **SOLAR_FLOORS), 5, (1, 5)),
    "moc_all_visc_rk2": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.4), dict(integrator="rk2", xb=("open_moc", "open_moc"), yb=("open_moc", "open_moc"),
                         eqs_block=[("global_viscosity", "0.1")], **SOLAR_FLOORS), 5, (1, 5)),
    "moc_x1_mixed_rk4": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.4), dict(integrator="rk4", xb=("open_moc", "reflect"), yb=("fixed", "open"), **SOLAR_FLOORS), 4, (1, 4)),
    "sm_sink_heat_mass_rk2": ("stratified_loop", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=PP, yb=("fixed", "fixed"), modules=[
        ("ambient_heating_sink", [("exp_mode", "true"), ("exp_base_heating_rate", "2.0e-5"), ("exp_scale_height", "8.0e8"), ("center_x", "2.0e9"), ("half_width", "1.5e9")]),
        ("localized_heating", [("start_time", "0.0"), ("duration", "5.0"), ("max_heating_rate", "1.0e-3"), ("stddev_x", "3.0"), ("stddev_y", "4.0"), ("center_x", "2.0"), ("center_y", "8.0"), ("ramp_time", "1.0")]),
        ("mass_injection", [("start_time", "0.5"), ("duration", "10.0"), ("max_injection_rate", "1.0e6"), ("stddev_x", "3.0"), ("stddev_y", "3.0"), ("center_x", "12.0"), ("center_y", "10.0")])],
        **SOLAR_FLOORS), 6, (1, 6)),
    "sm_momentum_divclean_rk2": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="rk2", xb=("fixed", "open"), yb=("reflect", "fixed"), modules=[
        ("momentum_injection", [("start_time", "0.0"), ("duration", "50.0"), ("max_accel", "1.0e3"), ("stddev_x", "4.0"), ("stddev_y", "3.0"), ("center_x", "13.0"), ("center_y", "9.0"),
                                ("dir_x", "1.0"), ("dir_y", "0.5"), ("template_angle", "20.0"), ("oscillatory", "true"), ("oscillation_period", "3.0")]),
        ("div_cleaning", [("epsilon", "0.1"), ("time_scale", "5.0")])], **SOLAR_FLOORS), 5, (1, 5)),
    "sm_field_heating_euler": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.5), dict(integrator="euler", xb=PP, yb=("fixed", "open"), modules=[
        ("field_heating", [("coeff", "1.0"), ("current_pow", "1.0"), ("b_pow", "0.5"), ("n_pow", "0.25"), ("roc_pow", "0.5")])], **SOLAR_FLOORS), 4, (1, 4)),
    "sm_outflow_dynamic_rk2": ("stratified_loop", dict(nx=NX, ny=NY, bump=0.4), dict(integrator="rk2", xb=PP, yb=("fixed", "open"), modules=[
        ("boundary_outflow", [("max_accel", "2.0e3"), ("falloff_length", "6.0e8"), ("boundary", "y_bound_2"), ("falloff_shape", "exp"), ("feather_length", "3.0e8"),
                              ("field_aligned_mode", "true"), ("dynamic_mode", "true"), ("dynamic_time", "10.0"), ("dynamic_target_speed", "2.0e6")])], **SOLAR_FLOORS), 5, (1, 5)),
    # (square grid: circularMask / currentThresholdMask loop j over result.rows(), anomalousresistivity.cpp:285,297 -- the reference aborts when xdim > ydim)
    "ar_floodfill_rk2": ("stratified_loop", dict(nx=26, ny=26, bump=0.5), dict(integrator="rk2", xb=("fixed", "open"), yb=("fixed", "open"), modules=[
        ("anomalous_resistivity", [("time_scale", "0.3"), ("safety_factor", "0.5"), ("flood_fill_threshold", "1.5"), ("smoothing_sigma", "1.0")])], **SOLAR_FLOORS), 3, (1, 3)),
    # output_to_file planes of anomalous_resistivity (anomalous_diffusivity, anomalous_template, joule_heating) and field_heating
    "ar_diag_planes": ("stratified_loop", dict(nx=26, ny=26, bump=0.5), dict(integrator="euler", xb=("fixed", "open"), yb=("fixed", "open"), modules=[
        ("anomalous_resistivity", [("time_scale", "0.3"), ("safety_factor", "0.5"), ("flood_fill_threshold", "1.5"), ("smoothing_sigma", "1.0"), ("output_to_file", "true")]),
        ("field_heating", [("coeff", "1.0e-7"), ("current_pow", "0.5"), ("b_pow", "1.0"), ("output_to_file", "true")])], **SOLAR_FLOORS), 3, (0, 1, 3)),
    # IdealMHD2E (source/equationsets/idealmhd2E.cpp): one fluid, separate ion / electron thermal energies
    "e2_mixed_rk2": ("two_energy", dict(nx=NX, ny=NY), dict(integrator="rk2", xb=("reflect", "open"), yb=("fixed", "open"), eqs="ideal_mhd_2E", **SOLAR_FLOORS), 8, (1, 8)),
    "e2_pp_ucnp_rk4": ("two_energy", dict(nx=NX, ny=NY), dict(integrator="rk4", xb=PP, yb=UC, eqs="ideal_mhd_2E", **SOLAR_FLOORS), 4, (1, 4)),
    # the UCNP configuration of that set: ideal_mhd_2E + eic_thermalization on the Sr+ cloud (eic_thermalization.cpp:27-44)
    "e2_ucnp_eic_rk2": ("ucnp_cloud_2e", dict(nx=NX - 1, ny=NY + 1, drift=20.0, bfield=0.01), dict(integrator="rk2", xb=UC, yb=("open_ucnp", "fixed"), eqs="ideal_mhd_2E", modules=[EIC], **UCNP_FLOORS), 6, (1, 6)),
    # ... with artificial_viscosity as well (the fourth module of the UCNP set): right-hand-side terms on a momentum and the electron energy, a hyper-viscous rk2 term
    "e2_ucnp_visc_rk2": ("ucnp_cloud_2e", dict(nx=NX - 1, ny=NY + 1, drift=20.0, bfield=0.01), dict(integrator="rk2", xb=UC, yb=("open_ucnp", "fixed"), eqs="ideal_mhd_2E", modules=[
        AV(visc_opt="local,global,global", visc_strength="0.5,3.0,0.3", visc_vars_to_diff="v_x,v_y,e_temp", visc_vars_to_evol="mom_x,mom_y,e_thermal_energy", visc_length="0,0,0",
           visc_species="i,i,e", hv_time_integrator="rk2", hv_epsilon="1.0", gradient_correction="true", visc_output_visc="false", visc_output_lap="false",
           visc_output_strength="false", visc_output_timescale="false")], **UCNP_FLOORS), 4, (1, 4)),
    # configs[0] of BASELINE.json: the reference's own example.state (fixed up: + be_z, mom_z, bi_z zero planes)
    "example_state_rk2": ("example_state", dict(), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), **SOLAR_FLOORS), 20, (1, 20)),
    # configs[1] of BASELINE.json: the reference's own shipped inputs with the solar module set of default.config:57-76 -- thermal conduction with flux
    # saturation, radiative losses, ambient heating -- on the gravity-stratified bipolar loop (example.state, fixed up) and on examples/solar/solar_gaussian.state
    "example_state_solar_modules_rk2": ("example_state", dict(), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), modules=[
        TC(flux_saturation="true", output_to_file="false"), RL(output_to_file="false"), AH()], **SOLAR_FLOORS), 6, (1, 6)),
    "solar_gaussian_state_modules_rk2": ("solar_gaussian_state", dict(), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "fixed"), modules=[
        TC(flux_saturation="true", output_to_file="false"), RL(output_to_file="false"), AH()], **SOLAR_FLOORS), 6, (1, 6)),
}


def example_state():
    meta, planes = refrun.read_state("/root/reference/example.state")
    z = np.zeros_like(planes["rho"])
    for k in ("be_z", "mom_z", "bi_z"):
        planes.setdefault(k, z.copy())
    order = ["d_x", "d_y", "pos_x", "pos_y", "be_x", "be_y", "be_z", "rho", "temp", "mom_x", "mom_y", "mom_z",
             "bi_x", "bi_y", "bi_z", "grav_x", "grav_y"]
    return dict(planes={k: planes[k] for k in order}, ion_mass=meta["ion_mass"], adiabatic_index=meta["adiabatic_index"])


def solar_gaussian_state():
    meta, planes = refrun.read_state("/root/reference/examples/solar/solar_gaussian.state")
    z = np.zeros_like(planes["rho"])
    for k in ("be_z", "mom_z", "bi_z"):
        planes.setdefault(k, z.copy())
    order = ["d_x", "d_y", "pos_x", "pos_y", "be_x", "be_y", "be_z", "rho", "temp", "mom_x", "mom_y", "mom_z",
             "bi_x", "bi_y", "bi_z", "grav_x", "grav_y"]
    return dict(planes={k: planes[k] for k in order}, ion_mass=meta["ion_mass"], adiabatic_index=meta["adiabatic_index"])


def interior(cfg, nx, ny):
    """bounds of the time-step minimum: the interior, widened by the ghost zone on open_moc sides (plasmadomain.cpp:138-159)"""
    wide = ("periodic", "open_moc")
    xl = 0 if cfg["xb"][0] in wide else 2
    xu = nx - 1 if cfg["xb"][1] in wide else nx - 3
    yl = 0 if cfg["yb"][0] in wide else 2
    yu = ny - 1 if cfg["yb"][1] in wide else ny - 3
    return xl, xu, yl, yu


def make(name):
    gen, gkw, ckw, nsteps, keep = CASES[name]
    s = example_state() if gen == "example_state" else solar_gaussian_state() if gen == "solar_gaussian_state" else getattr(synthetic, gen)(**gkw)
    tmp = Path(tempfile.mkdtemp(prefix="golden_"))
    try:
        refrun.write_state(tmp / "in.state", s["planes"], s["ion_mass"], s["adiabatic_index"])
        out_vars = OUT_VARS_2F if ckw.get("eqs") == "ideal_2F" else OUT_VARS_2E if ckw.get("eqs") == "ideal_mhd_2E" else OUT_VARS
        cfg = refrun.ideal_mhd_config(max_iterations=nsteps, output_flags=out_vars, std_out_interval=1, **ckw)
        _, stdout = refrun.run_reference(tmp / "in.state", cfg, tmp / "out", threads=8)
        _, frames = refrun.read_out(tmp / "out" / "mhd.out")
        assert len(frames) == nsteps + 1, (len(frames), nsteps)
        nx, ny = s["planes"]["d_x"].shape
        xl, xu, yl, yu = interior(ckw, nx, ny)
        eps = ckw.get("epsilon", 0.2)
        steps = np.array([eps * np.nanmin(f["dt"][xl:xu + 1, yl:yu + 1]) for f in frames[:-1]])
        times = np.array([f["t"] for f in frames])
        out = {"in_" + k: v for k, v in s["planes"].items()}
        out["ion_mass"] = np.float64(s["ion_mass"])
        out["adiabatic_index"] = np.float64(s["adiabatic_index"])
        out["steps"] = steps
        out["times_6digits"] = times
        for fi in keep:
            for v in out_vars:
                out["f%d_%s" % (fi, v)] = frames[fi][v]
            for v in frames[fi]:
                if v not in out_vars and v != "t":       # module output planes (thermal_conduction, rad, ...)
                    out["f%d_mod_%s" % (fi, v)] = frames[fi][v]
        desc = dict(case=name, out_vars=out_vars, generator=gen, gen_kwargs=gkw, n_steps=nsteps, keep=list(keep),
                    config={k: v for k, v in ckw.items()}, config_text=cfg,
                    subcycle_log=[ln for ln in stdout.splitlines() if "Subcycle" in ln][:nsteps])
        out["desc"] = np.array(json.dumps(desc))
        np.savez_compressed(HERE / (name + ".npz"), **out)
        print("%-20s steps=%d first dt=%s  -> %s.npz" % (name, nsteps, float(steps[0]).hex(), name))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    assert refrun.have_reference(), "build the reference first: make -C oracle ref"
    for nm in (sys.argv[1:] or list(CASES)):
        make(nm)
