"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/spruce_oracle.c (the CPU restatement of the
reference's per-timestep advance).  Allowed importers: tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline leg.  The product path (spruce_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
LIB_PATH = ORACLE_DIR / "_ref" / "liboracle.so"

BC = {"periodic": 0, "open": 1, "fixed": 2, "reflect": 3, "open_moc": 4, "open_ucnp": 5}   # plasmadomain.hpp:22-26
TI = {"euler": 0, "rk2": 1, "rk4": 2}                                                       # plasmadomain.hpp:29-32
VARS = ["rho", "temp", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "grav_x", "grav_y",
        "n", "press", "thermal_energy", "v_x", "v_y", "v_z", "kinetic_energy",
        "b_x", "b_y", "b_z", "b_mag", "b_hat_x", "b_hat_y", "b_hat_z", "dt"]                # idealmhd.hpp:19-23
EVOLVED = ["rho", "mom_x", "mom_y", "mom_z", "thermal_energy", "bi_x", "bi_y", "bi_z"]     # idealmhd.hpp:32-34
DOMAIN = {"d_x": 0, "d_y": 1, "be_x": 2, "be_y": 3, "be_z": 4, "pos_x": 5, "pos_y": 6, "mask": 7}

_lib = None


def build():
    subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int] * 7 + [C.c_double] * 8
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_plane.restype = C.POINTER(C.c_double)
        L.oracle_plane.argtypes = [C.c_void_p, C.c_int]
        L.oracle_setup.argtypes = [C.c_void_p]
        L.oracle_propagate.argtypes = [C.c_void_p]
        L.oracle_step.restype = C.c_double
        L.oracle_step.argtypes = [C.c_void_p]
        L.oracle_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.oracle_time.restype = C.c_double
        L.oracle_time.argtypes = [C.c_void_p]
        L.oracle_subcycles.restype = C.c_int
        L.oracle_subcycles.argtypes = [C.c_void_p, C.c_int]
        L.oracle_rhs.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_operator.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 3
        L.oracle_set_thermal_conduction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.oracle_set_radiative_losses.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
        L.oracle_set_viscosity.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.oracle_add_viscosity_term.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.oracle_set_anomalous_resistivity.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_anomalous_subcycles.restype = C.c_int
        L.oracle_small_module_hooks.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.oracle_set_time.argtypes = [C.c_void_p, C.c_double]
        L.oracle_anomalous_diffusivity.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_module_output.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.oracle_anomalous_iterate.argtypes = [C.c_void_p, C.c_double]
        L.oracle_anomalous_core.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.c_int]
        L.oracle_anomalous_state.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_anomalous_subcycles.argtypes = [C.c_void_p]
        L.oracle_add_small_module.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int]
        L.oracle_small_module_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.oracle_small_module_plane.restype = C.c_int
        L.oracle_outflow_mean.argtypes = [C.c_void_p, C.c_int]
        L.oracle_outflow_mean.restype = C.c_double
        L.oracle_coulomb_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.oracle_coulomb_plane.restype = C.c_int
        L.oracle_set_global_viscosity.argtypes = [C.c_void_p, C.c_double]
        L.oracle_apply_moc_thresholding.argtypes = [C.c_void_p]
        L.oracle_sg_filter.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_set_moc_limiting.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double]
        L.oracle_set_physical_viscosity.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_physical_viscosity_iterate.argtypes = [C.c_void_p, C.c_double]
        L.oracle_set_module_inactive.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_set_multispecies.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_ms_fraction.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.oracle_ms_plane.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.oracle_ms_plane.restype = C.c_int
        L.oracle_ms_reset.argtypes = [C.c_void_p]
        L.oracle2e_create.argtypes = [C.c_void_p]; L.oracle2e_create.restype = C.c_void_p
        L.oracle2e_destroy.argtypes = [C.c_void_p]
        L.oracle2e_set_eic.argtypes = [C.c_void_p, C.c_int]
        L.oracle2e_set_viscosity.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.oracle2e_add_viscosity_term.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.oracle2e_plane.argtypes = [C.c_void_p, C.c_int]; L.oracle2e_plane.restype = C.POINTER(C.c_double)
        L.oracle2e_setup.argtypes = [C.c_void_p]
        L.oracle2e_step.argtypes = [C.c_void_p]; L.oracle2e_step.restype = C.c_double
        L.oracle2e_rhs.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.oracle2e_time.argtypes = [C.c_void_p]; L.oracle2e_time.restype = C.c_double
        L.oracle2f_apply_ghosts.argtypes = [C.c_void_p]
        L.oracle2f_create.restype = C.c_void_p
        L.oracle2f_create.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle2f_destroy.argtypes = [C.c_void_p]
        L.oracle2f_plane.restype = C.POINTER(C.c_double)
        L.oracle2f_plane.argtypes = [C.c_void_p, C.c_int]
        L.oracle2f_setup.argtypes = [C.c_void_p]
        L.oracle2f_step.restype = C.c_double
        L.oracle2f_step.argtypes = [C.c_void_p]
        L.oracle2f_rhs.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.oracle2f_time.restype = C.c_double
        L.oracle2f_time.argtypes = [C.c_void_p]
        L.oracle_set_ambient_heating.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def anomalous_params(*, time_scale=1.0, frobenius_metric_coeff=1.0e50, smoothing_sigma=3.0, safety_factor=1.0, metric_smoothing=True,
                     time_integrator="euler", template_mode="flood_fill", flood_fill_max_radius=-1.0, flood_fill_argmin_radius=5.0e9,
                     flood_fill_min_current=-1.0, flood_fill_current_ramp_length=1.0e-5, flood_fill_threshold=1.0, resistivity_model="time_scale",
                     gradient_correction=False, resistivity_model_params=(0.0, 0.0, 0.0)):
    """the 17 numbers oracle_set_anomalous_resistivity takes, with the reference's defaults (anomalousresistivity.hpp:16-39)"""
    mp = list(resistivity_model_params) + [0.0, 0.0, 0.0]
    return np.array([time_scale, frobenius_metric_coeff, smoothing_sigma, safety_factor, float(metric_smoothing), TI[time_integrator or "euler"],
                     float(template_mode == "flood_fill"), flood_fill_max_radius, flood_fill_argmin_radius, flood_fill_min_current, flood_fill_current_ramp_length,
                     flood_fill_threshold, {"time_scale": 0.0, "syntelis_19": 1.0, "ys_94": 2.0}[resistivity_model], float(gradient_correction), mp[0], mp[1], mp[2]])


class Oracle:
    """One ideal-MHD domain evolved by the C restatement.  planes: the reference's .state planes."""

    def __init__(self, planes, ion_mass, adiabatic_index, *, xb=("periodic", "periodic"), yb=("periodic", "periodic"),
                 integrator="rk2", epsilon=0.2, density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6,
                 open_strength=1.0, open_decay=0.5, setup=True, moc_limiting=None):
        L = lib()
        nx, ny = planes["rho"].shape
        self.nx, self.ny = nx, ny
        self.h = L.oracle_create(nx, ny, BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]], TI[integrator],
                                 ion_mass, adiabatic_index, epsilon, density_min, temp_min, thermal_energy_min,
                                 open_strength, open_decay)
        for name, a in planes.items():
            if name in ("mask",):
                continue
            self.view(name)[...] = a
        if moc_limiting:                               # equation-set options are parsed before setupEquationSet: its derived pass already applies them
            self.set_moc_limiting(**moc_limiting)
        if setup:
            L.oracle_setup(self.h)

    def view(self, name) -> np.ndarray:
        which = DOMAIN[name] if name in DOMAIN else 100 + VARS.index(name)
        p = lib().oracle_plane(self.h, which)
        return np.ctypeslib.as_array(p, shape=(self.nx, self.ny))

    def get(self, name) -> np.ndarray:
        return self.view(name).copy()

    def evolved(self) -> dict:
        return {k: self.get(k) for k in EVOLVED}

    def step(self) -> float:
        return lib().oracle_step(self.h)

    def run(self, nsteps) -> np.ndarray:
        dts = np.zeros(nsteps)
        lib().oracle_run(self.h, nsteps, _dp(dts))
        return dts

    def propagate(self):
        lib().oracle_propagate(self.h)

    def rhs(self) -> np.ndarray:
        k = np.zeros((8, self.nx, self.ny))
        lib().oracle_rhs(self.h, _dp(k))
        return k

    def operator(self, op, index, q, vel=None) -> np.ndarray:
        ops = {"derivative1D": 0, "secondDerivative1D": 1, "laplacian": 2, "transportDerivative1D": 3}
        q = np.ascontiguousarray(q, dtype=np.float64)
        vel = q if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        out = np.zeros_like(q)
        lib().oracle_operator(self.h, ops[op], index, _dp(q), _dp(vel), _dp(out))
        return out

    @property
    def time(self) -> float:
        return lib().oracle_time(self.h)

    def subcycles(self, which) -> int:
        return lib().oracle_subcycles(self.h, {"thermal_conduction": 1, "radiative_losses": 2, "physical_viscosity": 5}[which])

    SMALL = {"ambient_heating_sink": (6, ["heating_rate", "exp_mode", "exp_base_heating_rate", "exp_scale_height", "center_x", "half_width"]),
             "localized_heating": (7, ["start_time", "duration", "max_heating_rate", "stddev_x", "stddev_y", "center_x", "center_y", "ramp_time"]),
             "mass_injection": (8, ["start_time", "duration", "max_injection_rate", "stddev_x", "stddev_y", "center_x", "center_y"]),
             "momentum_injection": (9, ["start_time", "duration", "max_accel", "stddev_x", "stddev_y", "center_x", "center_y", "dir_x", "dir_y", "template_angle", "oscillatory", "oscillation_period"]),
             "div_cleaning": (10, ["epsilon", "time_scale"]),
             "field_heating": (11, ["coeff", "current_pow", "b_pow", "n_pow", "roc_pow", "inactive_mode"]),
             # boundary: 0 x_bound_1, 1 x_bound_2, 2 y_bound_1, 3 y_bound_2 ; falloff_shape: 0 exp, 1 gaussian, 2 flat
             "sg_filtering": (13, ["filter_interval"]),
             "coulomb_explosion": (14, ["timescale", "lengthscale", "strength"]),
             "global_temperature": (15, ["gt_strength", "gt_use_diffusion"]),
             "boundary_outflow": (12, ["max_accel", "falloff_length", "boundary", "falloff_shape", "feather_length", "field_aligned_mode", "dynamic_mode", "dynamic_time", "dynamic_target_speed"])}

    def add_small_module(self, name: str, **kw):
        """Small solar source-term modules (oracle/solar_small_modules_oracle.inc); keyword names = the reference's config keys."""
        kind, keys = self.SMALL[name]
        p = np.array([float(kw.get(k, 0.0)) for k in keys])
        lib().oracle_add_small_module(self.h, kind, _dp(p), p.size)

    def small_module_plane(self, idx: int, k: int = 0):
        """Template / coefficient plane k of the idx-th small module (None until it exists); for checking the product's host-built templates."""
        out = np.zeros((self.nx, self.ny))
        return out if lib().oracle_small_module_plane(self.h, idx, k, _dp(out)) else None

    def coulomb_plane(self, idx: int, name: str):
        """output_to_file plane F_x / F_y / dP_x / dP_y of the coulomb_explosion module at position idx among the small modules; raises if the reference
        would have aborted (an empty radial bin, a radius outside the bin centres)"""
        out = np.zeros((self.nx, self.ny))
        rc = lib().oracle_coulomb_plane(self.h, idx, ["F_x", "F_y", "dP_x", "dP_y"].index(name), _dp(out))
        if rc != 1:
            raise RuntimeError("coulomb_explosion: rc %d (2: the reference aborts on this grid)" % rc)
        return out

    def outflow_mean(self, idx: int) -> float:
        """BoundaryOutflow::computeMeanOutflow of the idx-th small module on the current state."""
        return lib().oracle_outflow_mean(self.h, idx)

    def set_anomalous_resistivity(self, *, time_scale=1.0, frobenius_metric_coeff=1.0e50, smoothing_sigma=3.0, safety_factor=1.0, metric_smoothing=True,
                                  time_integrator="euler", template_mode="flood_fill", flood_fill_max_radius=-1.0, flood_fill_argmin_radius=5.0e9,
                                  flood_fill_min_current=-1.0, flood_fill_current_ramp_length=1.0e-5, flood_fill_threshold=1.0, resistivity_model="time_scale",
                                  gradient_correction=False, resistivity_model_params=(0.0, 0.0, 0.0)):
        """AnomalousResistivity with the reference's defaults (anomalousresistivity.hpp:16-39); call after setup."""
        p = anomalous_params(**{k: v for k, v in locals().items() if k != "self"})
        lib().oracle_set_anomalous_resistivity(self.h, _dp(p))

    def small_module_hooks(self, phase: int, step: float):
        """test accessor: the small solar modules' hooks of one phase only (0 preIterate, 1 iterate, 2 postIterate) on the current planes"""
        lib().oracle_small_module_hooks(self.h, C.c_int(phase), C.c_double(step))

    def set_time(self, t: float):
        lib().oracle_set_time(self.h, C.c_double(t))

    def anomalous_core(self, dt: float, raw_commit: bool = False):
        """test accessor: one iterateModule(dt) of anomalous_resistivity without write-back / propagateChanges -> (bi_x, bi_y, bi_z, thermal_energy)"""
        out = np.zeros((4, self.nx, self.ny))
        lib().oracle_anomalous_core(self.h, C.c_double(dt), _dp(out), C.c_int(int(raw_commit)))
        return out

    def set_module_inactive(self, module: str, on: bool = True):
        """inactive_mode of thermal_conduction / radiative_losses (thermalconduction.cpp:109, radiativelosses.cpp:98): evaluated for the output and cumulative planes, not applied"""
        lib().oracle_set_module_inactive(self.h, {"thermal_conduction": 1, "radiative_losses": 2}[module], int(on))

    def set_multispecies(self, on: bool = True, **fractions):
        """multispecies_mode = true; fractions: ms_electron_heating_fraction per module (thermal_conduction=, radiative_losses=, ambient_heating=, physical_viscosity=,
        ambient_heating_sink=, localized_heating=), the reference's defaults otherwise"""
        lib().oracle_set_multispecies(self.h, int(on))
        ids = {"thermal_conduction": 1, "radiative_losses": 2, "ambient_heating": 3, "physical_viscosity": 5, "ambient_heating_sink": 6, "localized_heating": 7}
        for k, v in fractions.items():
            lib().oracle_set_ms_fraction(self.h, ids[k], float(v))

    def ms_plane(self, name: str):
        """cumulative_electron_heating / cumulative_ion_heating / cumulative_joule_heating since the last ms_reset()"""
        out = np.zeros((self.nx, self.ny))
        ok = lib().oracle_ms_plane(self.h, {"cumulative_electron_heating": 0, "cumulative_ion_heating": 1, "cumulative_joule_heating": 2}[name], _dp(out))
        return out if ok else None

    def ms_reset(self):
        lib().oracle_ms_reset(self.h)

    def physical_viscosity_iterate(self, dt: float):
        """PhysicalViscosity::iterateModule alone, on the current state"""
        lib().oracle_physical_viscosity_iterate(self.h, C.c_double(dt))

    def anomalous_iterate(self, dt: float):
        """test accessor: one complete iterateModule(dt) of anomalous_resistivity (write-back and propagateChanges included), nothing else"""
        lib().oracle_anomalous_iterate(self.h, C.c_double(dt))

    def anomalous_state(self):
        ij = (C.c_int * 2)()
        t = np.zeros((self.nx, self.ny))
        lib().oracle_anomalous_state(self.h, ij, _dp(t))
        return (ij[0], ij[1]), t

    def module_output(self, name: str):
        """output_to_file plane of thermal_conduction / radiative_losses / physical_viscosity after the last step ("thermal_conduction", "flux_saturation", "rad",
        "viscous_heating", "viscous_force_x" / "_y" / "_z"); None before the module ran"""
        out = np.zeros((self.nx, self.ny))
        which = {"thermal_conduction": 0, "flux_saturation": 1, "rad": 2, "viscous_heating": 3, "viscous_force_x": 4, "viscous_force_y": 5, "viscous_force_z": 6}[name]
        ok = lib().oracle_module_output(self.h, which, _dp(out))
        return out if ok else None

    def anomalous_diffusivity(self):
        out = np.zeros((self.nx, self.ny))
        lib().oracle_anomalous_diffusivity(self.h, _dp(out))
        return out

    def anomalous_subcycles(self) -> int:
        return lib().oracle_anomalous_subcycles(self.h)

    def set_global_viscosity(self, v: float):
        lib().oracle_set_global_viscosity(self.h, v)

    def apply_moc_thresholding(self):
        """The two limiter passes alone, on the planes as they are (for checking the product's limiter on arbitrary input)."""
        lib().oracle_apply_moc_thresholding(self.h)

    def sg_filter(self, plane):
        """SGFilter::singleVarSavitzkyGolay (sgfilter.cpp:46-82) on an arbitrary plane; returns the filtered copy."""
        a = np.array(plane, dtype=np.float64, order="C", copy=True)
        lib().oracle_sg_filter(self.h, _dp(a))
        return a

    def set_moc_limiting(self, *, b_limiting=False, b_lower=0.1, b_upper=10.0, mom_limiting=False, mom_lower=0.1, mom_upper=10.0):
        """moc_b_limiting / moc_mom_limiting (idealmhd.cpp:107-223); call before the first step (the setup's derived pass has already run without them)."""
        lib().oracle_set_moc_limiting(self.h, int(b_limiting), b_lower, b_upper, int(mom_limiting), mom_lower, mom_upper)

    def set_physical_viscosity(self, coeff_plane, *, coeff, epsilon=1.0, heating_on=True, force_on=True, gradient_correction=False, integrator="euler",
                               inactive_mode=False):
        a = np.ascontiguousarray(coeff_plane, dtype=np.float64)
        lib().oracle_set_physical_viscosity(self.h, coeff, _dp(a), epsilon, int(heating_on), int(force_on), int(gradient_correction), TI[integrator], int(inactive_mode))

    def set_thermal_conduction(self, *, flux_saturation=False, integrator="euler", epsilon=0.1, dt_subcycle_min=1.0e-4, weakening_factor=1.0):
        lib().oracle_set_thermal_conduction(self.h, int(flux_saturation), TI[integrator], epsilon, dt_subcycle_min, weakening_factor)

    def set_radiative_losses(self, *, integrator="euler", cutoff_ramp=1.0e3, cutoff_temp=3.0e4, epsilon=0.1, prevent_subcycling=False):
        lib().oracle_set_radiative_losses(self.h, TI[integrator], cutoff_ramp, cutoff_temp, epsilon, int(prevent_subcycling))

    def set_ambient_heating(self, *, heating_rate=0.0, exp_mode=False, exp_base_heating_rate=0.0, exp_scale_height=1.0,
                            split_exp_mode=False, split_exp_scale_height=1.0, split_exp_start_height=0.0):
        lib().oracle_set_ambient_heating(self.h, heating_rate, int(exp_mode), exp_base_heating_rate, exp_scale_height,
                                         int(split_exp_mode), split_exp_scale_height, split_exp_start_height)

    def set_viscosity(self, terms, *, hv_integrator="euler", hv_epsilon=1.0, gradient_correction=False):
        """terms: list of dict(opt=local|global|boundary|boundary_global, strength=, var_diff=, var_evol=, species='i', strength_grid=None)"""
        lib().oracle_set_viscosity(self.h, TI[hv_integrator], hv_epsilon, int(gradient_correction))
        opts = {"local": 0, "global": 1, "boundary": 2, "boundary_global": 3}
        for tm in terms:
            sg = tm.get("strength_grid")
            sgp = _dp(np.ascontiguousarray(sg, dtype=np.float64)) if sg is not None else None
            lib().oracle_add_viscosity_term(self.h, opts[tm["opt"]], tm["strength"], VARS.index(tm["var_diff"]), VARS.index(tm["var_evol"]),
                                            ord(tm.get("species", "i")), sgp)

    def viscosity_output(self, which: str, term: int):
        """the plane Viscosity::fileOutput appends for term `term`: "dqdt", "lap", "str" or "dt" (viscosity.cpp:351-376), as the last evaluation left it"""
        out = np.zeros((self.nx, self.ny))
        ok = lib().oracle_viscosity_output(self.h, {"dqdt": 0, "lap": 1, "str": 2, "dt": 3}[which], term, _dp(out))
        return out if ok else None

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


VARS_2F = ["i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_temp", "e_temp", "bi_x", "bi_y", "bi_z", "E_x", "E_y", "E_z", "grav_x", "grav_y",
           "i_n", "e_n", "i_v_x", "i_v_y", "e_v_x", "e_v_y", "j_x", "j_y", "i_press", "e_press", "press", "i_thermal_energy", "e_thermal_energy",
           "rho", "rho_c", "n", "dn", "dt", "dt_i", "b_x", "b_y", "b_z", "b_mag", "b_mag_xy", "b_hat_x", "b_hat_y", "curlE_z", "divE", "divB", "i_dPdx", "e_dPdx"]   # ideal2F.hpp:24-31
EVOLVED_2F = ["i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_thermal_energy", "e_thermal_energy", "E_x", "E_y", "E_z", "bi_x", "bi_y", "bi_z"]


class Oracle2F:
    """One two-fluid (Ideal2F, use_sub_cycling = false) domain evolved by the C restatement (oracle/ideal2f_oracle.inc)."""

    def __init__(self, planes, ion_mass, adiabatic_index, *, xb=("open_ucnp", "open_ucnp"), yb=("open_ucnp", "open_ucnp"), integrator="rk2", epsilon=0.2,
                 density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1.0e-30, remove_curl_terms=False, eic=False, setup=True, **_unused):
        L = lib()
        nx, ny = planes["i_rho"].shape
        self.nx, self.ny = nx, ny
        self.base = L.oracle_create(nx, ny, BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]], TI[integrator], ion_mass, adiabatic_index, epsilon,
                                    density_min, temp_min, thermal_energy_min, 1.0, 0.5)
        for name, which in DOMAIN.items():
            if name in planes:
                np.ctypeslib.as_array(L.oracle_plane(self.base, which), shape=(nx, ny))[...] = planes[name]
        self.h = L.oracle2f_create(self.base, int(remove_curl_terms), int(eic))
        for name, a in planes.items():
            if name in VARS_2F:
                self.view(name)[...] = a
        if setup:
            L.oracle2f_setup(self.h)

    def view(self, name) -> np.ndarray:
        return np.ctypeslib.as_array(lib().oracle2f_plane(self.h, VARS_2F.index(name)), shape=(self.nx, self.ny))

    def get(self, name) -> np.ndarray:
        return self.view(name).copy()

    def step(self) -> float:
        return lib().oracle2f_step(self.h)

    def apply_ghosts(self):
        """updateGhostZones alone, on the planes as they are (for checking the product's ordered boundary passes on arbitrary input)."""
        lib().oracle2f_apply_ghosts(self.h)

    def rhs(self) -> np.ndarray:
        k = np.zeros((14, self.nx, self.ny))
        lib().oracle2f_rhs(self.h, _dp(k))
        return k

    @property
    def time(self) -> float:
        return lib().oracle2f_time(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib().oracle2f_destroy(self.h)
            lib().oracle_destroy(self.base)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


VARS_2E = ["rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "grav_x", "grav_y", "n", "i_press", "e_press", "press", "i_thermal_energy", "e_thermal_energy",
           "v_x", "v_y", "kinetic_energy", "b_x", "b_y", "b_mag", "b_hat_x", "b_hat_y", "dt"]                 # idealmhd2E.hpp:18-22
EVOLVED_2E = ["rho", "mom_x", "mom_y", "i_thermal_energy", "e_thermal_energy", "bi_x", "bi_y"]              # idealmhd2E.hpp:31-33


class Oracle2E:
    """One IdealMHD2E domain (one fluid, separate ion / electron thermal energies) evolved by the C restatement (oracle/ideal_mhd2e_oracle.inc)."""

    def __init__(self, planes, ion_mass, adiabatic_index, *, xb=("periodic", "periodic"), yb=("fixed", "fixed"), integrator="rk2", epsilon=0.2,
                 density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6, open_strength=1.0, open_decay=0.5, setup=True, eic=False, **_unused):
        L = lib()
        nx, ny = planes["rho"].shape
        self.nx, self.ny = nx, ny
        self.base = L.oracle_create(nx, ny, BC[xb[0]], BC[xb[1]], BC[yb[0]], BC[yb[1]], TI[integrator], ion_mass, adiabatic_index, epsilon,
                                    density_min, temp_min, thermal_energy_min, open_strength, open_decay)
        for name, which in DOMAIN.items():
            if name in planes:
                np.ctypeslib.as_array(L.oracle_plane(self.base, which), shape=(nx, ny))[...] = planes[name]
        self.h = L.oracle2e_create(self.base)
        L.oracle2e_set_eic(self.h, int(eic))             # eic_thermalization on this equation set (the UCNP configuration: ideal_mhd_2E + eic_thermalization)
        for name, a in planes.items():
            if name in VARS_2E:
                self.view(name)[...] = a
        if setup:
            L.oracle2e_setup(self.h)

    def view(self, name) -> np.ndarray:
        return np.ctypeslib.as_array(lib().oracle2e_plane(self.h, VARS_2E.index(name)), shape=(self.nx, self.ny))

    def get(self, name) -> np.ndarray:
        return self.view(name).copy()

    def set_viscosity(self, terms, *, hv_integrator="euler", hv_epsilon=1.0, gradient_correction=False):
        """artificial_viscosity on this set (configured after eic_thermalization); terms: list of dict(opt=local|global|boundary|boundary_global, strength=, var_diff=,
        var_evol=, species='i', strength_grid=None), variable names of idealmhd2E.hpp:18-22"""
        L = lib()
        L.oracle2e_set_viscosity(self.h, TI[hv_integrator], hv_epsilon, int(gradient_correction))
        opts = {"local": 0, "global": 1, "boundary": 2, "boundary_global": 3}
        for tm in terms:
            sg = tm.get("strength_grid")
            sgp = _dp(np.ascontiguousarray(sg, dtype=np.float64)) if sg is not None else None
            L.oracle2e_add_viscosity_term(self.h, opts[tm["opt"]], tm["strength"], VARS_2E.index(tm["var_diff"]), VARS_2E.index(tm["var_evol"]), ord(tm.get("species", "i")), sgp)

    def step(self) -> float:
        return lib().oracle2e_step(self.h)

    def rhs(self) -> np.ndarray:
        k = np.zeros((7, self.nx, self.ny))
        lib().oracle2e_rhs(self.h, _dp(k))
        return k

    @property
    def time(self) -> float:
        return lib().oracle2e_time(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib().oracle2e_destroy(self.h)
            lib().oracle_destroy(self.base)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
