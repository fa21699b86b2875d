"""Python mirror of the reference's PlasmaDomain / EquationSet surface for the per-timestep advance, on top of the
C ABI (spruce_b200.capi).  Method names follow the reference (source/mhd/plasmadomain.hpp, source/equationsets/
equationset.hpp): grid(name), propagateChanges(), computeTimeDerivatives(), advanceTime()/run().

All arithmetic happens on the GPU inside libspruce_b200.so; this file only moves planes and parameters.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

_DP = C.POINTER(C.c_double)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_DP)


def rank1_cell_sizes(d_x: np.ndarray, d_y: np.ndarray):
    """The reference keeps d_x, d_y as planes but requires d_x(i,j) = d_x(i), d_y(i,j) = d_y(j) (README.md:41)."""
    dx, dy = np.ascontiguousarray(d_x[:, 0]), np.ascontiguousarray(d_y[0, :])
    if not (np.array_equal(d_x, np.repeat(dx[:, None], d_x.shape[1], 1)) and np.array_equal(d_y, np.repeat(dy[None, :], d_y.shape[0], 0))):
        raise capi.SpruceError("d_x must vary with i only and d_y with j only (rectilinear grid, reference README.md:41)")
    return dx, dy


class PlasmaDomain:
    EVOLVED = ["rho", "mom_x", "mom_y", "mom_z", "thermal_energy", "bi_x", "bi_y", "bi_z"]          # idealmhd.hpp:32-34
    STATE = ["rho", "temp", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "grav_x", "grav_y"]   # idealmhd.hpp:28-30
    DOMAIN = ["be_x", "be_y", "be_z"]
    EVOLVED_2E = ["rho", "mom_x", "mom_y", "i_thermal_energy", "e_thermal_energy", "bi_x", "bi_y"]    # idealmhd2E.hpp:31-33
    STATE_2E = ["rho", "i_temp", "e_temp", "mom_x", "mom_y", "bi_x", "bi_y", "grav_x", "grav_y"]      # idealmhd2E.hpp:27-29
    # ideal2F.hpp:40-46
    EVOLVED_2F = ["i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_thermal_energy", "e_thermal_energy",
                  "E_x", "E_y", "E_z", "bi_x", "bi_y", "bi_z"]
    STATE_2F = ["i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_temp", "e_temp", "bi_x", "bi_y", "bi_z",
                "E_x", "E_y", "E_z", "grav_x", "grav_y"]

    def __init__(self, planes: dict, ion_mass: float, adiabatic_index: float, *, equation_set="ideal_mhd", eqs_options=None,
                 xb=("periodic", "periodic"), yb=("periodic", "periodic"), integrator="rk2", epsilon=0.2,
                 density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6, open_strength=1.0, open_decay=0.5,
                 time=0.0, device=-1, row0=0, nx_local=None, rank=0, n_ranks=1, setup=True):
        self.lib = capi.load()
        self.equation_set = equation_set
        if equation_set == "ideal_2F":
            self.EVOLVED, self.STATE = self.EVOLVED_2F, self.STATE_2F
        elif equation_set == "ideal_mhd_2E":
            self.EVOLVED, self.STATE = self.EVOLVED_2E, self.STATE_2E
        gx, gy = planes["d_x"].shape if planes["d_x"].ndim == 2 else (planes["d_x"].size, planes["d_y"].size)
        if planes["d_x"].ndim == 2:
            dx, dy = rank1_cell_sizes(planes["d_x"], planes["d_y"])
        else:
            dx, dy = np.ascontiguousarray(planes["d_x"], dtype=np.float64), np.ascontiguousarray(planes["d_y"], dtype=np.float64)
        self.xdim, self.ydim = int(gx), int(gy)
        self.row0 = int(row0)
        self.nx = int(self.xdim if nx_local is None else nx_local)
        cfg = capi.Config(abi_version=capi.ABI_VERSION, equation_set=capi.EQS[equation_set], xdim=self.xdim, ydim=self.ydim,
                          x_bound_1=capi.BC[xb[0]], x_bound_2=capi.BC[xb[1]], y_bound_1=capi.BC[yb[0]], y_bound_2=capi.BC[yb[1]],
                          time_integrator=capi.TI[integrator], device=device, row0=self.row0, nx_local=self.nx, rank=rank, n_ranks=n_ranks,
                          ion_mass=ion_mass, adiabatic_index=adiabatic_index, epsilon=epsilon, density_min=density_min,
                          temp_min=temp_min, thermal_energy_min=thermal_energy_min, open_boundary_strength=open_strength,
                          open_boundary_decay_base=open_decay, time=time)
        h = C.c_void_p()
        capi.check(self.lib.spruce_domain_create(C.byref(cfg), C.byref(h)))
        self.h = h
        capi.check(self.lib.spruce_set_cell_sizes(self.h, _dp(dx), dx.size, _dp(dy), dy.size))
        if equation_set == "ideal_2F":
            o = dict(use_sub_cycling=True, remove_curl_terms=False)          # Ideal2F defaults, ideal2F.hpp:62-65
            o.update(eqs_options or {})
            capi.check(self.lib.spruce_eqs_ideal2f_options(self.h, int(o["use_sub_cycling"]), int(o["remove_curl_terms"])))
        elif equation_set == "ideal_mhd" and eqs_options:                     # IdealMHD::parseEquationSetConfigs, idealmhd.cpp:12-40
            o = eqs_options
            if "global_viscosity" in o:
                capi.check(self.lib.spruce_eqs_ideal_mhd_options(self.h, float(o["global_viscosity"])))
            if o.get("moc_b_limiting") or o.get("moc_mom_limiting"):
                capi.check(self.lib.spruce_eqs_ideal_mhd_moc_limiting(self.h, int(bool(o.get("moc_b_limiting"))), float(o.get("moc_b_lower_lim", 0.1)),
                                                                      float(o.get("moc_b_upper_lim", 10.0)), int(bool(o.get("moc_mom_limiting"))),
                                                                      float(o.get("moc_mom_lower_lim", 0.1)), float(o.get("moc_mom_upper_lim", 10.0))))
        for name in self.DOMAIN + self.STATE:
            if name in planes:
                self.upload(name, planes[name])
        if setup:
            self.setup()

    # -- plane transfer (EquationSet::grid(name), equationset.cpp:113-121)
    def _local(self, a: np.ndarray) -> np.ndarray:
        a = np.asarray(a, dtype=np.float64)
        if a.shape[0] == self.xdim and self.nx != self.xdim:
            a = a[self.row0:self.row0 + self.nx]
        return np.ascontiguousarray(a)

    def upload(self, name: str, a: np.ndarray):
        a = self._local(a)
        capi.check(self.lib.spruce_grid_upload(self.h, name.encode(), _dp(a), a.size))

    def grid(self, name: str, out: np.ndarray = None) -> np.ndarray:
        """Plane `name` of the primary state (derived variables are evaluated on demand).  `out`: a caller-owned C-contiguous
        float64 (nx, ydim) buffer to fill -- e.g. pinned host memory, which the device copies into at full PCIe speed."""
        if out is None:
            out = np.empty((self.nx, self.ydim))
        elif out.shape != (self.nx, self.ydim) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise capi.SpruceError("out must be a C-contiguous float64 array of shape (%d, %d)" % (self.nx, self.ydim))
        capi.check(self.lib.spruce_grid_download(self.h, name.encode(), _dp(out), out.size))
        return out

    def evolved(self) -> dict:
        return {k: self.grid(k) for k in self.EVOLVED}

    # -- EquationSet
    def setup(self):
        capi.check(self.lib.spruce_eqs_setup(self.h))

    def propagateChanges(self):
        capi.check(self.lib.spruce_eqs_propagate_changes(self.h))

    def computeTimeDerivatives(self) -> np.ndarray:
        k = np.empty((len(self.EVOLVED), self.nx, self.ydim))
        capi.check(self.lib.spruce_eqs_time_derivatives(self.h, _dp(k), k.size))
        return k

    def next_step_size(self) -> float:
        s = C.c_double()
        capi.check(self.lib.spruce_next_step_size(self.h, C.byref(s)))
        return s.value

    # -- time loop (PlasmaDomain::advanceTime / run, evolution.cpp:8-82)
    def advance(self, n_steps: int, max_time: float = -1.0) -> np.ndarray:
        dts = np.zeros(max(n_steps, 1))
        done = C.c_int()
        capi.check(self.lib.spruce_advance(self.h, n_steps, max_time, _dp(dts), C.byref(done)))
        return dts[:done.value]

    def advanceTime(self) -> float:
        return float(self.advance(1)[0])

    @property
    def time(self) -> float:
        t = C.c_double(); it = C.c_int64()
        capi.check(self.lib.spruce_get_time(self.h, C.byref(t), C.byref(it)))
        return t.value

    @property
    def iter(self) -> int:
        t = C.c_double(); it = C.c_int64()
        capi.check(self.lib.spruce_get_time(self.h, C.byref(t), C.byref(it)))
        return it.value

    # -- modules (config block -> device module), call order = execution order (modulehandler.cpp:92-111)
    def set_thermal_conduction(self, *, flux_saturation=False, integrator="euler", epsilon=0.1, dt_subcycle_min=1.0e-4, weakening_factor=1.0):
        capi.check(self.lib.spruce_module_thermal_conduction(self.h, int(flux_saturation), capi.TI[integrator], epsilon, dt_subcycle_min, weakening_factor))

    def set_radiative_losses(self, *, integrator="euler", cutoff_ramp=1.0e3, cutoff_temp=3.0e4, epsilon=0.1, prevent_subcycling=False):
        capi.check(self.lib.spruce_module_radiative_losses(self.h, capi.TI[integrator], cutoff_ramp, cutoff_temp, epsilon, int(prevent_subcycling)))

    def set_ambient_heating_plane(self, heating: np.ndarray):
        a = self._local(heating)
        capi.check(self.lib.spruce_module_ambient_heating(self.h, _dp(a), a.size))

    # -- pointwise solar source terms (post-iterate hooks); keyword names are the reference's config keys
    def set_ambient_heating_sink_plane(self, reduction: np.ndarray):
        """reduction = the plane AmbientHeatingSink::setupModule builds (ambientheatingsink.cpp:27-33)."""
        a = self._local(reduction)
        capi.check(self.lib.spruce_module_ambient_heating_sink(self.h, _dp(a), a.size))

    def set_localized_heating(self, *, start_time, duration, max_heating_rate, stddev_x, stddev_y, center_x, center_y, ramp_time=0.0):
        capi.check(self.lib.spruce_module_localized_heating(self.h, start_time, duration, max_heating_rate, stddev_x, stddev_y, center_x, center_y, ramp_time))

    def set_mass_injection(self, *, start_time, duration, max_injection_rate, stddev_x, stddev_y, center_x, center_y):
        capi.check(self.lib.spruce_module_mass_injection(self.h, start_time, duration, max_injection_rate, stddev_x, stddev_y, center_x, center_y))

    def set_momentum_injection(self, *, start_time, duration, max_accel, stddev_x, stddev_y, center_x, center_y, dir_x, dir_y, template_angle=0.0,
                               oscillatory=False, oscillation_period=1.0):
        capi.check(self.lib.spruce_module_momentum_injection(self.h, start_time, duration, max_accel, stddev_x, stddev_y, center_x, center_y, dir_x, dir_y,
                                                             template_angle, int(oscillatory), oscillation_period))

    def set_div_cleaning(self, *, epsilon=0.1, time_scale=1.0):
        capi.check(self.lib.spruce_module_div_cleaning(self.h, epsilon, time_scale))

    def set_field_heating(self, *, coeff=0.0, current_pow=0.0, b_pow=0.0, n_pow=0.0, roc_pow=0.0, inactive_mode=False):
        capi.check(self.lib.spruce_module_field_heating(self.h, coeff, current_pow, b_pow, n_pow, roc_pow, int(inactive_mode)))

    def set_boundary_outflow(self, pos_x: np.ndarray, pos_y: np.ndarray, *, max_accel, falloff_length, boundary="y_bound_2", falloff_shape="exp", feather_length=0.0,
                             field_aligned_mode=False, dynamic_mode=False, dynamic_time=1.0, dynamic_target_speed=0.0):
        x = np.ascontiguousarray(pos_x, dtype=np.float64); y = np.ascontiguousarray(pos_y, dtype=np.float64)
        capi.check(self.lib.spruce_module_boundary_outflow(self.h, _dp(x), _dp(y), x.size, max_accel, falloff_length,
                                                           {"x_bound_1": 0, "x_bound_2": 1, "y_bound_1": 2, "y_bound_2": 3}[boundary],
                                                           {"exp": 0, "gaussian": 1, "flat": 2}[falloff_shape], feather_length, int(field_aligned_mode), int(dynamic_mode),
                                                           dynamic_time, dynamic_target_speed))

    def boundary_outflow_state(self):
        m, a = C.c_double(), C.c_double()
        capi.check(self.lib.spruce_module_boundary_outflow_state(self.h, C.byref(m), C.byref(a)))
        return m.value, a.value

    def set_anomalous_resistivity(self, pos_x: np.ndarray, pos_y: np.ndarray, *, time_scale=1.0, frobenius_metric_coeff=1.0e50, smoothing_sigma=3.0, safety_factor=1.0,
                                  metric_smoothing=True, time_integrator="euler", template_mode="flood_fill", flood_fill_max_radius=-1.0, flood_fill_argmin_radius=5.0e9,
                                  flood_fill_min_current=-1.0, flood_fill_current_ramp_length=1.0e-5, flood_fill_threshold=1.0, resistivity_model="time_scale",
                                  gradient_correction=False, resistivity_model_params=(0.0, 0.0, 0.0)):
        """AnomalousResistivity with the reference's config keys and defaults (anomalousresistivity.hpp:16-39)."""
        x = np.ascontiguousarray(pos_x, dtype=np.float64); y = np.ascontiguousarray(pos_y, dtype=np.float64)
        mp = list(resistivity_model_params) + [0.0, 0.0, 0.0]
        p = np.array([time_scale, frobenius_metric_coeff, smoothing_sigma, safety_factor, float(metric_smoothing), {"euler": 0, "rk2": 1, "rk4": 2}[time_integrator or "euler"],
                      float(template_mode == "flood_fill"), flood_fill_max_radius, flood_fill_argmin_radius, flood_fill_min_current, flood_fill_current_ramp_length,
                      flood_fill_threshold, {"time_scale": 0.0, "syntelis_19": 1.0, "ys_94": 2.0}[resistivity_model], float(gradient_correction), mp[0], mp[1], mp[2]], dtype=np.float64)
        capi.check(self.lib.spruce_module_anomalous_resistivity(self.h, _dp(x), _dp(y), x.size, _dp(p), p.size))

    def anomalous_resistivity_state(self):
        """(null_i, null_j), sub-cycles of the last step"""
        i, j, n = C.c_int(), C.c_int(), C.c_int()
        capi.check(self.lib.spruce_module_anomalous_resistivity_state(self.h, C.byref(i), C.byref(j), C.byref(n)))
        return (i.value, j.value), n.value

    def set_module_output_to_file(self, module: str, on: bool = True):
        """output_to_file = true of thermal_conduction / radiative_losses: keep the module's diagnostic planes (Module::fileOutput)."""
        capi.check(self.lib.spruce_module_output_to_file(self.h, module.encode(), int(on)))

    def module_output(self, name: str) -> np.ndarray:
        """'thermal_conduction', 'flux_saturation' or 'rad' of the last step."""
        out = np.empty((self.nx, self.ydim))
        capi.check(self.lib.spruce_module_output(self.h, name.encode(), _dp(out), out.size))
        return out

    def set_physical_viscosity(self, coeff_plane: np.ndarray, *, coeff, epsilon=1.0, heating_on=True, force_on=True, gradient_correction=False,
                               integrator="euler", inactive_mode=False):
        """coeff_plane = PhysicalViscosity::constructCoefficientGrid(coeff, ramp_length, buffer_length) (physicalviscosity.cpp:247-267)."""
        a = self._local(coeff_plane)
        capi.check(self.lib.spruce_module_physical_viscosity(self.h, coeff, _dp(a), a.size, epsilon, int(heating_on), int(force_on),
                                                             int(gradient_correction), capi.TI[integrator], int(inactive_mode)))

    def set_module_inactive(self, module: str, on: bool = True):
        """inactive_mode of thermal_conduction / radiative_losses: evaluated for the output and cumulative planes, not applied"""
        capi.check(self.lib.spruce_module_inactive_mode(self.h, module.encode(), int(on)))

    def set_multispecies(self, on: bool = True, **fractions):
        """multispecies_mode = true (plasmadomain.hpp:134-135); fractions: ms_electron_heating_fraction per module name (set after the module is configured)"""
        capi.check(self.lib.spruce_multispecies_mode(self.h, int(on)))
        for module, f in fractions.items():
            capi.check(self.lib.spruce_module_ms_fraction(self.h, module.encode(), float(f)))

    def multispecies_reset(self):
        """the reset the run loop makes after every stored frame (evolution.cpp:36-41)"""
        capi.check(self.lib.spruce_multispecies_reset(self.h))

    def set_eic_thermalization(self):
        capi.check(self.lib.spruce_module_eic_thermalization(self.h))

    def set_viscosity(self, terms, *, hv_integrator="euler", hv_epsilon=1.0, gradient_correction=False):
        """terms: list of dict(opt, strength, var_diff, var_evol, species='i', strength_grid=None) in config order."""
        capi.check(self.lib.spruce_module_viscosity(self.h, capi.TI[hv_integrator], hv_epsilon, int(gradient_correction)))
        for tm in terms:
            sg = tm.get("strength_grid")
            if sg is not None:
                sg = self._local(sg)
            capi.check(self.lib.spruce_module_viscosity_term(self.h, tm["opt"].encode(), tm["strength"], tm["var_diff"].encode(), tm["var_evol"].encode(),
                                                             tm.get("species", "i").encode(), _dp(sg) if sg is not None else None, sg.size if sg is not None else 0))

    def operator(self, op: str, index: int, q: np.ndarray, vel: np.ndarray = None) -> np.ndarray:
        q = self._local(q)
        out = np.empty_like(q)
        v = self._local(vel) if vel is not None else None
        capi.check(self.lib.spruce_operator(self.h, op.encode(), index, _dp(q), _dp(v) if v is not None else None, _dp(out), q.size))
        return out

    def operator2(self, op: str, a: np.ndarray, b: np.ndarray, c: np.ndarray = None) -> np.ndarray:
        """divergence2D(a_x, a_y), curl2D(x, y), transportDivergence2D(quantity, vel_x, vel_y) of PlasmaDomain (plasmadomain.hpp:201-242)."""
        a, b = self._local(a), self._local(b)
        cc = self._local(c) if c is not None else None
        out = np.empty_like(a)
        capi.check(self.lib.spruce_operator2(self.h, op.encode(), _dp(a), _dp(b), _dp(cc) if cc is not None else None, _dp(out), a.size))
        return out

    def subcycles(self, which: str) -> int:
        n = C.c_int()
        capi.check(self.lib.spruce_module_subcycles(self.h, which.encode(), C.byref(n)))
        return n.value

    # -- plumbing
    def synchronize(self):
        capi.check(self.lib.spruce_synchronize(self.h))

    def stream(self) -> int:
        s = C.c_void_p()
        capi.check(self.lib.spruce_stream(self.h, C.byref(s)))
        return s.value or 0

    def launch_count(self) -> int:
        n = C.c_int64()
        capi.check(self.lib.spruce_launch_count(self.h, C.byref(n)))
        return n.value

    def time_stage_kernel(self, reps=10) -> float:
        ms = C.c_float()
        capi.check(self.lib.spruce_time_stage_kernel(self.h, reps, C.byref(ms)))
        return ms.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.spruce_domain_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
