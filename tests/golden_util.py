"""Shared helpers for the parity tests: load a golden fixture (tests/golden/*.npz, produced by
tests/golden/make_golden.py from the unmodified reference binary) and describe it."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
OUT_VARS = ["rho", "temp", "thermal_energy", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "dt"]
EVOLVED = ["rho", "mom_x", "mom_y", "mom_z", "thermal_energy", "bi_x", "bi_y", "bi_z"]


EVOLVED_2F = ["i_rho", "e_rho", "i_mom_x", "i_mom_y", "e_mom_x", "e_mom_y", "i_thermal_energy", "e_thermal_energy",
              "E_x", "E_y", "E_z", "bi_x", "bi_y", "bi_z"]


EXTENDED = ("moc_", "sm_", "ar_", "e2_")      # open_moc boundary, small solar modules, anomalous resistivity, IdealMHD2E


def cases(prefixes=None, two_fluid=False, oracle_only=False, extended=False):
    """Fixture names.  tf_*: the two-fluid equation set.  extended: the fixtures of the 8f rows (separate tests: their device paths are
    not GPU-validated yet); the default lists leave them out."""
    names = sorted(p.stem for p in GOLDEN.glob("*.npz"))
    names = [n for n in names if n.startswith("tf_") == two_fluid]
    names = [n for n in names if n.startswith(EXTENDED) == extended]
    if oracle_only:
        names = [n for n in names if "_pv_" not in n]
    if prefixes:
        names = [n for n in names if any(n.startswith(p) for p in prefixes)]
    return names


class Golden:
    def __init__(self, name):
        z = np.load(GOLDEN / (name + ".npz"))
        self.name = name
        self.desc = json.loads(str(z["desc"]))
        self.planes = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
        self.ion_mass = float(z["ion_mass"])
        self.adiabatic_index = float(z["adiabatic_index"])
        self.steps = z["steps"]
        self.n_steps = self.desc["n_steps"]
        self.keep = self.desc["keep"]
        self.out_vars = self.desc.get("out_vars", OUT_VARS)
        self.frames = {fi: {v: z["f%d_%s" % (fi, v)] for v in self.out_vars} for fi in self.keep}
        self.module_planes = {fi: {k.split("_mod_", 1)[1]: z[k] for k in z.files if k.startswith("f%d_mod_" % fi)} for fi in self.keep}
        cfg = self.desc["config"]
        self.cfg = cfg
        self.kw = dict(xb=tuple(cfg["xb"]), yb=tuple(cfg["yb"]), integrator=cfg["integrator"],
                       epsilon=cfg.get("epsilon", 0.2), density_min=cfg["density_min"], temp_min=cfg["temp_min"],
                       thermal_energy_min=cfg["thermal_energy_min"],
                       open_strength=cfg.get("open_strength", 1.0), open_decay=cfg.get("open_decay", 0.5))
        self.modules = [(m[0], dict(m[1])) for m in cfg.get("modules", [])]
        self.equation_set = cfg.get("eqs", "ideal_mhd")
        self.eqs_options = {k: (v == "true") for k, v in cfg.get("eqs_block", [])}
        self.eqs_raw = {k: v for k, v in cfg.get("eqs_block", [])}

    def viscous_subcycle_counts(self):
        """'... , N Subcycle(s)' of the physical_viscosity message (physicalviscosity.cpp:283)."""
        out = []
        for ln in self.desc["subcycle_log"]:
            for part in ln.split("|"):
                if "Subcycle(s)" in part:
                    out.append(int(part.split("Subcycle(s)")[0].split(",")[-1].split()[-1]))
        return out

    def subcycle_counts(self, label):
        """Per-iteration sub-cycle counts parsed from the reference's stdout lines ('Thermal Subcycles: N')."""
        out = []
        for ln in self.desc["subcycle_log"]:
            for part in ln.split("|"):
                if part.startswith(label):
                    out.append(int(part.split(":")[1].split()[0]))
        return out


def module_kwargs(name, kv):
    """Translate a .config module block into keyword arguments shared by the oracle and the product API."""
    b = lambda s: s == "true"
    if name == "thermal_conduction":
        return dict(flux_saturation=b(kv.get("flux_saturation", "false")), integrator=kv.get("time_integrator", "euler"),
                    epsilon=float(kv["epsilon"]), dt_subcycle_min=float(kv["dt_subcycle_min"]),
                    weakening_factor=float(kv.get("weakening_factor", "1.0")))
    if name == "radiative_losses":
        return dict(integrator=kv.get("time_integrator", "euler"), cutoff_ramp=float(kv["cutoff_ramp"]),
                    cutoff_temp=float(kv["cutoff_temp"]), epsilon=float(kv["epsilon"]),
                    prevent_subcycling=b(kv.get("prevent_subcycling", "false")))
    if name == "ambient_heating":
        return dict(heating_rate=float(kv.get("heating_rate", "0.0")), exp_mode=b(kv.get("exp_mode", "false")),
                    exp_base_heating_rate=float(kv.get("exp_base_heating_rate", "0.0")),
                    exp_scale_height=float(kv.get("exp_scale_height", "1.0")),
                    split_exp_mode=b(kv.get("split_exp_mode", "false")),
                    split_exp_scale_height=float(kv.get("split_exp_scale_height", "1.0")),
                    split_exp_start_height=float(kv.get("split_exp_start_height", "0.0")))
    if name == "physical_viscosity":
        return dict(coeff=float(kv.get("coeff", "0.0")), ramp_length=float(kv.get("ramp_length", "0.0")), buffer_length=float(kv.get("buffer_length", "0.0")),
                    epsilon=float(kv.get("epsilon", "1.0")), heating_on=b(kv.get("heating_on", "true")), force_on=b(kv.get("force_on", "true")),
                    gradient_correction=b(kv.get("gradient_correction", "false")), integrator=kv.get("time_integrator", "euler") or "euler",
                    inactive_mode=b(kv.get("inactive_mode", "false")))
    if name == "artificial_viscosity":
        n = len(kv["visc_opt"].split(","))
        cols = {k: kv[k].split(",") for k in ("visc_opt", "visc_strength", "visc_vars_to_diff", "visc_vars_to_evol", "visc_length", "visc_species")}
        terms = [dict(opt=cols["visc_opt"][i], strength=float(cols["visc_strength"][i]), var_diff=cols["visc_vars_to_diff"][i],
                      var_evol=cols["visc_vars_to_evol"][i], length=float(cols["visc_length"][i]), species=cols["visc_species"][i]) for i in range(n)]
        return dict(terms=terms, hv_integrator=kv.get("hv_time_integrator", "euler"), hv_epsilon=float(kv.get("hv_epsilon", "1.0")),
                    gradient_correction=b(kv.get("gradient_correction", "false")))
    raise KeyError(name)


MS_PLANES = ("cumulative_electron_heating", "cumulative_ion_heating", "cumulative_joule_heating")


def multispecies_fractions(modules):
    """ms_electron_heating_fraction of every module block that sets it (the others keep the reference's defaults)"""
    return {name: float(kv["ms_electron_heating_fraction"]) for name, kv in modules if "ms_electron_heating_fraction" in kv}


def viscosity_plane_request(modules, pname):
    """'<evolved>_dqdt' / '_lap' / '_str' / '_dt' of Viscosity::fileOutput (viscosity.cpp:103-107, 351-376) -> (which, term index), or None for other names"""
    for suffix in ("dqdt", "lap", "str", "dt"):
        if pname.endswith("_" + suffix):
            for name, kv in modules:
                if name == "artificial_viscosity":
                    evol = kv["visc_vars_to_evol"].split(",")
                    if pname[:-len(suffix) - 1] in evol:
                        return suffix, evol.index(pname[:-len(suffix) - 1])
    return None


def boundary_viscosity_profile(pos_x, pos_y, strength, length, shape="gaussian"):
    """Viscosity::getBoundaryViscosity (reference source/modules/viscosity.cpp:278-325) for its four shapes -- gaussian, exp, exp_elliptical,
    gaussian_elliptical --, with the host libm: a static profile the reference also builds on the host."""
    import math
    ex = np.vectorize(math.exp)
    x_min, x_max, y_min, y_max = pos_x.min(), pos_x.max(), pos_y.min(), pos_y.max()
    r = np.zeros_like(pos_x)
    if shape == "gaussian":
        for a in ((pos_x - x_min), (pos_x - x_max), (pos_y - y_max), (pos_y - y_min)):
            q = a / length
            r = r + ex((q * q) * -2.3) * strength
    elif shape == "exp":
        for a, c in (((pos_x - x_min), -2.3), ((pos_x - x_max), 2.3), ((pos_y - y_max), 2.3), ((pos_y - y_min), -2.3)):
            r = r + ex((a * c) / length) * strength
    else:
        kind, ell = shape.split("_")
        assert ell == "elliptical" and kind in ("exp", "gaussian")
        xc, yc = 0.5 * (x_min + x_max), 0.5 * (y_min + y_max)
        s = ((pos_x - xc) * (pos_x - xc)) / math.pow(x_max - xc, 2.0) + ((pos_y - yc) * (pos_y - yc)) / math.pow(y_max - yc, 2.0)      # :309-310
        s_length = length / min(x_max - xc, y_max - yc)
        m = np.minimum(s - 1.0, 0.0)
        if kind == "exp":
            r = ex((m * 2.3) / s_length) * strength
        else:
            q = m / s_length
            r = ex((q * q) * -2.3) * strength
    return np.where(strength < r, strength, r)


def viscosity_terms_with_profiles(planes, terms, shape="gaussian"):
    out = []
    for tm in terms:
        tm = dict(tm)
        if tm["opt"] in ("boundary", "boundary_global"):
            tm["strength_grid"] = boundary_viscosity_profile(planes["pos_x"], planes["pos_y"], tm["strength"], tm["length"], shape)
        out.append(tm)
    return out


def same_bits(a, b):
    """Bit-for-bit equality up to the sign of zero (and NaN == NaN)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def mismatch(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    n = int(bad.sum())
    if n == 0:
        return "identical"
    with np.errstate(all="ignore"):
        rel = np.nanmax(np.abs(a - b)) / max(np.nanmax(np.abs(b)), 1e-300)
    idx = np.argwhere(bad)[:5].tolist()
    return "%d/%d cells differ, rel Linf %.3e, first at %s" % (n, a.size, rel, idx)


def physical_viscosity_coefficient(planes, coeff, ramp_length):
    """PhysicalViscosity::constructCoefficientGrid (reference source/modules/solar/physicalviscosity.cpp:247-267), on the host as the
    reference builds it (buffer_length is parsed but unused there)."""
    x, y = planes["pos_x"], planes["pos_y"]
    if ramp_length == 0.0:
        return coeff * np.ones_like(x)
    x_max, y_max, x_min, y_min = x.max(), y.max(), x.min(), y.min()
    xc, yc = 0.5 * (x_min + x_max), 0.5 * (y_min + y_max)
    s = (x - xc) ** 2 / (x_max - xc) ** 2.0 + (y - yc) ** 2 / (y_max - yc) ** 2.0
    s_length = ramp_length / min(x_max - xc, y_max - yc)
    import math
    arg = -2.3 * (np.maximum(s + 2.0 * s_length - 1.0, 0.0) / s_length) ** 2
    ex = np.vectorize(math.exp, otypes=[np.float64])(arg)       # libm's exp, as the reference's Grid::exp (numpy's SIMD exp differs in the last bit)
    res = 1.0 / 0.99 * np.maximum(ex - 0.01, 0.0)
    return coeff * res


def small_module_kwargs(name, kv):
    """A .config block of one of the small solar modules -> (oracle kwargs, product kwargs); keys are the reference's config keys."""
    def val(v):
        return 1.0 if v == "true" else 0.0 if v == "false" else float(v)
    if name == "boundary_outflow":
        B = dict(x_bound_1=0, x_bound_2=1, y_bound_1=2, y_bound_2=3); S = dict(exp=0, gaussian=1, flat=2)
        ora = {k: (B[v] if k == "boundary" else S[v] if k == "falloff_shape" else val(v)) for k, v in kv.items()}
        prod = {k: (v if k in ("boundary", "falloff_shape") else (v == "true") if v in ("true", "false") else float(v)) for k, v in kv.items()}
        return ora, prod
    kv = {k: v for k, v in kv.items() if k != "ms_electron_heating_fraction"}       # multispecies_mode: handed over separately (multispecies_fractions)
    ora = {k: val(v) for k, v in kv.items()}
    prod = {k: ((v == "true") if v in ("true", "false") else float(v)) for k, v in kv.items()}
    return ora, prod


def sink_reduction_plane(planes, kw, xb, yb):
    """AmbientHeatingSink::setupModule (ambientheatingsink.cpp:27-33), host libm -- a static plane the reference also builds on the host."""
    import math
    X, Y = planes["pos_x"], planes["pos_y"]
    nx, ny = X.shape
    mask = np.zeros((nx, ny))
    il = lambda b: 0 if b == "periodic" else 2
    ih = lambda b, n: n if b == "periodic" else n - 2
    mask[il(xb[0]):ih(xb[1], nx), il(yb[0]):ih(yb[1], ny)] = 1.0
    if kw.get("exp_mode"):
        ex = np.vectorize(math.exp)((-1.0 * Y) / kw["exp_scale_height"])
        q = (X - kw["center_x"]) / kw["half_width"]
        return ((mask * kw["exp_base_heating_rate"]) * ex) * np.maximum(1.0 - q * q, 0.0)
    return mask * kw.get("heating_rate", 0.0)
