#!/usr/bin/env python
"""Throughput of the secondary paths on one GPU (device-timed): two-fluid RK2 step, MHD + thermal conduction, MHD + physical viscosity -- each as shipped
and (key suffix " [general instances]") with SPRUCE_FAST_INTERIOR=0, i.e. the wrapping, range-testing stencil instances for every cell; saturated conduction also
with SPRUCE_TC_TWO_PASS=0 (" [five-point coefficient]").
usage: module_perf.py [size]"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from spruce_b200 import synthetic
from spruce_b200.domain import PlasmaDomain

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
out = {}
T0 = time.perf_counter()
BUDGET_S = float(os.environ.get("SPRUCE_PERF_BUDGET_S", "105"))      # the caller's limit is 150 s for the whole process: later entries are skipped rather than lost with all the others
def late():
    return time.perf_counter() - T0 > BUDGET_S

def timed(dom, steps):
    dom.advance(3)
    st = torch.cuda.ExternalStream(dom.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    dom.advance(steps)
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

s = synthetic.ucnp_cloud(n + 1, n + 1, drift=2.0e3, bfield=5.0)
for fast in ("1", "0"):
    os.environ["SPRUCE_FAST_INTERIOR"] = fast                      # read by spruce_domain_create
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], equation_set="ideal_2F", eqs_options=dict(use_sub_cycling=False), xb=("open_ucnp",) * 2, yb=("open_ucnp",) * 2,
                     integrator="rk2", density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1e-30)
    d.set_eic_thermalization()
    ms = timed(d, 10)
    out["ideal_2F+eic rk2 %d^2%s" % (n + 1, "" if fast == "1" else " [general instances]")] = dict(ms_per_step=ms, cell_updates_per_s=(n + 1) ** 2 / ms * 1e3, hbm_frac=(n + 1) ** 2 * 624 / (ms * 1e-3) / 6551.4e9)
    d.close()

def five_point():
    # saturated conduction with the coefficient evaluated at five points per cell (SPRUCE_TC_TWO_PASS=0) next to the two-pass form timed above
    try:
        if late():
            raise RuntimeError("skipped: time budget of this script")
        os.environ["SPRUCE_FAST_INTERIOR"] = "1"; os.environ["SPRUCE_TC_TWO_PASS"] = "0"
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **KW)
        d.set_thermal_conduction(flux_saturation=True, integrator="rk2", epsilon=0.1, dt_subcycle_min=1.0e-4)
        ms = timed(d, 10)
        out["mhd + thermal_conduction (saturated, rk2) %d^2 [five-point coefficient]" % n] = dict(ms_per_step=ms, cell_updates_per_s=n * n / ms * 1e3, subcycles_last_step=d.subcycles("thermal_conduction"))
        d.close()
    except Exception as e:                                                 # a side measurement: the others stand
        out["five-point coefficient"] = {"error": repr(e)[:200]}
    os.environ.pop("SPRUCE_TC_TWO_PASS", None)

KW = dict(xb=("periodic", "periodic"), yb=("periodic", "periodic"), integrator="rk2", epsilon=0.2, density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
s = synthetic.orszag_tang(n, n, temp_mod=0.1)
for name, setup in (("mhd only", lambda d: None),
                    ("mhd + thermal_conduction (unsaturated, euler)", lambda d: d.set_thermal_conduction(flux_saturation=False, integrator="euler", epsilon=0.1, dt_subcycle_min=1.0e-4)),
                    ("mhd + thermal_conduction (saturated, rk2)", lambda d: d.set_thermal_conduction(flux_saturation=True, integrator="rk2", epsilon=0.1, dt_subcycle_min=1.0e-4)),
                    ("mhd + physical_viscosity (euler)", lambda d: d.set_physical_viscosity(np.full((n, n), 4.0e-18), coeff=4.0e-18, epsilon=0.2))):
    for fast in ("1", "0"):
        if fast == "0" and name == "mhd only":
            continue
        if late():
            out["%s %d^2%s" % (name, n, "" if fast == "1" else " [general instances]")] = {"skipped": "time budget of this script"}
            continue
        os.environ["SPRUCE_FAST_INTERIOR"] = fast
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **KW)
        setup(d)
        ms = timed(d, 10)
        e = dict(ms_per_step=ms, cell_updates_per_s=n * n / ms * 1e3)
        for k in ("thermal_conduction", "physical_viscosity"):
            if k.split("_")[0] in name.replace("thermal", "thermal_conduction").replace("physical", "physical_viscosity") and k in name:
                e["subcycles_last_step"] = d.subcycles(k)
        out["%s %d^2%s" % (name, n, "" if fast == "1" else " [general instances]")] = e
        d.close()
        if fast == "1" and "saturated" in name:
            five_point()
print(json.dumps(out, indent=1))
