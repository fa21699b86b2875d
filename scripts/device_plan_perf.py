#!/usr/bin/env python
"""What the device-resident sub-cycle plan buys (SPRUCE_DEVICE_SUBCYCLES): the solar module set (thermal_conduction saturated rk2 + radiative_losses rk2 +
ambient_heating) on a gravity-stratified loop, batches of steps with the plan off (three host waits per step) and on (one per batch), device-timed; also whether both runs
arrived at the same step sizes and thermal energy bit for bit.  usage: device_plan_perf.py [size ...]"""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from spruce_b200 import synthetic
from spruce_b200.domain import PlasmaDomain

sizes = [int(a) for a in sys.argv[1:]] or [256, 1024]
out = {}
for n in sizes:
    s = synthetic.stratified_loop(n, n, bump=0.5)
    res = {}
    for on in ("0", "1"):
        os.environ["SPRUCE_DEVICE_SUBCYCLES"] = on                      # read by spruce_domain_create
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=("periodic", "periodic"), yb=("fixed", "open"), integrator="rk2")
        d.set_thermal_conduction(flux_saturation=True, integrator="rk2", epsilon=0.1, dt_subcycle_min=1.0e-4)
        d.set_radiative_losses(integrator="rk2")
        mask = np.zeros((n, n)); mask[:, 2:n - 2] = 1.0
        d.set_ambient_heating_plane(mask * 1.0e-4)
        dts = list(d.advance(5))
        st = torch.cuda.ExternalStream(d.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 40
        torch.cuda.synchronize(); e0.record(st)
        dts += list(d.advance(steps))
        e1.record(st); torch.cuda.synchronize()
        res[on] = dict(ms_per_step=e0.elapsed_time(e1) / steps, dts=[x.hex() for x in dts], e=d.grid("thermal_energy").copy(), tc=d.subcycles("thermal_conduction"),
                       rl=d.subcycles("radiative_losses"), plan=d.subcycles("device_plan"), budget=d.subcycles("device_plan_budget"), replans=d.subcycles("device_plan_replans"))
        d.close()
    a, b = res["0"], res["1"]
    out["solar modules %d^2" % n] = dict(ms_per_step_host_driven=a["ms_per_step"], ms_per_step_device_plan=b["ms_per_step"], device_plan_in_use=b["plan"],
                                         same_step_sizes=a["dts"] == b["dts"], same_thermal_energy_bits=bool(np.array_equal(a["e"].view(np.uint64), b["e"].view(np.uint64))),
                                         subcycles_last_step=dict(thermal_conduction=[a["tc"], b["tc"]], radiative_losses=[a["rl"], b["rl"]]),
                                         budget_after=b["budget"], replans=b["replans"])
os.environ.pop("SPRUCE_DEVICE_SUBCYCLES", None)
print(json.dumps(out, indent=1))
