"""torchrun --nproc-per-node N scripts/mgpu_check.py [nx ny steps integrator xbound transport]
N-GPU slab run vs the 1-GPU run of the same problem (rank 0 computes both): planes and step sizes must be identical
bit for bit (min/max reductions are exact and every cell sees the same operands)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spruce_b200 import synthetic  # noqa: E402
from spruce_b200.domain import PlasmaDomain  # noqa: E402
from spruce_b200.multigpu import SlabRunner  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 210
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
integ = sys.argv[4] if len(sys.argv) > 4 else "rk2"
xbound = sys.argv[5] if len(sys.argv) > 5 else "periodic"
transport = sys.argv[6] if len(sys.argv) > 6 else "p2p"
modules = [m for m in (sys.argv[7].split(",") if len(sys.argv) > 7 else []) if m]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def add_modules(dom):
    for m in modules:                       # same order on every domain = execution order
        if m == "tc":
            dom.set_thermal_conduction(flux_saturation=False, integrator="euler", epsilon=0.1, dt_subcycle_min=1.0e-4)
        elif m == "tcsat":
            dom.set_thermal_conduction(flux_saturation=True, integrator="rk2", epsilon=0.1, dt_subcycle_min=1.0e-4)
        elif m == "rl":
            dom.set_radiative_losses(integrator="rk2", cutoff_ramp=1.0e3, cutoff_temp=3.0e4, epsilon=0.1)
        elif m == "pv":
            dom.set_physical_viscosity(np.full((dom.nx, dom.ydim), 3.0e-15), coeff=3.0e-15, epsilon=0.1, integrator="rk2", gradient_correction=True)
        elif m == "av":
            prof = 0.8 * np.exp(-((s["planes"]["pos_y"] - s["planes"]["pos_y"].min()) / 6.0e8) ** 2)     # a boundary-style strength profile (global plane)
            dom.set_viscosity([dict(opt="local", strength=0.5, var_diff="v_x", var_evol="mom_x"), dict(opt="global", strength=3.0, var_diff="v_y", var_evol="mom_y"),
                               dict(opt="boundary", strength=0.8, var_diff="temp", var_evol="thermal_energy", strength_grid=prof)],
                              hv_integrator="rk2", hv_epsilon=1.0, gradient_correction=True)
        elif m in ("2f", "2e"):
            pass
        elif m == "2feic":
            dom.set_eic_thermalization()
        elif m == "ah":
            dom.set_ambient_heating_plane(np.full((dom.nx, dom.ydim), 1.0e-4))
        elif m == "dc":
            dom.set_div_cleaning(epsilon=0.1, time_scale=5.0)
        elif m == "fh":
            dom.set_field_heating(coeff=1.0, current_pow=1.0, b_pow=0.5, n_pow=0.25, roc_pow=0.5)
        elif m == "bo":
            dom.set_boundary_outflow(s["planes"]["pos_x"], s["planes"]["pos_y"], max_accel=2.0e3, falloff_length=6.0e8, boundary="x_bound_2", falloff_shape="exp",
                                     feather_length=3.0e8, field_aligned_mode=True, dynamic_mode=True, dynamic_time=10.0, dynamic_target_speed=2.0e6)
        elif m in ("moc", "mocv"):
            pass                                # a boundary condition, selected below
        elif m == "src":                        # the pointwise solar source terms (templates are built per slab from global cell indices)
            dom.set_ambient_heating_sink_plane(np.full((dom.xdim, dom.ydim), 1.0e-5))
            dom.set_localized_heating(start_time=0.0, duration=5.0, max_heating_rate=1.0e-3, stddev_x=9.0, stddev_y=4.0, center_x=0.5 * nx, center_y=8.0, ramp_time=1.0)
            dom.set_mass_injection(start_time=0.0, duration=10.0, max_injection_rate=1.0e6, stddev_x=7.0, stddev_y=3.0, center_x=0.5 * nx + 3.0, center_y=10.0)
            dom.set_momentum_injection(start_time=0.0, duration=50.0, max_accel=1.0e3, stddev_x=8.0, stddev_y=3.0, center_x=0.5 * nx - 2.0, center_y=9.0, dir_x=1.0, dir_y=0.5,
                                       template_angle=20.0, oscillatory=True, oscillation_period=3.0)
        else:
            raise SystemExit("unknown module " + m)


two_fluid = any(m.startswith("2f") for m in modules)
if two_fluid:
    s = synthetic.ucnp_cloud(nx, ny, drift=2.0e3, bfield=5.0)
    ub = ("open_ucnp", "open_ucnp")
    kw = dict(equation_set="ideal_2F", eqs_options=dict(use_sub_cycling=False), xb=("periodic", "periodic") if xbound == "periodic" else ub, yb=ub, integrator=integ,
              density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1e-30)
elif "2e" in modules:
    # IdealMHD2E on slabs: wall-type sides exercise the exchange of the primary state the boundary passes write
    s = synthetic.two_energy(nx, ny, loop=True)
    kw = dict(equation_set="ideal_mhd_2E", xb=("periodic", "periodic") if xbound == "periodic" else (xbound, "open"), yb=("fixed", "open"), integrator=integ,
              density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
elif "moc" in modules or "mocv" in modules:
    # open_moc sides: x periodic with a y side, or x sides (first / last slab) plus a y side -> corners on the end slabs
    s = synthetic.stratified_loop(nx, ny, bump=0.4)
    kw = dict(xb=("periodic", "periodic") if xbound == "periodic" else ("open_moc", "open_moc"), yb=("open_moc", "fixed") if xbound == "periodic" else ("fixed", "open_moc"),
              integrator=integ, density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6, eqs_options=dict(global_viscosity=0.2 if "mocv" in modules else 0.0))
elif modules:
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    kw = dict(xb=("periodic", "periodic") if xbound == "periodic" else (xbound, "open"), yb=("fixed", "fixed"), integrator=integ)
elif xbound in ("periodic", "periodic2d"):
    s = synthetic.orszag_tang(nx, ny, zfull=(xbound == "periodic"))       # periodic2d: zero z system / external field -> the 2-D kernel instance
    kw = dict(xb=("periodic", "periodic"), yb=("periodic", "periodic"), integrator=integ, density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
else:
    s = synthetic.stratified_loop(nx, ny)
    kw = dict(xb=(xbound, "open"), yb=("reflect", "fixed"), integrator=integ)
run = SlabRunner(s["planes"], s["ion_mass"], s["adiabatic_index"], rank=rank, world=world, device=local, transport=transport, **kw)
add_modules(run.dom)
run.step(steps)
run.dom.synchronize()
got = {v: run.gather(v) for v in (run.dom.EVOLVED + ["dt"])}
t_slab = run.dom.time
ok = True
if rank == 0:
    one = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], device=local, **kw)
    add_modules(one)
    dts = one.advance(steps)
    for m, key in (("tc", "thermal_conduction"), ("tcsat", "thermal_conduction"), ("rl", "radiative_losses"), ("pv", "physical_viscosity")):
        if m in modules:
            a, b = run.dom.subcycles(key), one.subcycles(key)
            print("subcycles", key, a, b)
            ok &= (a == b)
    for v, a in got.items():
        b = one.grid(v)
        same = bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))
        ok &= same
        if not same:
            bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
            print("MISMATCH", v, len(bad), bad[:5].tolist())
    ok &= (t_slab == one.time)
    print("mgpu_check world=%d %s %dx%d %s x=%s modules=%s steps=%d : %s (t=%r vs %r)" % (world, transport, nx, ny, integ, xbound, ",".join(modules) or "-", steps, "IDENTICAL" if ok else "DIFFERENT", t_slab, one.time))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
