"""The device-resident sub-cycle plan (SPRUCE_DEVICE_SUBCYCLES=1, read at spruce_domain_create; DESIGN.md section 4 "no host round trip with modules"): with
thermal_conduction / radiative_losses / ambient_heating the step size and both sub-cycle counts stay on the device and no step waits for the host.  The same run with the
plan on and off must give the same step sizes, the same sub-cycle counts and the same evolved and output planes BIT FOR BIT -- the host-driven form is the one the strict
oracle tests validated.  Also: a budget that is too small (SPRUCE_TC_BUDGET=1) must stop, re-plan and still arrive at the same bits; a run that reaches max_time inside a
batch must stop where the host-driven one stops; and an euler run that stops inside a batch must keep naming the right set as its primary state.  The host-executed form of
the first two checks is in tests/test_capi_hooks_emulated.py (test_planned_subcycles_*).  Written after the round's GPU budget was spent: non-strict, first executed by the
round-end suite."""
import os

import numpy as np
import pytest

from golden_util import same_bits
from spruce_b200 import synthetic

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(reason="written after round 2's GPU budget was spent: first executed by the round-end suite", strict=False)]

NX, NY = 70, 90
KW = dict(xb=("periodic", "periodic"), yb=("fixed", "open"))


class env:
    def __init__(self, **kv):
        self.kv, self.old = kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def solar_domain(mods, integrator="rk2", outputs=()):
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.stratified_loop(NX, NY, bump=0.5)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], integrator=integrator, **KW)
    for name, mk in mods:
        if name == "ambient_heating":
            mask = np.zeros((NX, NY)); mask[:, 2:NY - 2] = 1.0
            d.set_ambient_heating_plane(mask * mk["heating_rate"])
        else:
            getattr(d, "set_" + name)(**mk)
    for m in outputs:
        d.set_module_output_to_file(m)
    return d


def snapshot(d, names, outputs):
    out = {v: d.grid(v).copy() for v in d.EVOLVED + ["dt"]}
    for m in names:
        if m in ("thermal_conduction", "radiative_losses"):
            out["n_" + m] = d.subcycles(m)
    for o in outputs:
        out["out_" + o] = d.module_output(o).copy()
    return out


def equal(a, b):
    assert a.keys() == b.keys()
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert same_bits(a[k], b[k]), k
        else:
            assert a[k] == b[k], (k, a[k], b[k])


SOLAR = [("thermal_conduction", dict(flux_saturation=True, integrator="rk2")), ("radiative_losses", dict(integrator="rk2")), ("ambient_heating", dict(heating_rate=1.0e-4))]


@pytest.mark.parametrize("mods,outputs", [
    (SOLAR, ("thermal_conduction", "radiative_losses")),
    ([("thermal_conduction", dict(flux_saturation=False, integrator="euler"))], ()),
    ([("thermal_conduction", dict(flux_saturation=True, integrator="rk4")), ("radiative_losses", dict(integrator="rk4", prevent_subcycling=True))], ()),
    ([("radiative_losses", dict(integrator="euler")), ("ambient_heating", dict(heating_rate=1.0e-4))], ("radiative_losses",)),
])
@pytest.mark.parametrize("integrator", ["rk2", "euler"])
def test_device_plan_equals_host_driven_run(mods, outputs, integrator):
    names = [m[0] for m in mods]
    planes = {"thermal_conduction": ("thermal_conduction", "flux_saturation"), "radiative_losses": ("rad",)}
    outs = [p for m in outputs for p in planes[m] if not (p == "flux_saturation" and not dict(mods)["thermal_conduction"]["flux_saturation"])]
    res = []
    for on in ("0", "1"):
        with env(SPRUCE_DEVICE_SUBCYCLES=on):
            d = solar_domain(mods, integrator, outputs)
        assert d.subcycles("device_plan") == int(on)
        dts = list(d.advance(1)) + list(d.advance(5))               # a single step, then a batch no step of which may wait for the host
        res.append((dts, snapshot(d, names, outs)))
        d.close()
    assert [x.hex() for x in res[0][0]] == [x.hex() for x in res[1][0]] and len(res[0][0]) == 6
    equal(res[0][1], res[1][1])


def test_device_plan_with_too_small_a_budget_replans_and_arrives_at_the_same_bits():
    names = [m[0] for m in SOLAR]
    with env(SPRUCE_DEVICE_SUBCYCLES="0"):
        h = solar_domain(SOLAR)
    dh = list(h.advance(6))
    want = snapshot(h, names, ())
    assert want["n_thermal_conduction"] > 1, "the case must need more than one conduction sub-cycle"
    h.close()
    with env(SPRUCE_DEVICE_SUBCYCLES="1", SPRUCE_TC_BUDGET="1"):
        d = solar_domain(SOLAR)
    assert d.subcycles("device_plan_budget") == 1
    dd = list(d.advance(6))
    assert d.subcycles("device_plan_replans") >= 1 and d.subcycles("device_plan_budget") >= want["n_thermal_conduction"]
    assert [x.hex() for x in dd] == [x.hex() for x in dh]
    equal(snapshot(d, names, ()), want)
    d.close()


@pytest.mark.parametrize("integrator", ["rk2", "euler"])
def test_device_plan_stops_at_max_time_inside_a_batch(integrator):
    names = [m[0] for m in SOLAR]
    res = []
    for on in ("0", "1"):
        with env(SPRUCE_DEVICE_SUBCYCLES=on):
            d = solar_domain(SOLAR, integrator, ("thermal_conduction", "radiative_losses"))
        first = list(d.advance(4))
        tmax = float(np.sum(first)) + 2.5 * first[-1]
        rest = list(d.advance(8, max_time=tmax))                     # stops after about three of the eight steps; 8 - 3 is odd: the euler run exchanges its sets once too often
        assert 0 < len(rest) < 8
        more = list(d.advance(2))                                    # and goes on from the right state
        res.append((first + rest + more, snapshot(d, names, ("thermal_conduction", "flux_saturation", "rad")), d.iter))
        d.close()
    assert [x.hex() for x in res[0][0]] == [x.hex() for x in res[1][0]] and res[0][2] == res[1][2]
    equal(res[0][1], res[1][1])


@pytest.mark.parametrize("zfull", [False, True])
def test_euler_batch_that_stops_at_max_time_keeps_the_primary_state(zfull):
    """plain ideal MHD, euler: steps enqueued after max_time are no-ops on the device, but each still exchanges the roles of the two plane sets on the host; after an odd number
    of them the host must not name the stale set as the primary state (spruce_advance counts the exchanges and undoes the odd one)"""
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.orszag_tang(96, 80, zfull=zfull)
    kw = dict(density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30, integrator="euler")
    a = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    da = [a.advanceTime() for _ in range(7)]
    for n_batch in (8, 9):                                           # 3 resp. 4 no-op steps after the 5 real ones
        b = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
        tmax = float(np.sum(da[:4])) + 0.5 * da[4]
        db = list(b.advance(n_batch, max_time=tmax))
        assert [x.hex() for x in db] == [x.hex() for x in da[:5]] and b.iter == 5
        db += list(b.advance(2))
        assert [x.hex() for x in db] == [x.hex() for x in da]
        for v in PlasmaDomain.EVOLVED:
            assert same_bits(a.grid(v), b.grid(v)), (n_batch, v)
        b.close()
    a.close()


@pytest.mark.parametrize("args", [("128", "96", "4", "rk2", "periodic", "p2p", "tcsat,rl,ah"), ("96", "80", "3", "euler", "reflect", "p2p", "tc,rl")])
@pytest.mark.parametrize("budget", ["1", "6"])
def test_device_plan_on_two_slabs_equals_one_gpu(args, budget):
    """the plan-driven module steps on 2 slabs (the reduction words all-gathered over peer memory before the planning thread, skipped sub-cycle stages still taking part in the
    halo exchanges, a stop for a larger budget taken by both ranks at the same step) == the plan-driven run on one GPU, bit for bit (scripts/mgpu_check.py)"""
    import subprocess
    import sys
    import torch
    from pathlib import Path
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29533",
                        str(root / "scripts" / "mgpu_check.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300,
                       env=dict(os.environ, SPRUCE_DEVICE_SUBCYCLES="1", SPRUCE_TC_BUDGET=budget))
    out = r.stdout.decode()
    assert r.returncode == 0 and "IDENTICAL" in out, out[-3000:]
