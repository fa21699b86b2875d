"""The FAST stencil instances on the device (SPRUCE_FAST_INTERIOR, read at spruce_domain_create): the same run with them on (default) and off must
give the same step sizes and the same evolved planes BIT FOR BIT -- thermal conduction (unsaturated / saturated), physical viscosity, the two-fluid set with EIC
thermalization -- on wall and periodic sides, at sizes where most cells are deep-interior ones.  The general instances are the ones the strict oracle tests validated
in round 2; the host-executed form of this check is in tests/test_capi_hooks_emulated.py, tests/test_ideal2f_kernels_emulated.py and
tests/test_stencil_fast_instances_host.py.  Written after the round's GPU budget was spent: non-strict, first executed by the round-end suite."""
import os

import numpy as np
import pytest

from golden_util import same_bits
from spruce_b200 import synthetic

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(reason="written after round 2's GPU budget was spent: first executed by the round-end suite", strict=False)]

FLOORS = dict(density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)


def run_both(make, steps, switch="SPRUCE_FAST_INTERIOR"):
    from spruce_b200.domain import PlasmaDomain  # noqa: F401
    out = []
    for fast in ("1", "0"):
        old = os.environ.get(switch)
        os.environ[switch] = fast
        try:
            d = make()
        finally:
            if old is None:
                os.environ.pop(switch, None)
            else:
                os.environ[switch] = old
        dts = np.asarray(d.advance(steps))
        planes = {v: d.grid(v).copy() for v in d.EVOLVED}
        d.close()
        out.append((dts, planes))
    (da, pa), (db, pb) = out
    assert [x.hex() for x in da] == [x.hex() for x in db]
    for v in pa:
        assert same_bits(pa[v], pb[v]), v
    return out[0]


@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("fixed", "open")), (("reflect", "open"), ("periodic", "periodic")), (("fixed", "fixed"), ("fixed", "fixed"))])
@pytest.mark.parametrize("sat,integ", [(False, "euler"), (True, "rk2"), (True, "rk4")])
def test_thermal_conduction_fast_equals_general(xb, yb, sat, integ):
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.stratified_loop(150, 140, bump=0.5)

    def make():
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
        d.set_thermal_conduction(flux_saturation=sat, integrator=integ, epsilon=0.1, dt_subcycle_min=1.0e-4)
        return d
    dts, planes = run_both(make, 3)
    assert np.all(np.isfinite(planes["thermal_energy"]))


@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("fixed", "open")), (("reflect", "open"), ("periodic", "periodic")), (("fixed", "fixed"), ("fixed", "fixed"))])
@pytest.mark.parametrize("integ", ["euler", "rk2", "rk4"])
def test_saturated_conduction_two_pass_equals_the_five_point_form(xb, yb, integ):
    """SPRUCE_TC_TWO_PASS (default 1): the saturation coefficient and raw flux of every cell written once by k_tc_coef and differentiated as a plane, against the form that
    evaluates the coefficient at five points per cell (the one the round-2 strict tests validated): same step sizes, same planes, bit for bit"""
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.stratified_loop(150, 140, bump=0.5)

    def make():
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
        d.set_thermal_conduction(flux_saturation=True, integrator=integ, epsilon=0.1, dt_subcycle_min=1.0e-4)
        d.set_module_output_to_file("thermal_conduction")
        return d
    dts, planes = run_both(make, 3, switch="SPRUCE_TC_TWO_PASS")
    assert np.all(np.isfinite(planes["thermal_energy"]))


@pytest.mark.parametrize("xb,yb", [(("periodic", "periodic"), ("fixed", "open")), (("reflect", "open"), ("fixed", "fixed"))])
@pytest.mark.parametrize("gc,integ", [(False, "euler"), (True, "rk2")])
def test_physical_viscosity_fast_equals_general(xb, yb, gc, integ):
    from spruce_b200.domain import PlasmaDomain
    nx, ny = 140, 130
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    cg = np.full((nx, ny), 1.0e-14) * (1.0 + 0.5 * np.sin(np.arange(ny) / 9.0))[None, :]

    def make():
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator="rk2", **FLOORS)
        d.set_physical_viscosity(cg, coeff=1.0e-14, epsilon=0.1, gradient_correction=gc, integrator=integ)
        return d
    run_both(make, 3)


@pytest.mark.parametrize("xb,yb,integ,eic", [(("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp"), "rk2", True), (("periodic", "periodic"), ("fixed", "reflect"), "rk4", False),
                                            (("periodic", "periodic"), ("periodic", "periodic"), "euler", True)])
def test_two_fluid_fast_equals_general(xb, yb, integ, eic):
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.ucnp_cloud(131, 125, drift=2.0e3, bfield=5.0)

    def make():
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], equation_set="ideal_2F", eqs_options=dict(use_sub_cycling=False), xb=xb, yb=yb, integrator=integ,
                         density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1e-30)
        if eic:
            d.set_eic_thermalization()
        return d
    run_both(make, 4)
