"""Pins the oracle (oracle/spruce_oracle.c, the CPU restatement) against outputs of the UNMODIFIED reference
binary: every fixture in tests/golden/ must be reproduced bit-for-bit -- the step-size history (reference
source/mhd/evolution.cpp:62) and every evolved plane + temp + dt after the recorded iterations."""
import numpy as np
import pytest

from golden_util import (Golden, MS_PLANES, OUT_VARS, cases, mismatch, module_kwargs, multispecies_fractions, physical_viscosity_coefficient, same_bits, small_module_kwargs, viscosity_plane_request,
                         viscosity_terms_with_profiles)
from oracle.oracle import Oracle


def make_oracle(g: Golden) -> Oracle:
    o = Oracle(g.planes, g.ion_mass, g.adiabatic_index, **g.kw)
    for name, kv in g.modules:
        kw = module_kwargs(name, kv)
        if name == "artificial_viscosity":
            o.set_viscosity(viscosity_terms_with_profiles(g.planes, kw.pop("terms")), **kw)
        elif name == "physical_viscosity":
            ramp = kw.pop("ramp_length"); kw.pop("buffer_length")
            o.set_physical_viscosity(physical_viscosity_coefficient(g.planes, kw["coeff"], ramp), **kw)
        else:
            getattr(o, "set_" + name)(**kw)
        if name in ("thermal_conduction", "radiative_losses") and kv.get("inactive_mode") == "true":
            o.set_module_inactive(name)
    if g.cfg.get("multispecies"):
        o.set_multispecies(True, **multispecies_fractions(g.modules))
    return o


@pytest.mark.parametrize("name", cases())
def test_oracle_reproduces_reference(name):
    g = Golden(name)
    o = make_oracle(g)
    tc, rl, pv = [], [], []
    for it in range(1, g.n_steps + 1):
        step = o.step()
        tc.append(o.subcycles("thermal_conduction"))
        rl.append(o.subcycles("radiative_losses"))
        pv.append(o.subcycles("physical_viscosity"))
        assert step == g.steps[it - 1], "iteration %d: step %s != reference %s" % (it, step.hex(), float(g.steps[it - 1]).hex())
        if it in g.frames:
            for v in OUT_VARS:
                assert same_bits(o.get(v), g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(o.get(v), g.frames[it][v]))
            for pname, ref in g.module_planes.get(it, {}).items():          # output_to_file planes: "thermal_conduction", "flux_saturation", "rad"
                vq = viscosity_plane_request(g.modules, pname)
                got = o.ms_plane(pname) if pname in MS_PLANES else o.viscosity_output(*vq) if vq else o.module_output(pname)
                assert got is not None and same_bits(got, ref), "module plane %s after iteration %d: %s" % (pname, it, mismatch(got, ref))
        o.ms_reset()                                                        # the fixtures store every iteration: the cumulative planes restart after each (evolution.cpp:36-41)
    names = [m[0] for m in g.modules]
    if "thermal_conduction" in names:
        assert tc == g.subcycle_counts("Thermal Subcycles")
    if "radiative_losses" in names:
        assert rl == g.subcycle_counts("Radiative Subcycles")
    if "physical_viscosity" in names:
        assert pv == g.viscous_subcycle_counts()


@pytest.mark.parametrize("name", cases(two_fluid=True))
def test_two_fluid_oracle_reproduces_reference(name):
    """oracle/ideal2f_oracle.inc (Ideal2F + EIC thermalization) against the tf_* fixtures of the unmodified reference binary: bit for bit,
    EIC included (the restatement and the reference both call glibc's pow / log)."""
    from oracle.oracle import Oracle2F
    g = Golden(name)
    assert all(m[0] == "eic_thermalization" for m in g.modules)
    o = Oracle2F(g.planes, g.ion_mass, g.adiabatic_index, remove_curl_terms=g.eqs_options.get("remove_curl_terms", False), eic=bool(g.modules), **g.kw)
    for it in range(1, g.n_steps + 1):
        step = o.step()
        assert step == g.steps[it - 1], "iteration %d: step %s != reference %s" % (it, step.hex(), float(g.steps[it - 1]).hex())
        if it in g.frames:
            for v in g.out_vars:
                assert same_bits(o.get(v), g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(o.get(v), g.frames[it][v]))
    o.close()


@pytest.mark.parametrize("name", cases(extended=True))
def test_extended_oracle_reproduces_reference(name):
    """The restatements of the SURVEY 8f rows -- open_moc boundary (oracle/moc_oracle.inc, with global_viscosity), the small solar modules
    (solar_small_modules_oracle.inc) and anomalous_resistivity (anomalous_resistivity_oracle.inc) -- against committed fixtures of the unmodified
    reference binary (tests/golden/make_golden.py): step-size history and every plane, bit for bit.  (tests/test_oracle_vs_live_reference.py
    sweeps more configurations against live runs where the reference binary is present.)"""
    g = Golden(name)
    if g.equation_set == "ideal_mhd_2E":
        from oracle.oracle import Oracle2E
        o = Oracle2E(g.planes, g.ion_mass, g.adiabatic_index, eic=any(m == "eic_thermalization" for m, _ in g.modules), **g.kw)
        for mname, kv in g.modules:
            if mname == "artificial_viscosity":
                kw = module_kwargs(mname, kv)
                o.set_viscosity(viscosity_terms_with_profiles(g.planes, kw.pop("terms")), **kw)
        for it in range(1, g.n_steps + 1):
            step = o.step()
            assert step == g.steps[it - 1], "iteration %d: step %s != reference %s" % (it, step.hex(), float(g.steps[it - 1]).hex())
            if it in g.frames:
                for v in g.out_vars:
                    assert same_bits(o.get(v), g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(o.get(v), g.frames[it][v]))
        o.close()
        return
    o = Oracle(g.planes, g.ion_mass, g.adiabatic_index, **g.kw)
    if "global_viscosity" in g.eqs_raw:
        o.set_global_viscosity(float(g.eqs_raw["global_viscosity"]))
    for mname, kv in g.modules:
        if mname == "anomalous_resistivity":
            o.set_anomalous_resistivity(**{k: float(v) for k, v in kv.items() if k != "output_to_file"})
        else:
            o.add_small_module(mname, **small_module_kwargs(mname, kv)[0])
    if g.cfg.get("multispecies"):
        o.set_multispecies(True, **multispecies_fractions(g.modules))
    for it in range(1, g.n_steps + 1):
        step = o.step()
        assert step == g.steps[it - 1], "iteration %d: step %s != reference %s" % (it, step.hex(), float(g.steps[it - 1]).hex())
        if it in g.frames:
            for v in OUT_VARS:
                assert same_bits(o.get(v), g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(o.get(v), g.frames[it][v]))
            for pname in MS_PLANES:                                         # multispecies_mode: ion / electron shares of the sources, joule heating of anomalous_resistivity
                if pname in g.module_planes.get(it, {}):
                    assert same_bits(o.ms_plane(pname), g.module_planes[it][pname]), "%s after iteration %d: %s" % (pname, it, mismatch(o.ms_plane(pname), g.module_planes[it][pname]))
            # output_to_file planes the restatement can form (anomalousresistivity.cpp:320-329, fieldheating.cpp:73-80)
            mp = g.module_planes.get(it, {})
            if "anomalous_template" in mp:
                _, tm = o.anomalous_state()
                assert same_bits(tm, mp["anomalous_template"]), "anomalous_template after iteration %d: %s" % (it, mismatch(tm, mp["anomalous_template"]))
                prod = tm * o.anomalous_diffusivity()
                assert same_bits(prod, mp["anomalous_diffusivity"]), "anomalous_diffusivity after iteration %d: %s" % (it, mismatch(prod, mp["anomalous_diffusivity"]))
            if "field_heating" in mp:
                k = [m for m, _ in g.modules if m != "anomalous_resistivity"].index("field_heating")
                assert same_bits(o.small_module_plane(k, 0), mp["field_heating"]), "field_heating after iteration %d" % it
        o.ms_reset()
    o.close()
