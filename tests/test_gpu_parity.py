"""Parity tests proper (run on the B200 with -m gpu): the CUDA path, called through the C ABI
(spruce_b200.capi -> libspruce_b200.so), against
  (a) the golden fixtures produced by the unmodified reference binary (tests/golden/), bit-for-bit;
  (b) the CPU oracle on seeded inputs at sizes it finishes in seconds, bit-for-bit;
  (c) size-independent properties at BASELINE.json's 4096^2.
Tolerance: none -- integer-like exactness. Planes must be identical up to the sign of zero, step sizes identical
as doubles."""
import numpy as np
import pytest

from golden_util import Golden, MS_PLANES, OUT_VARS, cases, mismatch, module_kwargs, multispecies_fractions, same_bits, viscosity_plane_request

pytestmark = pytest.mark.gpu

BUILT_MODULES = {"thermal_conduction", "radiative_losses", "ambient_heating", "artificial_viscosity", "physical_viscosity"}
# Modules that call std::pow / std::log10: CUDA's libm and glibc differ in the last bit for some arguments, so these runs are
# held to the north star's tolerance (relative L-infinity <= 1e-9 per plane, same for the step sizes) instead of bit equality.
LIBM_MODULES = {"thermal_conduction", "radiative_losses", "physical_viscosity"}
REL_TOL = 1.0e-9


def rel_linf(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def make_domain(g: Golden):
    from spruce_b200.domain import PlasmaDomain
    d = PlasmaDomain(g.planes, g.ion_mass, g.adiabatic_index, **g.kw)
    for name, kv in g.modules:
        kw = module_kwargs(name, kv)
        if name == "ambient_heating":
            from ambient import heating_plane
            d.set_ambient_heating_plane(heating_plane(g, kw))
        elif name == "artificial_viscosity":
            from golden_util import viscosity_terms_with_profiles
            d.set_viscosity(viscosity_terms_with_profiles(g.planes, kw.pop("terms")), **kw)
        elif name == "physical_viscosity":
            from golden_util import physical_viscosity_coefficient
            ramp = kw.pop("ramp_length"); kw.pop("buffer_length")
            d.set_physical_viscosity(physical_viscosity_coefficient(g.planes, kw["coeff"], ramp), **kw)
        else:
            getattr(d, "set_" + name)(**kw)
    return d


def golden_cases():
    out = []
    for n in cases():
        g = Golden(n)
        if all(m[0] in BUILT_MODULES for m in g.modules):
            out.append(n)
    return out


# fixtures added after round 2's GPU budget was spent (they carry the output planes of physical_viscosity / artificial_viscosity): first executed by the round-end run
FIRST_RUN_FIXTURES = {"example_state_solar_modules_rk2", "solar_gaussian_state_modules_rk2", "loop_inactive_tc_rl_rk2", "loop_ms_solar_rk2", "loop_pv_diag_rk2", "ot_pv_diag_inactive", "loop_visc_diag_hv_rk2", "ot_visc_diag_hv_rk4", "ot_visc_diag_hv_euler"}
PV_PLANES = ("viscous_heating", "viscous_force_x", "viscous_force_y", "viscous_force_z")


@pytest.mark.parametrize("name", [pytest.param(n, marks=pytest.mark.xfail(reason="fixture written after round 2's GPU budget was spent; first device run", strict=False))
                                  if n in FIRST_RUN_FIXTURES else n for n in golden_cases()])
def test_golden_reference_outputs(name):
    g = Golden(name)
    d = make_domain(g)
    tcrl_out = []
    if name in FIRST_RUN_FIXTURES:                                    # inactive_mode / output_to_file of thermal_conduction and radiative_losses (only the newer fixtures look at them here)
        for mname, kv in g.modules:
            if mname in ("thermal_conduction", "radiative_losses"):
                if kv.get("inactive_mode") == "true":
                    d.set_module_inactive(mname)
                if kv.get("output_to_file") == "true":
                    d.set_module_output_to_file(mname)
                    tcrl_out += ["thermal_conduction", "flux_saturation"] if mname == "thermal_conduction" else ["rad"]
    ms_on = bool(g.cfg.get("multispecies"))
    if ms_on:                                                         # multispecies_mode = true: cumulative electron / ion / joule heating (plasmadomain.hpp:134-135)
        d.set_multispecies(True, **multispecies_fractions(g.modules))
    pv_out = any(m[0] == "physical_viscosity" and m[1].get("output_to_file") == "true" for m in g.modules)
    if pv_out:
        d.set_module_output_to_file("physical_viscosity")
    av_out = any(m[0] == "artificial_viscosity" and any(v == "true" for k, v in m[1].items() if k.startswith("visc_output_")) for m in g.modules)
    if av_out:
        d.set_module_output_to_file("artificial_viscosity")
        for pname, ref in g.module_planes.get(0, {}).items():          # before the first step: zero planes (viscosity.cpp:99-102)
            which, term = viscosity_plane_request(g.modules, pname)
            assert not ref.any() and not d.module_output("visc_%s:%d" % (which, term)).any(), pname
    exact = not any(m[0] in LIBM_MODULES for m in g.modules)
    done = 0
    tc, rl, pv = [], [], []
    for it in sorted(g.frames):
        dts = []
        for _ in range(it - done):
            if ms_on and done + len(dts) > 0:
                d.multispecies_reset()                               # the fixtures store every iteration: the planes restart after each (evolution.cpp:36-41)
            dts.append(d.advanceTime())
            if any(m[0] == "physical_viscosity" for m in g.modules):
                pv.append(d.subcycles("physical_viscosity"))
            if any(m[0] == "thermal_conduction" for m in g.modules):
                tc.append(d.subcycles("thermal_conduction"))
            if any(m[0] == "radiative_losses" for m in g.modules):
                rl.append(d.subcycles("radiative_losses"))
        dts = np.array(dts)
        ref = g.steps[done:it]
        assert len(dts) == len(ref)
        if exact:
            assert all(a == b for a, b in zip(dts, ref)), "step history differs at iteration %d: %s vs %s" % (
                done + int(np.argmax(dts != ref)) + 1, [x.hex() for x in dts[:3]], [float(x).hex() for x in ref[:3]])
        elif len(ref):                                               # (a fixture may keep frame 0: nothing stepped yet)
            assert np.max(np.abs(dts - ref) / ref) <= REL_TOL
        done = it
        for v in OUT_VARS:
            got = d.grid(v)
            if exact:
                assert same_bits(got, g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(got, g.frames[it][v]))
            else:
                assert rel_linf(got, g.frames[it][v]) <= REL_TOL, "%s after iteration %d: rel Linf %.3e" % (v, it, rel_linf(got, g.frames[it][v]))
        for pname in (tcrl_out if it > 0 else []):
            ref = g.module_planes[it][pname]
            assert rel_linf(d.module_output(pname), ref) <= REL_TOL, "%s after iteration %d: rel Linf %.3e" % (pname, it, rel_linf(d.module_output(pname), ref))
        if ms_on:
            for pname in MS_PLANES:
                ref = g.module_planes[it][pname]
                assert rel_linf(d.module_output(pname), ref) <= REL_TOL or not ref.any() and not d.module_output(pname).any(), "%s after iteration %d" % (pname, it)
        if av_out and it > 0:                                        # viscosity.cpp:351-376: what each term's last evaluation left, bit for bit
            for pname, ref in g.module_planes[it].items():
                which, term = viscosity_plane_request(g.modules, pname)
                got = d.module_output("visc_%s:%d" % (which, term))
                assert same_bits(got, ref), "%s after iteration %d: %s" % (pname, it, mismatch(got, ref))
        if pv_out:                                                   # physicalviscosity.cpp:292-308: the averages over the last step's sub-cycles
            for pname in PV_PLANES:
                ref = g.module_planes[it][pname]
                assert np.count_nonzero(ref) > 0 and rel_linf(d.module_output(pname), ref) <= REL_TOL, "%s after iteration %d: rel Linf %.3e" % (pname, it, rel_linf(d.module_output(pname), ref))
    if tc:
        assert tc == g.subcycle_counts("Thermal Subcycles")[:len(tc)]
    if rl:
        assert rl == g.subcycle_counts("Radiative Subcycles")[:len(rl)]
    if pv:
        assert pv == g.viscous_subcycle_counts()[:len(pv)]
    d.close()


@pytest.mark.parametrize("name", cases(two_fluid=True))
def test_two_fluid_golden_reference_outputs(name):
    """Ideal2F (+ EIC thermalization) against fixtures written by the unmodified reference binary.  Bit-for-bit without the
    module; eic_thermalization evaluates pow / log (glibc vs CUDA libm), so those runs are held to REL_TOL."""
    from spruce_b200.domain import PlasmaDomain
    g = Golden(name)
    d = PlasmaDomain(g.planes, g.ion_mass, g.adiabatic_index, equation_set="ideal_2F", eqs_options=g.eqs_options, **g.kw)
    exact = True
    for m, _ in g.modules:
        assert m == "eic_thermalization"
        d.set_eic_thermalization()
        exact = False
    done = 0
    for it in sorted(g.frames):
        dts = d.advance(it - done)
        ref = g.steps[done:it]
        assert len(dts) == len(ref)
        if exact:
            assert all(a == b for a, b in zip(dts, ref)), "step history differs: %s vs %s" % ([x.hex() for x in dts[:3]], [float(x).hex() for x in ref[:3]])
        else:
            assert np.max(np.abs(dts - ref) / ref) <= REL_TOL
        done = it
        for v in g.out_vars:
            got = d.grid(v)
            if exact:
                assert same_bits(got, g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(got, g.frames[it][v]))
            else:
                assert rel_linf(got, g.frames[it][v]) <= REL_TOL, "%s after iteration %d: rel Linf %.3e" % (v, it, rel_linf(got, g.frames[it][v]))
    d.close()


@pytest.mark.parametrize("nx,ny,integrator,xb,yb,eic,rct", [
    (151, 133, "rk2", ("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp"), False, False),
    (97, 121, "rk4", ("periodic", "periodic"), ("periodic", "periodic"), False, False),
    (66, 131, "euler", ("fixed", "reflect"), ("periodic", "periodic"), False, False),
    (130, 67, "rk2", ("periodic", "periodic"), ("open_ucnp", "open_ucnp"), False, True),
    (129, 129, "rk2", ("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp"), True, False),
])
def test_two_fluid_vs_oracle(nx, ny, integrator, xb, yb, eic, rct):
    """Ideal2F at sizes beyond the fixtures (odd, not multiples of the block width) against the pinned CPU restatement
    (oracle/ideal2f_oracle.inc): right-hand side, step sizes and every evolved + derived plane bit for bit; with EIC thermalization
    (libm pow / log on both sides, CUDA vs glibc) to REL_TOL."""
    from oracle.oracle import EVOLVED_2F, Oracle2F
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.ucnp_cloud(nx, ny, drift=2.0e3, bfield=5.0)
    kw = dict(xb=xb, yb=yb, integrator=integrator, density_min=1.0, temp_min=1.0e-3, thermal_energy_min=1e-30)
    o = Oracle2F(s["planes"], s["ion_mass"], s["adiabatic_index"], remove_curl_terms=rct, eic=eic, **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], equation_set="ideal_2F", eqs_options=dict(use_sub_cycling=False, remove_curl_terms=rct), **kw)
    if eic:
        d.set_eic_thermalization()
    same = (lambda a, b: same_bits(a, b)) if not eic else (lambda a, b: rel_linf(a, b) <= REL_TOL)
    for v in ("i_thermal_energy", "e_thermal_energy", "dt", "dt_i", "e_temp", "j_x", "divE"):
        assert same_bits(d.grid(v), o.get(v)), "after setup, %s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    k_dev, k_ref = d.computeTimeDerivatives(), o.rhs()
    for i, nm in enumerate(EVOLVED_2F):
        assert same(k_dev[i], k_ref[i]), "d(%s)/dt: %s" % (nm, mismatch(k_dev[i], k_ref[i]))
    nsteps = 4
    ref = np.array([o.step() for _ in range(nsteps)])
    dts = d.advance(nsteps)
    assert (np.array_equal(dts, ref) if not eic else np.max(np.abs(dts - ref) / ref) <= REL_TOL), (dts, ref)
    for v in EVOLVED_2F + ["dt", "dt_i", "i_temp", "e_temp", "j_y", "rho_c", "divB", "curlE_z", "i_dPdx", "b_hat_y", "press"]:
        assert same(d.grid(v), o.get(v)), "%s after %d steps: %s" % (v, nsteps, mismatch(d.grid(v), o.get(v)))
    o.close(); d.close()


def test_two_fluid_refusals():
    """use_sub_cycling = true (the Ideal2F default) cannot run in the reference; eic_thermalization needs Ideal2F grids."""
    from spruce_b200 import capi, synthetic
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.ucnp_cloud(24, 20)
    with pytest.raises(capi.SpruceError, match="use_sub_cycling"):
        PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], equation_set="ideal_2F", xb=("open_ucnp",) * 2, yb=("open_ucnp",) * 2)
    with pytest.raises(capi.SpruceError, match="open"):
        PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], equation_set="ideal_2F", xb=("open", "open"), eqs_options=dict(use_sub_cycling=False))
    o = synthetic.orszag_tang(24, 20)
    d = PlasmaDomain(o["planes"], o["ion_mass"], o["adiabatic_index"])
    with pytest.raises(capi.SpruceError, match="e_temp"):
        d.set_eic_thermalization()
    d.close()


@pytest.mark.parametrize("integrator,xb,yb", [
    ("rk2", ("periodic", "periodic"), ("periodic", "periodic")),
    ("rk4", ("periodic", "periodic"), ("periodic", "periodic")),
    ("euler", ("periodic", "periodic"), ("periodic", "periodic")),
])
def test_orszag_tang_vs_oracle(integrator, xb, yb):
    """Non-square grid that is not a multiple of the tile width, all z-components active."""
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.orszag_tang(150, 203, zfull=True)
    kw = dict(xb=xb, yb=yb, integrator=integrator, density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    for v in ("thermal_energy", "n", "dt", "temp", "v_x", "b_hat_y", "kinetic_energy"):
        assert same_bits(d.grid(v), o.get(v)), "after setup, %s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    k_dev, k_ref = d.computeTimeDerivatives(), o.rhs()
    for i, nm in enumerate(PlasmaDomain.EVOLVED):
        assert same_bits(k_dev[i], k_ref[i]), "d(%s)/dt: %s" % (nm, mismatch(k_dev[i], k_ref[i]))
    nsteps = 6
    ref = o.run(nsteps)
    dts = d.advance(nsteps)
    assert [x.hex() for x in dts] == [x.hex() for x in ref]
    for v in PlasmaDomain.EVOLVED + ["dt", "temp", "press", "b_mag"]:
        assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))
    assert d.time == o.time


@pytest.mark.parametrize("xb,yb", [
    (("periodic", "periodic"), ("fixed", "fixed")),
    (("reflect", "reflect"), ("open", "open")),
    (("open", "fixed"), ("reflect", "open")),
    (("open_ucnp", "open_ucnp"), ("open_ucnp", "open_ucnp")),
    (("fixed", "fixed"), ("periodic", "periodic")),
])
def test_stratified_loop_boundaries_vs_oracle(xb, yb):
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.stratified_loop(97, 140)
    kw = dict(xb=xb, yb=yb, integrator="rk2", density_min=3.0e8)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    nsteps = 12
    ref = o.run(nsteps)
    dts = d.advance(nsteps)
    assert [x.hex() for x in dts] == [x.hex() for x in ref]
    for v in PlasmaDomain.EVOLVED + ["dt", "temp"]:
        assert same_bits(d.grid(v), o.get(v)), "%s: %s" % (v, mismatch(d.grid(v), o.get(v)))


def test_edge_sizes():
    """Smallest legal grid (2*N_GHOST+1 per axis, plasmadomain.cpp:140), odd sizes (UCNP grids are odd), one-strip grids."""
    from oracle.oracle import Oracle
    from spruce_b200 import capi, synthetic
    from spruce_b200.domain import PlasmaDomain
    for nx, ny, xb, yb in [(5, 5, ("fixed", "fixed"), ("fixed", "fixed")), (5, 7, ("periodic", "periodic"), ("periodic", "periodic")),
                           (33, 33, ("open_ucnp",) * 2, ("open_ucnp",) * 2), (9, 65, ("periodic",) * 2, ("open", "open")),
                           (131, 6, ("reflect", "reflect"), ("periodic", "periodic"))]:
        s = synthetic.stratified_loop(nx, ny)
        kw = dict(xb=xb, yb=yb, integrator="rk2")
        o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
        ref, dts = o.run(3), d.advance(3)
        assert [x.hex() for x in dts] == [x.hex() for x in ref], (nx, ny)
        for v in PlasmaDomain.EVOLVED:
            assert same_bits(d.grid(v), o.get(v)), "%dx%d %s: %s" % (nx, ny, v, mismatch(d.grid(v), o.get(v)))
    s = synthetic.stratified_loop(4, 9)
    with pytest.raises(capi.SpruceError, match="too small"):
        PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"])


def test_batched_advance_equals_single_steps_and_max_time_stop():
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.orszag_tang(96, 80, zfull=True)
    kw = dict(density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
    a = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    b = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    da = a.advance(9)
    db = np.array([b.advanceTime() for _ in range(9)])
    assert np.array_equal(da, db)
    for v in PlasmaDomain.EVOLVED:
        assert np.array_equal(a.grid(v), b.grid(v))
    # PlasmaDomain::run stops at the first iteration with time >= max_time (evolution.cpp:26)
    c = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    tmax = float(np.sum(da[:4])) + 0.5 * da[4]
    dc = c.advance(9, max_time=tmax)
    assert len(dc) == 5 and np.array_equal(dc, da[:5]) and c.iter == 5


def test_full_size_4096_translation_invariance():
    """BASELINE.json's 4096^2 workload.  On a UNIFORM doubly periodic grid the update commutes with a cyclic shift of the
    input, bit for bit (every cell sees the same operands), which the oracle cannot check at this size but the
    arithmetic guarantees: advance(shift(U)) == shift(advance(U))."""
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    n = 4096
    s = synthetic.orszag_tang(n, n, stretch=0.0)
    dxu = np.full(n, 1.0e9 / n)
    P = dict(s["planes"])
    P["d_x"], P["d_y"] = dxu, dxu
    kw = dict(density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
    a = PlasmaDomain(P, s["ion_mass"], s["adiabatic_index"], **kw)
    sx, sy = 37, 1001
    Q = {k: (np.roll(v, (sx, sy), axis=(0, 1)) if getattr(v, "ndim", 0) == 2 else v) for k, v in P.items()}
    b = PlasmaDomain(Q, s["ion_mass"], s["adiabatic_index"], **kw)
    da, db = a.advance(3), b.advance(3)
    assert np.array_equal(da, db) and np.all(da > 0)
    for v in ("rho", "mom_x", "thermal_energy", "bi_y"):
        assert np.array_equal(np.roll(a.grid(v), (sx, sy), axis=(0, 1)), b.grid(v)), v
    # mass is transported in flux form: the cell-volume weighted sum is conserved to rounding
    m0 = float(np.sum(P["rho"])) ; m1 = float(np.sum(a.grid("rho")))
    assert abs(m1 - m0) / m0 < 1e-12


def test_full_size_4096_vs_oracle():
    """BASELINE.json's 4096^2 workload itself (OT-4096: non-uniform doubly periodic grid, RK2), 3 steps against the CPU oracle, bit for bit:
    step sizes and all eight evolved planes (SURVEY 8d's protocol).  About a minute of oracle time on the box's host cores."""
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    n = 4096
    s = synthetic.orszag_tang(n, n)
    kw = dict(xb=("periodic", "periodic"), yb=("periodic", "periodic"), integrator="rk2", epsilon=0.2, density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    dts = d.advance(3)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    ref = o.run(3)
    assert [x.hex() for x in dts] == [x.hex() for x in ref], (dts, ref)
    for v in PlasmaDomain.EVOLVED:
        a, r = d.grid(v), o.get(v)
        assert same_bits(a, r), "%s: %s" % (v, mismatch(a, r))


@pytest.mark.parametrize("integ,mods", [
    ("euler", [("thermal_conduction", dict(flux_saturation=True, integrator="rk4"))]),
    ("euler", [("field_heating", dict(coeff=1.0e-3, current_pow=1.0, b_pow=0.5, n_pow=0.25))]),
    ("euler", [("div_cleaning", dict(epsilon=0.1, time_scale=1.0e-3))]),
    ("rk2", [("thermal_conduction", dict(flux_saturation=False, integrator="rk2")), ("div_cleaning", dict(epsilon=0.1, time_scale=1.0e-3))]),
    ("rk4", [("thermal_conduction", dict(flux_saturation=True, integrator="euler"))]),
])
def test_2d_instance_with_modules_keeps_the_z_planes_zero(integ, mods):
    """A purely 2-D state (mom_z = bi_z = be = 0: the stage kernel's 6-quantity instance, which neither reads nor writes the z planes)
    together with modules that use the stage copy's planes as scratch: mom_z / bi_z must stay zero and everything must match the oracle.
    (Round 1's advisor finding: with euler the scratch planes were swapped in as the primary state.)"""
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.orszag_tang(84, 70, temp_mod=0.1)
    kw = dict(xb=("periodic", "periodic"), yb=("periodic", "periodic"), integrator=integ, density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    for name, mk in mods:
        if name in Oracle.SMALL:
            o.add_small_module(name, **mk)
        else:
            getattr(o, "set_" + name)(**mk)
        getattr(d, "set_" + name)(**mk)
    for it in range(4):
        a, b = d.advanceTime(), o.step()
        assert abs(a - b) / b <= REL_TOL, (it, a, b)
    for v in ("mom_z", "bi_z"):
        assert not np.any(d.grid(v)), v
    for v in PlasmaDomain.EVOLVED + ["temp", "dt"]:
        assert rel_linf(d.grid(v), o.get(v)) <= REL_TOL, "%s: rel Linf %.3e" % (v, rel_linf(d.grid(v), o.get(v)))


@pytest.mark.parametrize("mods", [
    [("thermal_conduction", dict(flux_saturation=False, integrator="euler"))],
    [("thermal_conduction", dict(flux_saturation=True, integrator="rk2"))],
    [("thermal_conduction", dict(flux_saturation=True, integrator="rk4"))],
    [("radiative_losses", dict(integrator="euler"))],
    [("radiative_losses", dict(integrator="rk4", prevent_subcycling=True))],
    [("thermal_conduction", dict(flux_saturation=True)), ("radiative_losses", dict(integrator="rk2")), ("ambient_heating", dict(heating_rate=1.0e-4))],
])
def test_solar_modules_vs_oracle(mods):
    """cfg-2 style run (thermal conduction + radiative losses + ambient heating on a gravity-stratified loop) at a size the
    oracle finishes in seconds: equal sub-cycle counts, planes and step sizes within 1e-9 relative."""
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.stratified_loop(70, 90, bump=0.5)
    kw = dict(xb=("periodic", "periodic"), yb=("fixed", "open"), integrator="rk2")
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    for name, mk in mods:
        if name == "ambient_heating":
            o.set_ambient_heating(**mk)
            mask = np.zeros((70, 90)); mask[:, 2:88] = 1.0
            d.set_ambient_heating_plane(mask * mk["heating_rate"])
        else:
            getattr(o, "set_" + name)(**mk)
            getattr(d, "set_" + name)(**mk)
    names = [m[0] for m in mods]
    for it in range(5):
        a, b = d.advanceTime(), o.step()
        assert abs(a - b) / b <= REL_TOL, (it, a, b)
        if "thermal_conduction" in names:
            assert d.subcycles("thermal_conduction") == o.subcycles("thermal_conduction")
        if "radiative_losses" in names:
            assert d.subcycles("radiative_losses") == o.subcycles("radiative_losses")
    for v in PlasmaDomain.EVOLVED + ["temp", "dt"]:
        assert rel_linf(d.grid(v), o.get(v)) <= REL_TOL, "%s: rel Linf %.3e" % (v, rel_linf(d.grid(v), o.get(v)))


@pytest.mark.parametrize("args", [("300", "210", "6", "rk2", "periodic", "p2p"), ("131", "96", "5", "rk4", "periodic", "p2p"), ("120", "80", "6", "rk2", "fixed", "p2p"),
                                  ("400", "130", "5", "rk2", "periodic", "p2p"), ("380", "70", "4", "euler", "periodic", "p2p"),      # >= 168 rows per slab: edge / interior launches overlap the exchange
                                  ("90", "70", "4", "euler", "reflect", "p2p"), ("300", "210", "6", "rk2", "periodic", "nccl"), ("120", "80", "6", "rk4", "fixed", "nccl")])
def test_slab_decomposition_equals_single_gpu(args):
    """N-GPU == 1-GPU bit for bit, for both halo transports (library peer stores over NVLink; NCCL send/recv).
    Needs >= 2 visible GPUs; spawns torchrun with 2 ranks."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(root / "scripts" / "mgpu_check.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode()
    assert r.returncode == 0 and "IDENTICAL" in out, out[-3000:]


@pytest.mark.parametrize("args", [("128", "96", "4", "rk2", "periodic", "p2p", "tc,rl,ah"), ("96", "80", "3", "euler", "reflect", "p2p", "tc,rl"),
                                  ("200", "64", "3", "rk4", "periodic", "p2p", "ah,tcsat,rl"), ("128", "96", "3", "rk2", "periodic", "p2p", "pv,tc"),
                                  ("120", "90", "3", "rk4", "periodic", "p2p", "av"), ("130", "97", "4", "rk2", "ucnp", "p2p", "2feic"), ("131", "96", "3", "rk4", "periodic", "p2p", "2f")])
def test_slab_decomposition_with_modules_equals_single_gpu(args):
    """("2f": the two-fluid equation set on slabs.)  Thermal conduction (halo exchange of the temperature plane per sub-cycle stage, sub-cycle counts from min / max all-gathers
    over peer memory), radiative losses and ambient heating on 2 slabs == 1 GPU, bit for bit, with equal sub-cycle counts."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29518", str(root / "scripts" / "mgpu_check.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode()
    assert r.returncode == 0 and "IDENTICAL" in out, out[-3000:]


def test_operators_vs_oracle():
    """spruce_operator: the PlasmaDomain differential operators on a host plane (source/mhd/derivs.cpp), bit for bit."""
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    for xb, yb, gen in [(("periodic", "periodic"), ("periodic", "periodic"), lambda: synthetic.orszag_tang(70, 51, zfull=True)),
                        (("fixed", "open"), ("reflect", "fixed"), lambda: synthetic.stratified_loop(45, 66))]:
        s = gen()
        kw = dict(xb=xb, yb=yb)
        o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
        d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
        q, v = o.get("temp"), o.get("v_x") - 0.3 * o.get("v_y")
        for op in ("derivative1D", "secondDerivative1D", "transportDerivative1D"):
            for index in (0, 1):
                a, b = d.operator(op, index, q, v), o.operator(op, index, q, v)
                assert same_bits(a, b), "%s index %d: %s" % (op, index, mismatch(a, b))
        assert same_bits(d.operator("laplacian", 0, q), o.operator("laplacian", 0, q))
