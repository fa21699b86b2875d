"""Parity tests proper (run on the B200 with -m gpu): the CUDA path, called through the C ABI
(spruce_b200.capi -> libspruce_b200.so), against
  (a) the golden fixtures produced by the unmodified reference binary (tests/golden/), bit-for-bit;
  (b) the CPU oracle on seeded inputs at sizes it finishes in seconds, bit-for-bit;
  (c) size-independent properties at BASELINE.json's 4096^2.
Tolerance: none -- integer-like exactness. Planes must be identical up to the sign of zero, step sizes identical
as doubles."""
import numpy as np
import pytest

from golden_util import Golden, OUT_VARS, cases, mismatch, module_kwargs, same_bits

pytestmark = pytest.mark.gpu

BUILT_MODULES = set()      # module names the device library implements so far (extended as they land)


def make_domain(g: Golden):
    from spruce_b200.domain import PlasmaDomain
    d = PlasmaDomain(g.planes, g.ion_mass, g.adiabatic_index, **g.kw)
    for name, kv in g.modules:
        kw = module_kwargs(name, kv)
        if name == "ambient_heating":
            from ambient import heating_plane
            d.set_ambient_heating_plane(heating_plane(g, kw))
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


@pytest.mark.parametrize("name", golden_cases())
def test_golden_reference_outputs(name):
    g = Golden(name)
    d = make_domain(g)
    done = 0
    for it in sorted(g.frames):
        dts = d.advance(it - done)
        ref = g.steps[done:it]
        assert len(dts) == len(ref)
        assert all(a == b for a, b in zip(dts, ref)), "step history differs at iteration %d: %s vs %s" % (
            done + int(np.argmax(dts != ref)) + 1, [x.hex() for x in dts[:3]], [float(x).hex() for x in ref[:3]])
        done = it
        for v in OUT_VARS:
            got = d.grid(v)
            assert same_bits(got, g.frames[it][v]), "%s after iteration %d: %s" % (v, it, mismatch(got, g.frames[it][v]))
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
