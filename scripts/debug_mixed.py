import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle.oracle import Oracle
from spruce_b200 import synthetic
from spruce_b200.domain import PlasmaDomain
s = synthetic.stratified_loop(97, 140)
for xb, yb in [(("open", "fixed"), ("reflect", "open")), (("open", "open"), ("reflect", "reflect")), (("open", "open"), ("fixed", "fixed")), (("fixed", "open"), ("open", "reflect"))]:
    kw = dict(xb=xb, yb=yb, integrator="rk2", density_min=3.0e8)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
    for it in range(0, 3):
        if it:
            a, b = d.advanceTime(), o.step()
            print(xb, yb, "it", it, "dt equal", a == b)
        for v in PlasmaDomain.EVOLVED:
            A, B = d.grid(v), o.get(v)
            bad = ~((A == B) | (np.isnan(A) & np.isnan(B)))
            if bad.any():
                idx = np.argwhere(bad)
                rows = sorted(set(idx[:, 0].tolist())); cols = sorted(set(idx[:, 1].tolist()))
                rel = np.abs(A - B)[bad].max() / np.abs(B).max()
                print("  it", it, v, "n_bad", int(bad.sum()), "rows", rows[:8], "cols", cols[:6], "...", cols[-3:], "rel", rel)
