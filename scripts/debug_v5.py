import os, sys, numpy as np
sys.path.insert(0, '.')
from spruce_b200 import synthetic
from spruce_b200.domain import PlasmaDomain
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = synthetic.orszag_tang(n, n)
kw = dict(density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
os.environ["SPRUCE_STAGE_KERNEL"] = "5"
a = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
os.environ["SPRUCE_STAGE_KERNEL"] = "4"
b = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **kw)
ka, kb = a.computeTimeDerivatives(), b.computeTimeDerivatives()
for i, nm in enumerate(PlasmaDomain.EVOLVED):
    bad = ~((ka[i] == kb[i]) | (np.isnan(ka[i]) & np.isnan(kb[i])))
    if bad.any():
        idx = np.argwhere(bad)
        print("RHS", nm, "bad", int(bad.sum()), "rows%64", sorted(set((idx[:, 0] % 64).tolist()))[:20], "cols%62", sorted(set((idx[:, 1] % 62).tolist()))[:20], "first", idx[:4].tolist())
da, db = a.advance(1), b.advance(1)
print("dt", da, db)
for nm in PlasmaDomain.EVOLVED:
    A, B = a.grid(nm), b.grid(nm)
    bad = ~((A == B) | (np.isnan(A) & np.isnan(B)))
    if bad.any():
        idx = np.argwhere(bad)
        print("STEP", nm, "bad", int(bad.sum()), "rows%64", sorted(set((idx[:, 0] % 64).tolist()))[:20], "cols%62", sorted(set((idx[:, 1] % 62).tolist()))[:20], "first", idx[:4].tolist())
print("done")
