import sys
import numpy as np
sys.path.insert(0, ".")
from spruce_b200 import synthetic
from spruce_b200.domain import PlasmaDomain
KW = dict(xb=("periodic", "periodic"), yb=("periodic", "periodic"), integrator="rk2", epsilon=0.2, density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30)
for n, zf, tm, coeff in ((64, True, 0.0, 4e-16), (64, False, 0.0, 4e-16), (64, False, 0.1, 4e-16), (512, False, 0.1, 4e-16), (2048, False, 0.1, 4e-16), (2048, False, 0.1, 4e-18)):
    s = synthetic.orszag_tang(n, n, zfull=zf, temp_mod=tm)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **KW)
    d.set_physical_viscosity(np.full((n, n), coeff), coeff=coeff, epsilon=0.2)
    dts = d.advance(2)
    print(n, zf, tm, coeff, "dt", dts, "nsub", d.subcycles("physical_viscosity"), "finite", np.isfinite(d.grid("thermal_energy")).all())
    d.close()
