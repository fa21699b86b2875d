"""Stage-kernel launch time against the number of waves: the same 4096-column grid with 512, 1024, 2048, 4096 rows on one GPU."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spruce_b200 import synthetic
from spruce_b200.domain import PlasmaDomain
KW = dict(xb=("periodic", "periodic"), yb=("periodic", "periodic"), integrator="rk2", epsilon=0.2, density_min=1.0, temp_min=1.0, thermal_energy_min=1.0e-30)
for nx in (512, 1024, 4096):
    s = synthetic.orszag_tang(nx, 4096)
    d = PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"], **KW)
    d.advance(3)
    ms = d.time_stage_kernel(20)
    import torch
    st = torch.cuda.ExternalStream(d.stream())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(st); d.advance(40); b.record(st); torch.cuda.synchronize()
    print(json.dumps({"rows": nx, "stagger": os.environ.get("SPRUCE_STAGGER_NS", "0"), "chunk": os.environ.get("SPRUCE_CHUNK_ROWS", "auto"), "stage1_ms": ms, "per_row_us": 1e3 * ms / nx * 8, "step_ms": a.elapsed_time(b) / 40}))
    d.close()
