"""CPU oracle run of the benchmark job itself (OT-4096, RK2, K steps from the initial state) -> the same dt_hash / state_hash bench.py prints.
The hashes of the GPU runs (1, 2, 4, 8 GPUs) must equal these: a bit-for-bit check of the whole device path at the full benchmark size.
usage: python scripts/oracle_bench_hash.py [size] [steps,steps,...]   (about 16 s per 4096^2 step on 8 host cores)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle
from spruce_b200 import synthetic
from spruce_b200.domain import PlasmaDomain
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
marks = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "20,50,100").split(",")]
s = synthetic.orszag_tang(n, n)
o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **bench.KW)
dts, t0 = [], time.time()
for k in range(1, max(marks) + 1):
    dts.append(o.step())
    if k in marks:
        line = {"size": n, "steps": k, "dt_hash": bench.dt_hash(dts), "state_hash": bench.state_hash([bench.row_digests(o.get(v)) for v in PlasmaDomain.EVOLVED]), "oracle_seconds": time.time() - t0}
        print(json.dumps(line), flush=True)
