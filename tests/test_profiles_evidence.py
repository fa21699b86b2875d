"""The committed evidence must agree with itself: every multi-GPU bench line kept in profiles/r2_scaling.jsonl carries the step-size and state hashes
the CPU oracle produced for the same job (profiles/r2_oracle_bench_hashes.jsonl, written by scripts/oracle_bench_hash.py), and the hash functions of
bench.py reproduce the oracle's line for a small job run here."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def lines(name):
    return [json.loads(l) for l in (ROOT / "profiles" / name).read_text().splitlines() if l.strip()]


def test_gpu_bench_lines_carry_the_oracle_hashes():
    oracle = {l["steps"]: l for l in lines("r2_oracle_bench_hashes.jsonl") if l["size"] == 4096}
    assert set(oracle) >= {20, 50, 100}
    runs = lines("r2_scaling.jsonl")
    assert {r["n_gpus"] for r in runs} >= {2, 4, 8}
    for r in runs:
        o = oracle[r["steps"]]
        assert r["parity_vs_1gpu"] is True
        assert (r["dt_hash"], r["state_hash"]) == (o["dt_hash"], o["state_hash"]), (r["n_gpus"], r["steps"], r["note"])


def test_hash_script_and_bench_agree_on_a_small_job():
    """The same two functions hash the oracle's run and the device run: 3 steps of a 64 x 48 Orszag-Tang job through the oracle, hashed both ways."""
    import bench
    from oracle.oracle import Oracle
    from spruce_b200 import synthetic
    from spruce_b200.domain import PlasmaDomain
    s = synthetic.orszag_tang(64, 48)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], **bench.KW)
    dts = [o.step() for _ in range(3)]
    planes = [o.get(v) for v in PlasmaDomain.EVOLVED]
    h1 = bench.state_hash([bench.row_digests(p) for p in planes])
    # the slab form: every rank hashes its rows, rank 0 concatenates the digests in rank order
    parts = [(0, 20), (20, 41), (41, 64)]
    h2 = bench.state_hash([b"".join(bench.row_digests(p[a:b]) for a, b in parts) for p in planes])
    assert h1 == h2 and len(h1) == 16
    assert bench.dt_hash(dts) == bench.dt_hash(np.array(dts))
