"""bench.py's own control flow without a GPU: the device library and the CUDA timing calls are replaced by stand-ins, the script runs its N = 1 leg
end to end and must print ONE JSON line with every key the contract names (metric, value, unit, e2e, roofline, gpu_launches, clocks, hashes, the
relaxed / full-instance measurements).  Guards the round-end bench run against Python-level mistakes that no GPU is needed to find."""
import io
import json
import sys
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


class FakeDomain:
    EVOLVED = ["rho", "mom_x", "mom_y", "mom_z", "thermal_energy", "bi_x", "bi_y", "bi_z"]
    created = []

    def __init__(self, planes, ion_mass, gamma, device=0, **kw):
        import os
        self.n = planes["rho"].shape
        self.launches = 0
        self.env = {k: os.environ.get(k) for k in ("SPRUCE_ARITH", "SPRUCE_STAGE_VARIANTS")}
        self.nonzero_z = bool(np.any(planes["mom_z"]))
        FakeDomain.created.append(self)

    def stream(self): return 0
    def launch_count(self): return self.launches
    def advance(self, k):
        self.launches += 3 * k
        return np.full(k, 0.0125)
    def time_stage_kernel(self, reps=10): return 0.9
    def subcycles(self, name): return 3
    def grid(self, name, out=None):
        a = np.zeros(self.n) if out is None else out
        a[...] = 1.0
        return a
    def set_thermal_conduction(self, **kw): pass
    def close(self): pass


class FakeEvent:
    t = 0.0
    def __init__(self, enable_timing=False): self.at = 0.0
    def record(self, stream=None):
        FakeEvent.t += 100.0
        self.at = FakeEvent.t
    def elapsed_time(self, other): return other.at - self.at


@pytest.mark.parametrize("extra_args", [[], ["--no-extra"], ["--arith", "relaxed"], ["--workload", "mhd_tc"]])
def test_single_gpu_leg_prints_the_contract_line(monkeypatch, extra_args):
    import torch
    import bench
    from spruce_b200 import build, domain
    monkeypatch.setattr(domain, "PlasmaDomain", FakeDomain)
    monkeypatch.setattr(build, "build", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda h: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    FakeDomain.created.clear()
    monkeypatch.setattr(sys, "argv", ["bench.py", "--size", "96", "--steps", "7", "--warmup", "3", "--no-cpu-baseline", *extra_args])
    monkeypatch.delenv("SPRUCE_ARITH", raising=False)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 7 and line["scaling"] == "strong" and line["dtype"] == "f64" and line["gpu_launches"] == 21
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and line["e2e"]["h2d_bytes_per_step"] > 0
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert len(line["dt_hash"]) == 16 and len(line["state_hash"]) == 16
    assert "extra_error" not in line
    if not extra_args:
        assert line["roofline_relaxed"]["frac"] > 0 and line["value_full_instance"]["value"] > 0
        relaxed = [d for d in FakeDomain.created if d.env["SPRUCE_ARITH"] == "relaxed"]
        assert len(relaxed) == 1 and not relaxed[0].nonzero_z          # the relaxed run is the bench workload itself ...
        assert FakeDomain.created[-1].nonzero_z and FakeDomain.created[-1].env["SPRUCE_ARITH"] is None      # ... the full-instance run is exact, with the z system switched on
        import os
        assert os.environ.get("SPRUCE_ARITH") is None
    else:
        assert "roofline_relaxed" not in line


def test_hashes_do_not_depend_on_the_partition_or_on_the_sign_of_zero():
    import bench
    rng = np.random.default_rng(5)
    a = rng.standard_normal((40, 17))
    a[3, 4] = 0.0
    whole = bench.row_digests(a)
    parts = b"".join(bench.row_digests(a[i:j]) for i, j in ((0, 13), (13, 14), (14, 40)))
    assert whole == parts
    b = a.copy(); b[3, 4] = -0.0
    assert bench.row_digests(b) == whole
    b[3, 4] = 1e-300
    assert bench.row_digests(b) != whole
    assert bench.dt_hash([0.1, 0.2]) != bench.dt_hash([0.1, 0.2000000000000001])


def test_extras_get_only_what_is_left_of_the_run_budget(monkeypatch):
    """the side measurements must not stretch the run: with the budget spent they are skipped, otherwise the script is told how long it may take"""
    import bench
    monkeypatch.setattr(bench, "RUN_BUDGET_S", 30.0)
    assert "skipped" in bench.extras_subprocess("single", 1, 330)
    seen = {}

    class P:
        pid, returncode = 0, 0
        def __init__(self, cmd, **kw): seen["cmd"] = cmd
        def communicate(self, timeout=None):
            seen["timeout"] = timeout
            return b'{"ok": 1}\n', b""
    monkeypatch.setattr(bench, "RUN_BUDGET_S", 1.0e4)
    monkeypatch.setattr(bench.subprocess, "Popen", P)
    assert bench.extras_subprocess("single", 1, 330, reserve_s=25.0) == {"ok": 1}
    assert seen["timeout"] == 330 and seen["cmd"][-2:] == ["--budget", "320"]
    monkeypatch.setattr(bench, "RUN_BUDGET_S", (bench.time.perf_counter() - bench.T_START) + 125.0)
    bench.extras_subprocess("single", 1, 330, reserve_s=25.0)
    assert 95 <= seen["timeout"] <= 100
