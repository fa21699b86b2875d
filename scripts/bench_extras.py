#!/usr/bin/env python
"""Additional measurements bench.py attaches to its line as "extras" -- run by rank 0 in a SUBPROCESS with a timeout after the line's own numbers are taken, so that
nothing here can cost the bench line (a failure or a timeout becomes {"error": ...}).  Prints one JSON object.

  --mode single      (N = 1)  the drop-in binary end to end (spruce_b200/bin/run on OT-1024: text .state in, K steps, text end.state out -- SURVEY 8f-1's point:
                              through the reference's own surface the decimal text dominates), and the per-step cost of the secondary device paths
                              (scripts/module_perf.py: two-fluid, thermal conduction, physical viscosity)
  --mode ranks N     (N > 1)  `run -g N` (host/slabcomm.hpp: the drop-in binary on N GPUs) against the same binary on one GPU: mhd.out and end.state must be
                              byte-identical
No reference code and nothing under oracle/ is used here."""
import argparse
import json
import os
import signal
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
RUN = ROOT / "spruce_b200" / "bin" / "run"
DOMAIN_GRIDS = ["d_x", "d_y", "pos_x", "pos_y", "be_x", "be_y", "be_z"]
STATE_VARS = ["rho", "temp", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "grav_x", "grav_y"]


DEADLINE = [float("inf")]               # time.perf_counter() by which this script must have printed its line (--budget)


class NoTimeLeft(RuntimeError):
    pass


def remaining(limit_s, least_s=15):
    """the part's own limit, cut to what is left of the script's budget; a part that would get less than `least_s` is skipped"""
    left = DEADLINE[0] - time.perf_counter()
    if left < least_s:
        raise NoTimeLeft("skipped: %.0f s left of the budget bench.py gave" % max(left, 0.0))
    return min(limit_s, left)


def run_bounded(cmd, limit_s, env=None):
    """subprocess in its own process group; on a timeout the whole group goes (a torchrun launcher would otherwise leave its workers on the GPUs)"""
    limit_s = remaining(limit_s)
    p = subprocess.Popen(cmd, cwd=str(ROOT), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, start_new_session=True)
    try:
        out, err = p.communicate(timeout=limit_s)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        p.communicate()
        raise RuntimeError("%s ... did not finish within %d s" % (" ".join(cmd[:4]), limit_s))
    return p.returncode, out, err


def write_state(path, s):
    """The reference's .state text (fileio.cpp:14-80): header, then name / rows of comma-separated values; 17 significant digits are lossless."""
    P = s["planes"]
    nx, ny = P["rho"].shape
    with open(path, "w") as f:
        f.write("xdim,ydim\n%d,%d\nion_mass\n%.17g\nadiabatic_index\n%.17g\nt=0\n" % (nx, ny, s["ion_mass"], s["adiabatic_index"]))
        for name in DOMAIN_GRIDS + STATE_VARS:
            f.write(name + "\n")
            f.write("\n".join(",".join(map(repr, row)) for row in np.asarray(P[name], dtype=np.float64).tolist()) + "\n")


def config(steps, out_every=-1):
    return ("ideal_mhd = true\n{\n}\ntime_integrator = rk2\nmax_iterations = %d\niter_output_interval = %d\ntime_output_interval = -1.0\nstd_out_interval = -1\nwrite_interval = 1\n"
            "write_precision = 17\nduration = 1.0e30\nx_bound_1 = periodic\nx_bound_2 = periodic\ny_bound_1 = periodic\ny_bound_2 = periodic\nopen_boundary_strength = 1.0\n"
            "open_boundary_decay_base = 0.5\nepsilon = 0.2\ndensity_min = 1.0\ntemp_min = 1.0\nthermal_energy_min = 1.0e-30\noutput_flags = rho, temp, mom_x, mom_y, bi_x, bi_y, dt\n") % (steps, out_every)


def run_binary(state, out, steps, gpus=1, out_every=-1, timeout=40):
    out.mkdir(parents=True, exist_ok=True)
    (out / "run.config").write_text(config(steps, out_every))
    t0 = time.perf_counter()
    rc, _, err = run_bounded([str(RUN), "-m", "input", "-o", str(out), "-s", str(state)] + (["-g", str(gpus)] if gpus > 1 else []), timeout)
    wall = time.perf_counter() - t0
    if rc not in (-6, 134) or b"successfully reached" not in err:          # a completed run ends with abort(), like the reference (evolution.cpp:54-56)
        raise RuntimeError("run rc=%s: %s" % (rc, err.decode(errors="replace")[-400:]))
    return wall


def clean_env():
    """the environment without the variables of an enclosing torchrun (this script runs below a rank of one)"""
    drop = ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE", "ROLE_RANK", "ROLE_WORLD_SIZE", "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT",
            "OMP_NUM_THREADS")
    return {k: v for k, v in os.environ.items() if k not in drop and not k.startswith("TORCHELASTIC_") and not k.startswith("TORCH_NCCL_")}


def conduction_workload(n_gpus, size, steps, limit_s, device_plan=False):
    """BASELINE.json configs[4] / SURVEY 8d cfg-C (MHD + thermal conduction) through bench.py's own --workload mhd_tc leg, as a bounded side measurement"""
    cmd = [sys.executable]
    if n_gpus > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n_gpus), "--master-addr", "127.0.0.1", "--master-port", "29873"]
    cmd += [str(ROOT / "bench.py"), "--gpus", str(n_gpus), "--steps", str(steps), "--warmup", "3", "--workload", "mhd_tc", "--size", str(size), "--no-extra", "--no-cpu-baseline"]
    try:
        remaining(limit_s, least_s=45)                  # a whole bench.py process: not worth starting with less
        rc, so, se = run_bounded(cmd, limit_s, env=dict(clean_env(), **({"SPRUCE_DEVICE_SUBCYCLES": "1"} if device_plan else {})))
        lines = [ln for ln in so.decode(errors="replace").splitlines() if ln.startswith("{")]
        if not lines:
            return {"error": "rc %s: %s" % (rc, se.decode(errors="replace")[-300:])}
        l = json.loads(lines[-1])
        return {"workload": l["config"]["workload"], "n_gpus": l["n_gpus"], "size": size, "steps": steps, "value": l["value"], "unit": l["unit"], "ms_per_step": l["ms_per_step"],
                "thermal_conduction_subcycles_last_step": l["config"].get("thermal_conduction_subcycles_last_step"), "device_resident_subcycle_plan": l["config"].get("device_resident_subcycle_plan"), "e2e_value": (l.get("e2e") or {}).get("value"),
                "parity_vs_1gpu": l.get("parity_vs_1gpu")}
    except Exception as e:
        return {"error": repr(e)[:300]}


def mode_single(tmp):
    from spruce_b200 import synthetic
    out = {}
    n, steps = 1024, 100
    try:
        s = synthetic.orszag_tang(n, n)
        state = tmp / "ot.state"
        write_state(state, s)
        w1 = run_binary(state, tmp / "a", 1, timeout=60)
        wk = run_binary(state, tmp / "b", 1 + steps, timeout=60)
        out["dropin_binary_e2e"] = {
            "workload": "OT-%d through spruce_b200/bin/run: parse %.0f MB of .state text, set up, %d RK2 steps, write end.state + mhd.out" % (n, state.stat().st_size / 1e6, 1 + steps),
            "wall_s": wk, "wall_s_one_step_job": w1, "value": n * n * (1 + steps) / wk, "unit": "cell-updates/s",
            "stepping_only_value": n * n * steps / max(wk - w1, 1e-9), "note": "stepping_only = wall(%d steps) - wall(1 step): what remains of the job is text I/O and set-up" % (1 + steps)}
    except Exception as e:
        out["dropin_binary_e2e"] = {"error": repr(e)[:300]}
    out["conduction_workload_4096"] = conduction_workload(1, 4096, 10, 120)
    # the secondary device paths, as shipped and with the general (wrapping, range-testing) stencil instances for every cell: what the FAST instances buy
    try:
        remaining(150, least_s=45)
        rc, so, se = run_bounded([sys.executable, str(ROOT / "scripts" / "module_perf.py"), "2048"], 150)
        out["secondary_paths_2048"] = json.loads(so.decode()) if rc == 0 else {"error": se.decode(errors="replace")[-300:]}
    except Exception as e:
        out["secondary_paths_2048"] = {"error": repr(e)[:300]}
    # the solar module set with the device-resident sub-cycle plan off and on (SPRUCE_DEVICE_SUBCYCLES): what the host round trips cost, and the same bits either way
    try:
        remaining(60, least_s=30)
        rc, so, se = run_bounded([sys.executable, str(ROOT / "scripts" / "device_plan_perf.py"), "256", "1024"], 60)
        out["device_plan_solar_modules"] = json.loads(so.decode()) if rc == 0 else {"error": se.decode(errors="replace")[-300:]}
    except Exception as e:
        out["device_plan_solar_modules"] = {"error": repr(e)[:300]}
    # the conduction workload once more with the device-resident sub-cycle plan, if the budget still allows a whole bench.py process
    out["conduction_workload_4096_device_plan"] = conduction_workload(1, 4096, 10, 100, device_plan=True)
    return out


def mode_ranks(tmp, n_gpus):
    from spruce_b200 import synthetic
    nx, ny, steps = 96 * n_gpus + 5, 130, 12                       # uneven slabs
    res = {}
    for tag, zfull in (("2d", False), ("zfull", True)):
        s = synthetic.orszag_tang(nx, ny, zfull=zfull)
        state = tmp / ("in_%s.state" % tag)
        write_state(state, s)
        w1 = run_binary(state, tmp / ("one_" + tag), steps, 1, out_every=4)
        wn = run_binary(state, tmp / ("many_" + tag), steps, n_gpus, out_every=4)
        same = all((tmp / ("one_" + tag) / f).read_bytes() == (tmp / ("many_" + tag) / f).read_bytes() for f in ("mhd.out", "end.state"))
        res[tag] = {"files_equal": bool(same), "wall_s_1": w1, "wall_s_n": wn}
    return {"dropin_binary_ranks": {"grid": [nx, ny], "steps": steps, "n_gpus": n_gpus, "files_equal_to_one_gpu": all(v["files_equal"] for v in res.values()), "cases": res,
                                    "what": "spruce_b200/bin/run -g N (one forked rank per GPU, slabs along x, peer-store halo exchange, output gathered on rank 0) vs the same binary on one GPU"}}


def mode_ranks_all(tmp, n_gpus):
    out = {}
    try:
        out.update(mode_ranks(tmp, n_gpus))
    except Exception as e:
        out["dropin_binary_ranks"] = {"error": repr(e)[:300]}
    # cfg-C: 16384^2 needs the memory of >= 4 GPUs' slabs to stay small next to the enclosing run's; 8192^2 below that
    out["conduction_workload"] = conduction_workload(n_gpus, 16384 if n_gpus >= 4 else 8192, 5, 160)
    # the same with the device-resident sub-cycle plan (no rank waits for its host inside a batch), if the budget still allows it; its parity_vs_1gpu is the plan's slab check
    out["conduction_workload_device_plan"] = conduction_workload(n_gpus, 16384 if n_gpus >= 4 else 8192, 5, 120, device_plan=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True, choices=["single", "ranks"])
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--budget", type=float, default=0.0, help="seconds this script may take in all (0: no limit beyond the parts' own)")
    a = ap.parse_args()
    if a.budget > 0:
        DEADLINE[0] = time.perf_counter() + a.budget
    with tempfile.TemporaryDirectory(prefix="spruce_extras_") as d:
        try:
            out = mode_single(Path(d)) if a.mode == "single" else mode_ranks_all(Path(d), a.gpus)
        except Exception as e:
            out = {"error": repr(e)[:400]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
