"""compute-sanitizer (memcheck, racecheck, synccheck) over the drop-in binary (spruce_b200/bin/run -> libspruce_b200.so) on small runs that take every shipped
instance of the stage kernel -- the 6-quantity and the 12-quantity instance, first / last / euler stage variants, rk4's run-time-stage instance --, the ghost-zone,
propagate and module kernels, and (2 GPUs) the peer-store halo kernels.  SURVEY section 5 prescribes these runs; round 2's GPU budget ended before they could be made by
hand, so they are tests: non-strict (first executed by the round-end suite), each bounded by a timeout.  A run is clean when the tool exits with status 0 under
--error-exitcode.  The binary is given a wall-clock limit (-r) so that it leaves through main() instead of the reference's abort-on-success."""
import os
import shutil
import signal
import subprocess
import time
from pathlib import Path

import pytest

from oracle import refrun
from spruce_b200 import synthetic

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
OURS = ROOT / "spruce_b200" / "bin" / "run"
SANITIZER = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
RUN_LIMIT_S = 60             # one tool run
BUDGET_S = 200               # all of this file: the suite it belongs to has to end within the driver's limit whatever the tools do
T0 = time.monotonic()

SOLAR = [("thermal_conduction", [("flux_saturation", "true"), ("epsilon", "0.1"), ("dt_subcycle_min", "1.0e-4")]),
         ("radiative_losses", [("cutoff_ramp", "1.0e3"), ("cutoff_temp", "3.0e4"), ("epsilon", "0.1")]), ("ambient_heating", [("heating_rate", "1.0e-4")])]
OT = dict(density_min=1.0, temp_min=1.0, thermal_energy_min=1e-30, xb=("periodic", "periodic"), yb=("periodic", "periodic"))
RUNS = {
    # 200 rows: several row chunks per launch; 70 columns: two column strips, the second one partial
    "ot_2d_rk2": (lambda: synthetic.orszag_tang(200, 70), dict(integrator="rk2", **OT)),
    "ot_zfull_rk2": (lambda: synthetic.orszag_tang(120, 70, zfull=True), dict(integrator="rk2", **OT)),
    "ot_zfull_rk4": (lambda: synthetic.orszag_tang(64, 70, zfull=True), dict(integrator="rk4", **OT)),
    "loop_walls_euler_modules": (lambda: synthetic.stratified_loop(48, 70, bump=0.5), dict(integrator="euler", xb=("reflect", "open"), yb=("fixed", "open"), modules=SOLAR)),
    "loop_moc_rk2": (lambda: synthetic.stratified_loop(48, 40), dict(integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "open_moc"))),
}


def sanitize(tool, name, tmp_path, gpus=1, relaxed=False, extra_env=None):
    if not Path(SANITIZER).exists():
        pytest.skip("compute-sanitizer is not installed")
    gen, ckw = RUNS[name]
    s = gen()
    state = tmp_path / "in.state"
    refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"])
    out = tmp_path / "out"
    out.mkdir()
    (out / "run.config").write_text(refrun.ideal_mhd_config(std_out_interval=-1, max_iterations=1000, iter_output_interval=-1, **ckw))
    cmd = [SANITIZER, "--tool", tool, "--error-exitcode", "99", "--target-processes", "all", str(OURS), "-m", "input", "-o", str(out), "-s", str(state), "-r", "1.0e-6"]
    if gpus > 1:
        cmd += ["-g", str(gpus)]
    env = dict(os.environ, SPRUCE_NVTX="0", **({"SPRUCE_ARITH": "relaxed"} if relaxed else {}), **(extra_env or {}))
    if time.monotonic() - T0 > BUDGET_S:
        pytest.skip("the sanitizer runs of this file have used their %d s" % BUDGET_S)
    p = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, start_new_session=True)      # its own process group: a timeout takes the tool AND the binary down
    try:
        raw, _ = p.communicate(timeout=RUN_LIMIT_S)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        p.communicate()
        pytest.fail("%s on %s did not finish within %d s" % (tool, name, RUN_LIMIT_S))
    r = p
    text = raw.decode(errors="replace")
    assert r.returncode == 0, "%s on %s: exit status %d\n%s" % (tool, name, r.returncode, text[-3000:])
    assert "ERROR SUMMARY: 0 errors" in text or "RACECHECK SUMMARY: 0 hazards" in text, text[-2000:]
    assert (out / "end.state").exists()                    # the run did step (16 steps, then the wall-clock limit)


FIRST_RUN = pytest.mark.xfail(reason="written after round 2's GPU budget was spent: first executed by the round-end suite", strict=False)


@FIRST_RUN
@pytest.mark.parametrize("tool,name", [("racecheck", "ot_2d_rk2"), ("memcheck", "ot_2d_rk2"), ("racecheck", "ot_zfull_rk2"), ("memcheck", "ot_zfull_rk2")])
def test_stage_kernel_first(tool, name, tmp_path):
    """the two shipped instances of the stage kernel before anything else: this file works against a time budget"""
    sanitize(tool, name, tmp_path)


@FIRST_RUN
@pytest.mark.parametrize("name,relaxed", [("ot_zfull_rk4", False), ("ot_2d_rk2", True), ("loop_walls_euler_modules", False)])
def test_racecheck_clean(name, relaxed, tmp_path):
    """shared-memory hazards of the stage kernel: the cp.async ring rows, the role-specialised warps' private exchange slots, the block reductions"""
    sanitize("racecheck", name, tmp_path, relaxed=relaxed)


@FIRST_RUN
@pytest.mark.parametrize("name", [n for n in RUNS if n not in ("ot_2d_rk2", "ot_zfull_rk2")])
def test_memcheck_clean(name, tmp_path):
    sanitize("memcheck", name, tmp_path)


@FIRST_RUN
def test_memcheck_clean_with_the_device_resident_subcycle_plan(tmp_path):
    """the solar module set planned on the device (k_sub_plan, plan-driven k_tc_stage / k_rl, k_tc_derive); a budget of 1 also takes the stop-and-enqueue-again path"""
    sanitize("memcheck", "loop_walls_euler_modules", tmp_path, extra_env={"SPRUCE_DEVICE_SUBCYCLES": "1", "SPRUCE_TC_BUDGET": "1"})


@FIRST_RUN
def test_synccheck_clean(tmp_path):
    sanitize("synccheck", "ot_zfull_rk2", tmp_path)        # the pair-wise named barrier of the 12-quantity instance


@FIRST_RUN
@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_two_gpu_halo_kernels_clean(tool, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sanitize(tool, "ot_2d_rk2", tmp_path, gpus=2)
