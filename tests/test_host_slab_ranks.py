"""`run -g N` -- the drop-in binary on N GPUs (spruce_b200/host/slabcomm.hpp) -- without a GPU: N forked ranks over the recording stand-in for the device
library.  Checked: the row ranges every rank creates its handle with (multigpu.py's partition), the slab form of every plane argument, the set-up handshake
of the peer-store transport in the documented order (zero-plane masks OR-ed over the ranks, 64-byte IPC handles gathered in rank order, connect, eqs_setup,
initial exchange), one spruce_advance sequence on every rank, and the files: mhd.out / end.state of an N-rank run equal the 1-rank run's byte for byte
(state planes gathered through the shared buffer; a derived plane shows each rank's rows where they belong).  A rank that dies takes the others down."""
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import refrun
from spruce_b200 import synthetic
from spruce_b200.multigpu import partition

ROOT = Path(__file__).resolve().parents[1]
STUB_SRC = ROOT / "tests" / "hostcheck" / "capi_stub.c"
STUB = ROOT / "tests" / "hostcheck" / "_build" / "libcapi_stub.so"
OURS = ROOT / "spruce_b200" / "bin" / "run"


@pytest.fixture(scope="module")
def stub():
    STUB.parent.mkdir(exist_ok=True)
    hdr = ROOT / "include" / "spruce_b200.h"
    if not STUB.exists() or STUB.stat().st_mtime < max(STUB_SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["gcc", "-std=gnu11", "-O1", "-Wall", "-Werror", "-shared", "-fPIC", "-I", str(ROOT / "include"), str(STUB_SRC), "-o", str(STUB)], check=True)
    subprocess.run(["make", "-C", str(ROOT / "spruce_b200" / "host")], check=True, stdout=subprocess.DEVNULL)
    return STUB


def run(stub, tmp_path, s, cfg, n, tag, extra_env=None):
    state = tmp_path / "in.state"
    if not state.exists():
        refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"])
    out = tmp_path / ("out_" + tag)
    out.mkdir()
    (out / "run.config").write_text(cfg)
    log = tmp_path / ("calls_%s.log" % tag)
    env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(log), **(extra_env or {}))
    args = [str(OURS), "-m", "input", "-o", str(out), "-s", str(state)] + (["-g", str(n)] if n > 1 else [])
    r = subprocess.run(args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    logs = [Path(str(log) + ".r%d" % k).read_text().splitlines() for k in range(n)] if n > 1 else [log.read_text().splitlines()]
    return r, out, logs


MODULES = [("ambient_heating", [("heating_rate", "1.0e-4")]),
           ("thermal_conduction", [("flux_saturation", "true"), ("epsilon", "0.1"), ("dt_subcycle_min", "1.0e-4"), ("output_to_file", "true")])]


@pytest.mark.parametrize("n", [2, 3])
def test_n_ranks_write_the_files_one_rank_writes(stub, tmp_path, n):
    nx, ny = 23, 18                                        # 23 rows over 3 ranks: 8 + 8 + 7
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "open"), max_iterations=6, iter_output_interval=3, modules=MODULES,
                                  output_flags=("rho", "temp", "press", "mom_x", "bi_y", "dt"))
    r1, out1, (log1,) = run(stub, tmp_path, s, cfg, 1, "one")
    rn, outn, logs = run(stub, tmp_path, s, cfg, n, "n%d" % n)
    for r in (r1, rn):
        assert r.returncode in (-6, 134), r.stderr.decode()[-2000:]
        assert r.stderr.decode().count("Simulation successfully reached max simulation time or iterations") == 1      # rank 0 alone reports the end
    assert rn.stdout.decode().replace("out_n%d" % n, "out_one") == r1.stdout.decode()                                 # one stdout: rank 0's
    # end.state holds state variables only: what was uploaded comes back through the gather -> identical files
    assert (outn / "end.state").read_bytes() == (out1 / "end.state").read_bytes()
    # mhd.out: identical except for the derived planes, where the stand-in answers 1 + rank: each rank's rows must sit at its row range
    _, f1 = refrun.read_out(out1 / "mhd.out")
    _, fn = refrun.read_out(outn / "mhd.out")
    assert len(f1) == len(fn) == 3
    parts = partition(nx, n)
    for a, b in zip(f1, fn):
        assert a["t"] == b["t"]
        for v in ("rho", "mom_x", "bi_y"):
            assert np.array_equal(a[v], b[v])
        for v in ("press", "dt"):
            for k, (r0, rows) in enumerate(parts):
                assert np.all(b[v][r0:r0 + rows] == 1.0 + k)
        for v in ("thermal_conduction", "flux_saturation"):            # the module's device planes take the same gather
            assert np.array_equal(a[v], b[v])
    # per rank: the handle's row range, slab-sized plane arguments, the transport handshake in order, the same advance calls as the single rank
    adv1 = [ln for ln in log1 if ln.startswith("spruce_advance")]
    for k, log in enumerate(logs):
        create = dict(re.findall(r"(\w+)=(\S+)", next(ln for ln in log if ln.startswith("spruce_domain_create"))))
        assert (int(create["row0"]), int(create["nx_local"]), int(create["n_ranks"])) == (parts[k][0], parts[k][1], n)
        heating = log[log.index("spruce_module_ambient_heating") + 1]
        assert "count=%d " % (parts[k][1] * ny) in heating
        names = [ln.split()[0] for ln in log if ln.startswith("spruce_")]
        order = [c for c in names if c in ("spruce_plane_activity", "spruce_mgpu_ipc_export", "spruce_mgpu_ipc_connect", "spruce_eqs_setup", "spruce_mgpu_initial_exchange",
                                          "spruce_module_ambient_heating", "spruce_module_thermal_conduction")]
        assert order == ["spruce_plane_activity", "spruce_plane_activity", "spruce_mgpu_ipc_export", "spruce_mgpu_ipc_connect", "spruce_eqs_setup", "spruce_mgpu_initial_exchange",
                         "spruce_module_ambient_heating", "spruce_module_thermal_conduction"]
        assert "spruce_plane_activity set=%d" % ((1 << n) - 1) in log                       # the OR of every rank's mask
        assert "spruce_mgpu_ipc_connect n=%d handles_in_rank_order=1" % n in log
        assert [ln for ln in log if ln.startswith("spruce_advance")] == adv1


def test_host_resident_modules_and_too_many_ranks_are_refused(stub, tmp_path):
    s = synthetic.stratified_loop(20, 18)
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="euler", xb=("periodic", "periodic"), yb=("fixed", "open"), max_iterations=2, modules=[("sg_filtering", [("filter_interval", "1")])])
    r, _, _ = run(stub, tmp_path, s, cfg, 1, "sg1")
    assert "successfully reached" in r.stderr.decode()
    state = tmp_path / "in.state"
    for tag, n, cfg2, msg in (("sg2", 2, cfg, "one rank only"), ("many", 7, refrun.ideal_mhd_config(max_iterations=1, xb=("periodic", "periodic"), yb=("fixed", "open")), "at least 4 rows")):
        out = tmp_path / ("out_" + tag)
        out.mkdir()
        (out / "run.config").write_text(cfg2)
        env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(tmp_path / ("calls_%s.log" % tag)))
        p = subprocess.run([str(OURS), "-m", "input", "-o", str(out), "-s", str(state), "-g", str(n)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
        err = p.stderr.decode()
        assert p.returncode == 1 and msg in err and "successfully reached" not in err and "the other ranks were stopped" in err, err[-1500:]


def test_continue_mode_on_two_ranks_appends_like_one_rank(stub, tmp_path):
    """-m continue -g 2: every rank reads end.state (the time included), rank 0 appends to mhd.out; same files and the same exit status (SIGABRT on success) as one rank"""
    s = synthetic.stratified_loop(22, 18)
    state = tmp_path / "in.state"
    refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"])
    cfg = refrun.ideal_mhd_config(std_out_interval=1, max_iterations=3, integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "open"), duration=100.0)
    outs = {}
    for tag, g in (("one", []), ("two", ["-g", "2"])):
        out = tmp_path / ("out_" + tag)
        out.mkdir()
        (out / "run.config").write_text(cfg)
        env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(tmp_path / ("log_" + tag)))
        for args in (["-m", "input", "-o", str(out), "-s", str(state)], ["-m", "continue", "-o", str(out), "-d", "1.0"]):
            r = subprocess.run([str(OURS)] + args + g, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
            assert r.returncode in (-6, 134) and "successfully reached" in r.stderr.decode(), r.stderr.decode()[-1500:]
        outs[tag] = out
    assert (outs["one"] / "end.state").read_bytes() == (outs["two"] / "end.state").read_bytes()
    a, b = (outs["one"] / "mhd.out").read_text(), (outs["two"] / "mhd.out").read_text()
    assert a.count("\nt=") == b.count("\nt=") == 6 and [ln for ln in a.splitlines() if ln.startswith("t=")] == [ln for ln in b.splitlines() if ln.startswith("t=")]


def test_a_rank_that_dies_mid_run_takes_the_others_down(stub, tmp_path):
    """rank 1's device call fails after two steps: it aborts with the library's message, the parent releases the other ranks from their barrier (slabcomm.hpp: failure
    flag) and reports status 1 -- promptly, nobody is left waiting"""
    import time
    s = synthetic.stratified_loop(24, 18)
    cfg = refrun.ideal_mhd_config(std_out_interval=1, max_iterations=6, iter_output_interval=1, integrator="euler", xb=("periodic", "periodic"), yb=("fixed", "open"))
    state = tmp_path / "in.state"
    refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"])
    out = tmp_path / "out"
    out.mkdir()
    (out / "run.config").write_text(cfg)
    env = dict(os.environ, LD_PRELOAD=str(stub), SPRUCE_STUB_LOG=str(tmp_path / "log"), SPRUCE_STUB_FAIL_RANK="1")
    t0 = time.perf_counter()
    r = subprocess.run([str(OURS), "-m", "input", "-o", str(out), "-s", str(state), "-g", "3"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert time.perf_counter() - t0 < 20.0
    err = r.stderr.decode()
    assert r.returncode == 1 and "rank 1 ended early" in err and "successfully reached" not in err, err[-1500:]


@pytest.mark.parametrize("n", [2, 3])
def test_cumulative_and_viscosity_output_planes_are_gathered_like_every_other_plane(stub, tmp_path, n):
    """multispecies_mode and the output planes of artificial_viscosity / physical_viscosity on N ranks: every rank enables them, hands its ms_electron_heating_fraction over,
    resets the cumulative planes after every store; the planes reach mhd.out through the same row gather (the stand-in fills module planes with 7.0 on every rank)"""
    nx, ny = 23, 18
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    modules = [("ambient_heating", [("heating_rate", "1.0e-4"), ("ms_electron_heating_fraction", "0.3")]),
               ("artificial_viscosity", [("visc_opt", "global,local"), ("visc_strength", "3.0,0.5"), ("visc_vars_to_diff", "v_x,temp"), ("visc_vars_to_evol", "mom_x,thermal_energy"),
                                         ("visc_length", "0,0"), ("visc_species", "i,i"), ("hv_time_integrator", "rk2"), ("visc_output_visc", "true"), ("visc_output_lap", "true"),
                                         ("visc_output_strength", "false"), ("visc_output_timescale", "true")]),
               ("physical_viscosity", [("coeff", "1.0e-14"), ("epsilon", "0.1"), ("output_to_file", "true")])]
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "open"), max_iterations=4, iter_output_interval=2, modules=modules,
                                  output_flags=("rho", "mom_x"), multispecies=True)
    r1, out1, (log1,) = run(stub, tmp_path, s, cfg, 1, "one")
    rn, outn, logs = run(stub, tmp_path, s, cfg, n, "n%d" % n)
    for r in (r1, rn):
        assert r.returncode in (-6, 134), r.stderr.decode()[-2000:]
    assert (outn / "mhd.out").read_bytes() == (out1 / "mhd.out").read_bytes()
    _, frames = refrun.read_out(outn / "mhd.out")
    expected = ["cumulative_electron_heating", "cumulative_ion_heating", "cumulative_joule_heating", "mom_x_dqdt", "thermal_energy_dqdt", "mom_x_lap", "thermal_energy_lap",
                "mom_x_dt", "thermal_energy_dt", "viscous_heating", "viscous_force_x", "viscous_force_y", "viscous_force_z"]
    for f in frames:
        assert [k for k in f if k not in ("t", "rho", "mom_x")] == expected
        assert all(np.all(f[k] == 7.0) for k in expected)
    for log in logs:
        names = [ln.split()[0] for ln in log if ln.startswith("spruce_")]
        assert names.count("spruce_multispecies_mode") == 1 and names.count("spruce_multispecies_reset") == 2
        assert "spruce_module_ms_fraction ambient_heating fraction=0.29999999999999999" in log
        assert [ln.split()[1] for ln in log if ln.startswith("spruce_module_output_to_file")] == ["artificial_viscosity", "physical_viscosity"]
        assert [ln for ln in log if ln.startswith("spruce_module_output ")] == [ln for ln in log1 if ln.startswith("spruce_module_output ")]
