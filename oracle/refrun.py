"""TEST INFRASTRUCTURE ONLY -- helpers that drive the compiled, unmodified reference binary
(oracle/_ref/run, built by oracle/Makefile from /root/reference) and read/write its file formats.

Nothing in the product path (spruce_b200/) imports this module.  Allowed users: tests/,
tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

File formats follow the reference's reader/writer:
  .state  : source/mhd/fileio.cpp:14-80 (reader), :221-255 (writer)
  .config : source/mhd/fileio.cpp:84-122
  mhd.out : source/mhd/fileio.cpp:125-215
"""
from __future__ import annotations

import os
import subprocess
import time
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REF_BIN = ORACLE_DIR / "_ref" / "run"

DOMAIN_GRIDS = ["d_x", "d_y", "pos_x", "pos_y", "be_x", "be_y", "be_z"]  # plasmadomain.hpp:35-36
IDEALMHD_STATE = ["rho", "temp", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "grav_x", "grav_y"]  # idealmhd.hpp:28-30


def have_reference() -> bool:
    return REF_BIN.exists() and os.access(REF_BIN, os.X_OK)


def _fmt_plane(a: np.ndarray) -> str:
    # 17 significant digits: lossless for the reference's std::stod parser
    return "\n".join(",".join("%.17g" % v for v in row) for row in a) + "\n"


def write_state(path, planes: dict, ion_mass: float, adiabatic_index: float, t: float = 0.0, comments=()):
    """planes: name -> (xdim, ydim) float64 array, element [i, j] (i = x index, j = y index, j contiguous)."""
    first = next(iter(planes.values()))
    xdim, ydim = first.shape
    with open(path, "w") as f:
        for c in comments:
            f.write(c.rstrip("\n") + "\n")
        f.write("xdim,ydim\n%d,%d\n" % (xdim, ydim))
        f.write("ion_mass\n%.17g\n" % ion_mass)
        f.write("adiabatic_index\n%.17g\n" % adiabatic_index)
        f.write("t=%.17g\n" % t)
        for name, a in planes.items():
            assert a.shape == (xdim, ydim), name
            f.write(name + "\n")
            f.write(_fmt_plane(np.asarray(a, dtype=np.float64)))


def read_state(path):
    """Returns (meta, planes). meta has xdim, ydim, ion_mass, adiabatic_index, t, comments."""
    with open(path) as f:
        lines = f.read().split("\n")
    k = 0
    comments = []
    while lines[k] == "" or lines[k].startswith("#"):
        if lines[k].startswith("#"):
            comments.append(lines[k])
        k += 1
    assert lines[k].replace(" ", "") == "xdim,ydim"
    xdim, ydim = (int(s) for s in lines[k + 1].split(","))
    assert lines[k + 2].strip() == "ion_mass"
    ion_mass = float(lines[k + 3])
    assert lines[k + 4].strip() == "adiabatic_index"
    gamma = float(lines[k + 5])
    assert lines[k + 6].startswith("t=")
    t = float(lines[k + 6][2:])
    k += 7
    planes = {}
    while k < len(lines) and lines[k].strip() != "":
        name = lines[k].strip()
        rows = [np.array(lines[k + 1 + i].split(","), dtype=np.float64) for i in range(xdim)]
        planes[name] = np.stack(rows)
        assert planes[name].shape == (xdim, ydim), (name, planes[name].shape)
        k += 1 + xdim
    meta = dict(xdim=xdim, ydim=ydim, ion_mass=ion_mass, adiabatic_index=gamma, t=t, comments=comments)
    return meta, planes


def read_out(path):
    """Parse mhd.out -> (preamble planes, list of frames); each frame = dict(t=..., planes...)."""
    with open(path) as f:
        lines = f.read().split("\n")
    k = 0
    while lines[k].startswith("#") or lines[k] == "":
        k += 1
    assert lines[k].strip() == "xdim,ydim"
    xdim, ydim = (int(s) for s in lines[k + 1].split(","))
    k += 2
    pre = {}
    frames = []
    cur = None
    while k < len(lines):
        ln = lines[k].strip()
        if ln == "":
            k += 1
            continue
        if ln.startswith("t="):
            cur = {"t": float(ln[2:])}
            frames.append(cur)
            k += 1
            continue
        name = ln
        rows = [np.array(lines[k + 1 + i].split(","), dtype=np.float64) for i in range(xdim)]
        a = np.stack(rows)
        (pre if cur is None else cur)[name] = a
        k += 1 + xdim
    return pre, frames


def ideal_mhd_config(*, integrator="rk2", max_iterations=10, xb=("periodic", "periodic"), yb=("fixed", "fixed"),
                     epsilon=0.2, density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6,
                     output_flags=("rho", "temp", "thermal_energy", "mom_x", "mom_y", "mom_z", "bi_x", "bi_y", "bi_z", "dt"),
                     iter_output_interval=1, write_precision=17, open_strength=1.0, open_decay=0.5,
                     eqs="ideal_mhd", eqs_block=(), modules=(), std_out_interval=-1, duration=1.0e30,
                     write_interval=1, multispecies=False):
    """Config text accepted by the current reference code: equation-set block FIRST,
    braces in column 0, no blank lines inside it.  modules: iterable of (name, [(key, value), ...])."""
    L = ["%s = true" % eqs, "{"]
    L += ["%s = %s" % kv for kv in eqs_block]
    L += ["}"]
    L += [
        "time_integrator = %s" % integrator,
        "max_iterations = %d" % max_iterations,
        "iter_output_interval = %d" % iter_output_interval,
        "time_output_interval = -1.0",
        "std_out_interval = %d" % std_out_interval,
        "write_interval = %d" % write_interval,
        "write_precision = %d" % write_precision,
        "duration = %.17g" % duration,
        "x_bound_1 = %s" % xb[0], "x_bound_2 = %s" % xb[1],
        "y_bound_1 = %s" % yb[0], "y_bound_2 = %s" % yb[1],
        "open_boundary_strength = %.17g" % open_strength,
        "open_boundary_decay_base = %.17g" % open_decay,
        "epsilon = %.17g" % epsilon,
        "density_min = %.17g" % density_min,
        "temp_min = %.17g" % temp_min,
        "thermal_energy_min = %.17g" % thermal_energy_min,
    ]
    if output_flags:
        L.append("output_flags = " + ", ".join(output_flags))
    if multispecies:
        L.append("multispecies_mode = true")
    for name, kvs in modules:
        L += ["%s = true" % name, "{"]
        L += ["%s = %s" % kv for kv in kvs]
        L += ["}"]
    return "\n".join(L) + "\n"


def run_reference(state_path, config_text, out_dir, threads=None, timeout=3600, extra_args=()):
    """Run oracle/_ref/run -m input -o out_dir -s state (config placed inside out_dir, no -c:
    SURVEY.md 8c gotcha 1).  Exit status 134 == success (evolution.cpp:54-56).  Returns wall seconds."""
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    for p in out_dir.glob("*.config"):
        p.unlink()
    (out_dir / "run.config").write_text(config_text)
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    t0 = time.perf_counter()
    r = subprocess.run([str(REF_BIN), "-m", "input", "-o", str(out_dir), "-s", str(state_path), *extra_args],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    wall = time.perf_counter() - t0
    if r.returncode not in (0, 134, -6):
        raise RuntimeError("reference run failed rc=%s\nstdout:%s\nstderr:%s" % (
            r.returncode, r.stdout.decode()[-2000:], r.stderr.decode()[-2000:]))
    return wall, r.stdout.decode()
