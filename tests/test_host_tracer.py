"""tracer_particles, a host-resident module of the host shell:
  * its bilinearInterpolate (spruce_b200/host/tracer.hpp) against the REFERENCE's own function (source/mhd/utils.cpp:55-75, compiled from its source by
    `make -C oracle refutils`), bit for bit, on points inside, on and outside the coordinate range (0.0 past the end, the zero-width bracket before the start);
  * the shell over the recording stand-in for the device library (uniform v_x = v_y = 1, step 0.5): init.tpstate is copied, particles.tpout / end.tpstate are
    written at the reference's cadence with its text format, a periodic axis wraps a particle, a wall axis removes it."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import refrun
from spruce_b200 import synthetic

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "tracer_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libtracer_check.so"
REFLIB = ROOT / "oracle" / "_ref" / "libref_utils.so"
DP = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def libs():
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC, ROOT / "spruce_b200" / "host" / "tracer.hpp", ROOT / "spruce_b200" / "host" / "grid.hpp"]
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    if not REFLIB.exists():
        if not Path("/root/reference/source").is_dir():
            pytest.skip("the reference sources are not here and oracle/_ref/libref_utils.so was not prebuilt")
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "refutils"], check=True, stdout=subprocess.DEVNULL)
    ours, ref = C.CDLL(str(LIB)), C.CDLL(str(REFLIB))
    for f in (ours.ours_bilinear, ref.ref_bilinear):
        f.restype = C.c_double
        f.argtypes = [C.c_double, C.c_double, DP, C.c_int, C.c_int, DP, DP]
    return ours, ref


def test_bilinear_interpolate_equals_the_reference_function(libs):
    ours, ref = libs
    rng = np.random.default_rng(2)
    nx, ny = 23, 17
    x = np.cumsum(rng.uniform(0.5, 2.0, nx)) * 1.0e7 - 3.0e7
    y = np.cumsum(rng.uniform(0.5, 2.0, ny)) * 3.0e6
    q = np.ascontiguousarray(rng.standard_normal((nx, ny)) * 1.0e5)
    px = np.concatenate([rng.uniform(x[0], x[-1], 4000), x[:5], [x[0] - 1.0, x[-1], x[-1] + 1.0, x[0]], rng.uniform(x[0] - 1e7, x[-1] + 1e7, 500)])
    py = np.concatenate([rng.uniform(y[0], y[-1], 4000), y[:5], [y[3], y[2], y[1], y[0] - 5.0], rng.uniform(y[0] - 1e6, y[-1] + 1e6, 500)])
    dp = lambda a: a.ctypes.data_as(DP)
    n_edge = 0
    for a, b in zip(px, py):
        u, v = ours.ours_bilinear(a, b, dp(q), nx, ny, dp(x), dp(y)), ref.ref_bilinear(a, b, dp(q), nx, ny, dp(x), dp(y))
        assert (np.float64(u).view(np.uint64) == np.float64(v).view(np.uint64)) or (np.isnan(u) and np.isnan(v)), (a, b, u, v)
        n_edge += (not np.isfinite(v)) or v == 0.0
    assert n_edge > 20                                   # the sample did reach the end-of-range and zero-width branches


STUB_SRC = ROOT / "tests" / "hostcheck" / "capi_stub.c"
STUB = ROOT / "tests" / "hostcheck" / "_build" / "libcapi_stub.so"
OURS = ROOT / "spruce_b200" / "bin" / "run"


def test_shell_writes_the_particle_files(tmp_path):
    STUB.parent.mkdir(exist_ok=True)
    hdr = ROOT / "include" / "spruce_b200.h"
    if not STUB.exists() or STUB.stat().st_mtime < max(STUB_SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["gcc", "-std=gnu11", "-O1", "-Wall", "-Werror", "-shared", "-fPIC", "-I", str(ROOT / "include"), str(STUB_SRC), "-o", str(STUB)], check=True)
    subprocess.run(["make", "-C", str(ROOT / "spruce_b200" / "host")], check=True, stdout=subprocess.DEVNULL)
    nx, ny = 20, 18
    s = synthetic.stratified_loop(nx, ny)
    X, Y = s["planes"]["pos_x"][:, 0], s["planes"]["pos_y"][0, :]
    state = tmp_path / "in.state"
    refrun.write_state(state, s["planes"], s["ion_mass"], s["adiabatic_index"])
    init = tmp_path / "my.tpstate"
    # one particle in the middle, one that leaves through the periodic x side, one that leaves through the wall at the top
    # (a particle crosses an upper edge only when its half-step point is still inside -- beyond the last coordinate the interpolated velocity is 0.0, utils.cpp:60 --
    #  i.e. from within (0.25, 0.5) of the edge at unit speed and step 0.5)
    p0 = [(0.5 * (X[3] + X[4]), 0.5 * (Y[5] + Y[6]), "mid dle"), (X[-1] - 0.4, Y[4] + 0.1, "wraps"), (X[7], Y[-1] - 0.4, "leaves"), (X[9], Y[-1] - 0.1, "stalls")]
    init.write_text("# particles\n" + "".join("%.17g, %.17g # %s\n" % p for p in p0))
    out = tmp_path / "out"
    out.mkdir()
    cfg = refrun.ideal_mhd_config(std_out_interval=1, integrator="rk2", xb=("periodic", "periodic"), yb=("fixed", "open"), max_iterations=4, iter_output_interval=2,
                                  modules=[("tracer_particles", [("init_file", str(init))])])
    (out / "run.config").write_text(cfg)
    env = dict(os.environ, LD_PRELOAD=str(STUB), SPRUCE_STUB_LOG=str(tmp_path / "calls.log"))
    r = subprocess.run([str(OURS), "-m", "input", "-o", str(out), "-s", str(state)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode in (-6, 134) and "successfully reached" in r.stderr.decode(), r.stderr.decode()[-2000:]
    assert "Tracer Particles On" in r.stdout.decode()
    assert (out / "init.tpstate").read_text() == init.read_text()
    # the stand-in's derived planes are 1.0 everywhere and every step is 0.5: a midpoint step moves a particle by 0.5 * interp(1) in x and y
    blocks = (out / "particles.tpout").read_text().strip().split("t=")[1:]
    assert [float(b.splitlines()[0]) for b in blocks] == [0.0, 1.0, 2.0]            # start, and after iterations 2 and 4 (iter_output_interval = 2)
    last = [ln for ln in blocks[-1].splitlines()[1:]]
    end = (out / "end.tpstate").read_text().strip().splitlines()
    assert last == end and len(end) == 3                                            # the particle that crossed the open top has left
    labels = [ln.split("#")[1] for ln in end]
    assert labels == ["middle", "wraps", "stalls"]                                  # whitespace is stripped from the whole line, label included (clearWhitespace)
    ys = [float(ln.split("#")[0].split(",")[1]) for ln in end]
    assert abs(ys[2] - (Y[-1] - 0.1)) < 1e-5                                                     # its half-step point lies past the last coordinate: zero velocity, as in the reference
    xs = [float(ln.split("#")[0].split(",")[0]) for ln in end]
    assert abs(xs[0] - (p0[0][0] + 2.0)) < 1e-6 * abs(X[-1] - X[0])                 # 4 steps of 0.5 at unit speed
    assert X[0] <= xs[1] < X[0] + 2.0                                               # wrapped to the other side of the periodic axis
