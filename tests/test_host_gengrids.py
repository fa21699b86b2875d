"""spruce_b200/bin/gengrids -- the drop-in for the reference's UCNP problem generator (execs/gengrids.cpp: `.settings` sweep + `.config` template -> one directory per set of
conditions with plasma.settings, ucnp.config and init.state) -- against the REFERENCE'S OWN generator, compiled from its sources where they lie into
oracle/_ref/gengrids (oracle/Makefile).  Every file of every set must be byte-identical.  The sweeps take every branch: uniform / non-uniform grids, gaussian / exponential /
uniform clouds, the ion hole, both density modulations, all three runnable equation sets, `runs`, the set-number offset, `%` comments, units given as multiples of other rows and
of the derived plasma characteristics, config keys overridden, appended and left alone.  Host-only: no device library, no GPU."""
import filecmp
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
OURS = ROOT / "spruce_b200" / "bin" / "gengrids"
REF = ROOT / "oracle" / "_ref" / "gengrids"

BASE = """% UCNP sweep (comment line)
Nx = cgs = {nx}
Ny = cgs = {ny}
m_i = cgs = 1.455e-22
adiabatic_index = cgs = 1.66666667
grid_opt = opt = {grid_opt}
grid_growth = cgs = 1.05
grid_spread = cgs = 0.95
x_lim = sig_x = 4
y_lim = sig_y = 3.5  % in units of another row; the blank before the percent sign goes with the comment
dBdx = cgs = {dbdx}
n = cgs = {n}
n_min = n = 0.001
n_dist = opt = {n_dist}
sig_x = cgs = 0.1
sig_y = cgs = 0.12
n_hole = opt = {hole}
n_hole_amp = n = 0.2
n_hole_size = sig_x = 0.3
n_shock = opt = {shock}
n_shock_amp = n = 0.05
n_shock_lam = sig_x = 0.7
n_shock_sig = sig_x = 1.5
n_iaw = opt = {iaw}
n_iaw_amp = cgs = 0.1
n_iaw_sig = sig_x = 0.5
n_iaw_phase = cgs = 30
Te = cgs = {te}
Ti = cgs = 1
duration = tau = 0.3
epsilon = cgs = 0.15
{extra}"""

CONFIG = """{eqs} = true
{{
{block}}}
time_integrator = rk2
duration = 1.0e-6  # replaced by the sweep's value
epsilon = 0.2
x_bound_1 = open_ucnp
x_bound_2 = open_ucnp
y_bound_1 = open_ucnp
y_bound_2 = open_ucnp
density_min = 1.0
temp_min = 1.0e-3
thermal_energy_min = 1.0e-30
eic_thermalization = true
{{
}}
{tail}"""

CASES = {
    # name: (settings kwargs, config kwargs, extra command line)
    "sweep_2e_nonuniform_runs": (dict(nx=21, ny=17, grid_opt="non-uniform", dbdx=150, n="1e9, 3e9", n_dist="gaussian, exponential", hole="true", shock="false", iaw="true", te=20,
                                      extra="runs = opt = 2\n"), dict(eqs="ideal_mhd_2E", block="", tail=""), ["-o", "1", "-a", "0"]),
    "two_fluid_uniform_grid_shock_offset": (dict(nx=16, ny=24, grid_opt="uniform", dbdx=0, n="2e9", n_dist="gaussian", hole="false", shock="true", iaw="false", te="20, 40, 80",
                                                 extra="timescale = tau_x = 0.5\nlengthscale = sig = 2\n"), dict(eqs="ideal_2F", block="use_sub_cycling = false\n", tail="coulomb_explosion = true\n{\nstrength = 1.0e-3\n}"),
                                            ["-a", "5"]),
    "one_fluid_uniform_cloud": (dict(nx=15, ny=13, grid_opt="non-uniform", dbdx=75.5, n="5e8", n_dist="uniform", hole="true", shock="true", iaw="true", te=35,
                                     extra="gt_strength = cgs = 3.7\nmax_iterations = cgs = 50\n"), dict(eqs="ideal_mhd", block="", tail="global_temperature = true\n{\ngt_species = i\ngt_strength = 1.0\n}\n"), []),
}


def generate(binary, tmp, name, settings, config, extra):
    d = tmp / name
    d.mkdir(parents=True)
    (d / "sweep.settings").write_text(BASE.format(**settings))
    (d / "template.config").write_text(CONFIG.format(**config))
    r = subprocess.run([str(binary), "-p", str(d / "out"), "-s", str(d / "sweep.settings"), "-c", str(d / "template.config"), *extra], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode == 0, r.stderr.decode()[-1500:]
    return d / "out", r.stdout.decode()


@pytest.fixture(scope="module")
def binaries():
    subprocess.run(["make", "-C", str(ROOT / "spruce_b200" / "host"), "../bin/gengrids"], check=True, stdout=subprocess.DEVNULL)
    if not REF.exists():
        pytest.skip("oracle/_ref/gengrids is built where /root/reference exists (make -C oracle gengrids) and travels with the repo")
    return OURS, REF


@pytest.mark.parametrize("name", list(CASES))
def test_generated_files_are_byte_identical_to_the_reference_generators(binaries, tmp_path, name):
    ours, ref = binaries
    settings, config, extra = CASES[name]
    a, out_a = generate(ours, tmp_path / "ours", name, settings, config, extra)
    b, out_b = generate(ref, tmp_path / "ref", name, settings, config, extra)
    assert out_a == out_b                                                  # "Valid task array: [0 N]" and the peak density the n_shock branch prints
    files_a = sorted(p.relative_to(a) for p in a.rglob("*") if p.is_file())
    files_b = sorted(p.relative_to(b) for p in b.rglob("*") if p.is_file())
    assert files_a == files_b and len(files_a) % 3 == 0 and len(files_a) >= 3
    for f in files_a:
        assert filecmp.cmp(a / f, b / f, shallow=False), "%s differs from the reference generator's" % f
    assert {f.name for f in files_a} == {"plasma.settings", "ucnp.config", "init.state"}


def test_refusals_and_overwrite_guard(binaries, tmp_path):
    ours, _ = binaries
    settings, config, _ = CASES["one_fluid_uniform_cloud"]
    out, _ = generate(ours, tmp_path, "first", settings, config, [])
    d = tmp_path / "first"
    again = subprocess.run([str(ours), "-p", str(out), "-s", str(d / "sweep.settings"), "-c", str(d / "template.config")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert again.returncode != 0 and b"folder already exists and overwrite_flag=0" in again.stderr
    (d / "even.settings").write_text(BASE.format(**dict(settings, nx=14)))
    even = subprocess.run([str(ours), "-p", str(tmp_path / "even"), "-s", str(d / "even.settings"), "-c", str(d / "template.config")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert even.returncode != 0 and b"Number of grids must be odd" in even.stderr
    (d / "bad.settings").write_text(BASE.format(**settings).replace("x_lim = sig_x = 4", "x_lim = furlongs = 4"))
    bad = subprocess.run([str(ours), "-p", str(tmp_path / "bad"), "-s", str(d / "bad.settings"), "-c", str(d / "template.config")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert bad.returncode != 0 and b"The units for variable <x_lim> are not valid." in bad.stderr
