"""The product's anomalous_resistivity module -- the functors and the sequence of passes of spruce_b200/csrc/anomres_cells.hpp, compiled for the HOST by
tests/hostcheck/anomres_host_check.cpp with loops in place of kernel launches -- against the CPU restatement (oracle/anomalous_resistivity_oracle.inc,
pinned to live reference runs): field, thermal energy, template, tracked null point and sub-cycle count, bit for bit."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from golden_util import mismatch, same_bits
from oracle.oracle import Oracle, anomalous_params
from spruce_b200 import synthetic
from test_oracle_vs_live_reference import AR_CASES, ar_kwargs, interior

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "anomres_host_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libanomres_host_check.so"

EXTRA = [
    ("ar_periodic_x_euler", dict(time_scale="0.5", safety_factor="0.4", flood_fill_threshold="2.0", smoothing_sigma="0.7", gradient_correction="true"), ("periodic", "periodic"), ("fixed", "open"), "euler"),
    ("ar_moc_bounds_rk4", dict(time_scale="0.4", time_integrator="rk4", flood_fill_threshold="1.2", metric_smoothing="false", flood_fill_max_radius="2.0e9"), ("open_moc", "open_moc"), ("fixed", "open_moc"), "rk2"),
    ("ar_frobenius_plain", dict(time_scale="2.0", template_mode="frobenius", frobenius_metric_coeff="1.0e58", metric_smoothing="false"), ("fixed", "fixed"), ("fixed", "open"), "euler"),
]


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC, ROOT / "spruce_b200" / "csrc" / "anomres_cells.hpp"]
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    L = C.CDLL(str(LIB))
    L.anomres_host_run.restype = C.c_int
    return L


@pytest.mark.parametrize("reverse", [0, 1])
@pytest.mark.parametrize("name,kv,xb,yb,integrator", AR_CASES + EXTRA, ids=[m[0] for m in AR_CASES + EXTRA])
def test_product_anomalous_resistivity_equals_oracle(lib, name, kv, xb, yb, integrator, reverse):
    nx, ny = 23, 23
    s = synthetic.stratified_loop(nx, ny, bump=0.5)
    floors = dict(density_min=1.0e7, temp_min=1.0e4, thermal_energy_min=1.0e-6)
    o = Oracle(s["planes"], s["ion_mass"], s["adiabatic_index"], xb=xb, yb=yb, integrator=integrator, **floors)
    o.run(2)                                               # a state with currents and a moved null point
    names = ["bi_x", "bi_y", "bi_z", "thermal_energy", "be_x", "be_y", "be_z", "n", "dt"]
    planes = [np.ascontiguousarray(o.get(v), dtype=np.float64).copy() for v in names]
    planes += [np.ascontiguousarray(s["planes"][v], dtype=np.float64) for v in ("pos_x", "pos_y")]
    a = ar_kwargs(kv)
    p = anomalous_params(**a)
    o.set_anomalous_resistivity(**a)                       # setupModule on this state
    iters, dt = 3, 0.2 * float(np.min(planes[8][2:-2, 2:-2]))
    ref = None
    for _ in range(iters):
        ref = o.anomalous_core(dt, raw_commit=True)
    (ri, rj), rt = o.anomalous_state()
    dx = np.ascontiguousarray(s["planes"]["d_x"][:, 0]); dy = np.ascontiguousarray(s["planes"]["d_y"][0, :])
    arr = (C.c_void_p * 11)(*[q.ctypes.data for q in planes])
    xl, xu, yl, yu = interior(tuple(b if b != "open_moc" else "open" for b in xb), tuple(b if b != "open_moc" else "open" for b in yb), nx, ny)
    bounds = (C.c_int * 4)(xl, xu, yl, yu)
    per = (C.c_int * 2)(int(xb[0] == "periodic"), int(yb[0] == "periodic"))
    moc = (C.c_int * 4)(*[int(b == "open_moc") for b in (xb[0], xb[1], yb[0], yb[1])])
    out = np.zeros((4, nx, ny)); tmpl = np.zeros((nx, ny)); diff = np.zeros((nx, ny)); ij = (C.c_int * 2)(); nsub = C.c_int()
    vp = lambda q: q.ctypes.data_as(C.c_void_p)
    rc = lib.anomres_host_run(arr, vp(dx), vp(dy), C.c_int(nx), C.c_int(ny), bounds, per, moc, vp(p), C.c_double(0.2), C.c_double(dt), C.c_int(iters), C.c_int(reverse),
                              vp(out), vp(tmpl), vp(diff), ij, C.byref(nsub))
    assert rc == 0
    assert (ij[0], ij[1]) == (ri, rj)
    assert nsub.value == o.anomalous_subcycles()
    assert same_bits(tmpl, rt), "%s template: %s" % (name, mismatch(tmpl, rt))
    assert np.count_nonzero(rt) > 0, "the case does not exercise the template"
    assert same_bits(diff, o.anomalous_diffusivity()), "%s diffusivity: %s" % (name, mismatch(diff, o.anomalous_diffusivity()))       # the other factor of the anomalous_diffusivity output plane
    for q, nm in enumerate(["bi_x", "bi_y", "bi_z", "thermal_energy"]):
        assert same_bits(out[q], ref[q]), "%s %s: %s" % (name, nm, mismatch(out[q], ref[q]))
    assert not same_bits(out[3], planes[3]), "the case does not heat anything"
    o.close()
