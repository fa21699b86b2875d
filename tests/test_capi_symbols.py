"""CPU-side checks of the drop-in boundary: the in-tree CUDA library builds, loads without a GPU, and exports every
symbol include/spruce_b200.h declares (no compute calls here); with no device the product path fails loudly."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    txt = (ROOT / "include" / "spruce_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(spruce_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from spruce_b200 import build, capi
    lib_path = build.build()
    assert lib_path.exists()
    L = ctypes.CDLL(str(lib_path))
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "library does not export %s" % n
        assert n in capi.SYMBOLS, "python binding lacks %s" % n
    assert sorted(capi.SYMBOLS) == names
    assert capi.load().spruce_abi_version() == capi.ABI_VERSION


def test_config_struct_matches_header_field_order():
    from spruce_b200 import capi
    txt = (ROOT / "include" / "spruce_b200.h").read_text()
    body = re.search(r"typedef struct spruce_config \{(.*?)\} spruce_config;", txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        m = re.match(r"\s*(int32_t|double)\s+(.*)", decl.strip(), flags=re.S)
        if m:
            fields += [(m.group(1), f.strip()) for f in m.group(2).split(",")]
    got = [("int32_t" if t is ctypes.c_int32 else "double", n) for n, t in capi.Config._fields_]
    assert got == fields


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from spruce_b200 import capi, domain, synthetic
    s = synthetic.orszag_tang(16, 12)
    with pytest.raises(capi.SpruceError, match="no CUDA device"):
        domain.PlasmaDomain(s["planes"], s["ion_mass"], s["adiabatic_index"])


def test_product_path_never_imports_oracle():
    for p in (ROOT / "spruce_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h"):
            txt = p.read_text()
            assert "oracle" not in txt, "%s mentions the oracle" % p


def test_rank1_cell_size_validation():
    from spruce_b200 import capi, domain
    dx = np.repeat(np.linspace(1, 2, 6)[:, None], 5, 1)
    dy = np.repeat(np.linspace(1, 2, 5)[None, :], 6, 0)
    a, b = domain.rank1_cell_sizes(dx, dy)
    assert a.shape == (6,) and b.shape == (5,)
    dx[2, 3] *= 1.0001
    with pytest.raises(capi.SpruceError):
        domain.rank1_cell_sizes(dx, dy)


def test_create_argument_errors_mirror_the_reference_messages():
    """spruce_domain_create validates its arguments before it touches a device: the messages are the reference's own
    (plasmadomain.cpp:140 'Grid too small for ghost zones', evolution.cpp 'Boundary cond'n must be defined') or name the unbuilt feature."""
    import ctypes as C
    from spruce_b200 import capi
    L = capi.load()

    def create(**kw):
        base = dict(abi_version=capi.ABI_VERSION, equation_set=0, xdim=16, ydim=12, x_bound_1=0, x_bound_2=0, y_bound_1=0, y_bound_2=0, time_integrator=1,
                    device=-1, row0=0, nx_local=16, rank=0, n_ranks=1, ion_mass=1.6726e-24, adiabatic_index=5.0 / 3.0, epsilon=0.2, density_min=1.0,
                    temp_min=1.0, thermal_energy_min=1e-30, open_boundary_strength=1.0, open_boundary_decay_base=0.5, time=0.0)
        base.update(kw)
        cfg = capi.Config(**base)
        h = C.c_void_p()
        rc = L.spruce_domain_create(C.byref(cfg), C.byref(h))
        return rc, L.spruce_last_error().decode()

    rc, msg = create(abi_version=capi.ABI_VERSION + 7)
    assert rc != 0 and "ABI version" in msg
    rc, msg = create(xdim=4, nx_local=4)
    assert rc != 0 and "Grid too small for ghost zones" in msg
    rc, msg = create(x_bound_1=9)
    assert rc != 0 and "Boundary cond'n must be defined" in msg
    rc, msg = create(y_bound_2=capi.BC["open_moc"], y_bound_1=capi.BC["fixed"], equation_set=capi.EQS["ideal_2F"], x_bound_1=capi.BC["fixed"], x_bound_2=capi.BC["fixed"])
    assert rc != 0 and "ideal_mhd only" in msg
    rc, msg = create(y_bound_2=capi.BC["open_moc"])
    assert rc != 0 and "pairs" in msg
    rc, msg = create(equation_set=1)
    assert rc != 0 and "equation set" in msg
    rc, msg = create(equation_set=capi.EQS["ideal_2F"], x_bound_1=capi.BC["open"])
    assert rc != 0 and "open boundaries" in msg
    rc, msg = create(n_ranks=4, row0=12, nx_local=8)
    assert rc != 0 and "bad slab" in msg
    rc, msg = create(time_integrator=7)
    assert rc != 0 and "time integrator" in msg
