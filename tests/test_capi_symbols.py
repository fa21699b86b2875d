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
