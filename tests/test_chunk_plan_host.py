"""spruce_b200/csrc/chunk_plan.hpp compiled for the HOST (tests/hostcheck/chunk_plan_check.cpp): for every slab shape -- rows from the smallest
domain to 5000, narrow to 16384-column grids, both CTA capacities, plain / overlapped launches, with and without SPRUCE_CHUNK_ROWS -- the launches
capi.cu forms from the plan cover every row exactly once through the row mapping of k_mhd_stage_xy, no CTA exceeds the shared-memory x tables, and
the halo rows of an overlapped slab leave from the edge launch."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "hostcheck" / "chunk_plan_check.cpp"
LIB = ROOT / "tests" / "hostcheck" / "_build" / "libchunk_plan_check.so"


@pytest.fixture(scope="module")
def lib():
    LIB.parent.mkdir(exist_ok=True)
    deps = [SRC, ROOT / "spruce_b200" / "csrc" / "chunk_plan.hpp"]
    if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    return C.CDLL(str(LIB))


def check(lib, nx, ny, cap, split, override=0):
    rows, ctas = C.c_int(), C.c_int()
    strips = (ny + 61) // 62
    rc = lib.chunk_plan_check(nx, strips, cap, int(split), override, C.byref(rows), C.byref(ctas))
    assert rc == 0, "plan of nx=%d ny=%d cap=%d split=%s override=%d fails check %d (rows %d)" % (nx, ny, cap, split, override, rc, rows.value)
    return rows.value, ctas.value


@pytest.mark.parametrize("cap", [148 * 5, 148 * 4])
def test_every_row_is_covered_exactly_once(lib, cap):
    for ny in (5, 62, 63, 200, 4096, 16384):
        for nx in list(range(5, 700)) + [1024, 2048, 2049, 4096, 4999, 16384]:
            check(lib, nx, ny, cap, False)
            if nx >= 3 * 56:                       # can_split (capi.cu)
                check(lib, nx, ny, cap, True)


def test_chunk_rows_override(lib):
    for override in (8, 11, 12, 16, 24, 47, 56, 64):
        for nx in (168, 181, 512, 2048, 4096):
            check(lib, nx, 4096, 740, False, override)
            check(lib, nx, 4096, 740, True, override)


def test_the_benchmark_slabs_fill_whole_waves(lib):
    """4096 columns = 67 strips, 740 resident CTAs of the 2-D instance: the slab of 1 / 2 / 4 / 8 ranks takes 7 / 4 / 2 / 1 waves with at most one wave's 8 percent to spare."""
    for world, waves in ((1, 7), (2, 4), (4, 2), (8, 1)):
        rows, ctas = check(lib, 4096 // world, 4096, 740, world > 1)
        assert 44 <= rows <= 56
        assert (waves - 1) * 740 < ctas <= waves * 740, (world, rows, ctas)
        assert ctas >= 0.92 * waves * 740, (world, rows, ctas)
