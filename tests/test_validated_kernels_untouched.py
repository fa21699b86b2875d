"""The kernels the round-1 GPU runs validated (parity suite green on a B200, profiles/) must stay instruction-for-instruction identical while
unvalidated paths are added around them: scripts/sass_identity.py compares the built library with profiles/r1_validated_kernels.sass.gz.
A deliberate change to one of those kernels has to be re-validated on a GPU; then refresh the baseline:
    cuobjdump -sass spruce_b200/lib/libspruce_b200.so | gzip -9 > profiles/r1_validated_kernels.sass.gz"""
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs cuobjdump (CUDA toolkit)")
def test_gpu_validated_kernels_are_sass_identical_to_the_validated_build():
    from spruce_b200 import build
    lib = build.build()
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "sass_identity.py"), str(ROOT / "profiles" / "r1_validated_kernels.sass.gz"), str(lib)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    first = r.stdout.splitlines()[0]
    assert first.startswith("identical: 33") and "different: 0" in first, first
