"""The Python of the GPU tests without a GPU: every `-m gpu` test body of tests/test_gpu_parity.py, test_gpu_extended.py, test_gpu_fast_instances.py and test_gpu_device_plan.py runs in
this process over the recording stand-in for the device library (tests/hostcheck/capi_stub.c, built by tests/test_host_shell_calls.py) with `assert` statements
stripped:

    SPRUCE_STUB_LOG=/tmp/stub.log python -O scripts/gpu_tests_over_stub.py

The numbers are the stub's, so nothing numerical is checked; what surfaces is every Python-level mistake -- a wrong keyword, a missing method, an empty-array reduction,
a fixture key that does not exist -- in tests that were written after a round's GPU budget was spent and would otherwise meet a device for the first time at the round's
end.  Lines starting with BUG name them (tests that start the real binaries or pass pytest.param objects through this script's simple parametrize expansion show up too:
those are this script's limits, not the tests').  Test infrastructure only."""
import sys, textwrap, traceback, os, inspect
ROOT = __import__("pathlib").Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from pathlib import Path
import numpy as np
import ctypes as C
from spruce_b200 import capi
capi.LIB_PATH = ROOT / "tests" / "hostcheck" / "_build" / "libcapi_stub.so"        # only here: the product itself never loads anything but its own library
_L = C.CDLL(str(capi.LIB_PATH))
for k in [k for k in list(capi.SYMBOLS) if not hasattr(_L, k)]: del capi.SYMBOLS[k]
from spruce_b200 import domain as _dom
_dom.PlasmaDomain.advance = lambda self, k, max_time=None: np.full(int(k), 0.5)
_dom.PlasmaDomain.subcycles = lambda self, name: 1
_dom.PlasmaDomain.launch_count = lambda self: 0
import pytest
import test_gpu_extended as ge, test_gpu_parity as gp, test_gpu_fast_instances as gf, test_gpu_device_plan as gd
def run_inproc(code, env, timeout=90):
    old = {k: os.environ.get(k) for k in env}; os.environ.update(env)
    try:
        exec(compile(textwrap.dedent(code), "isolated", "exec"), {"__name__": "__main__"})
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
    return "ok"
for m in (ge, gf):
    if hasattr(m, "run_isolated"): m.run_isolated = run_inproc
bugs = 0
def params_of(fn):
    out = [dict()]
    for mk in getattr(fn, "pytestmark", []):
        if mk.name != "parametrize": continue
        names = [n.strip() for n in mk.args[0].split(",")] if isinstance(mk.args[0], str) else list(mk.args[0])
        vals = []
        for v in mk.args[1]:
            v = getattr(v, "values", v)            # pytest.param(...)
            if len(names) == 1:
                v = v[0] if (isinstance(v, tuple) and hasattr(mk.args[1][0], "values")) else v
                vals.append({names[0]: v})
            else:
                vals.append(dict(zip(names, v)))
        out = [dict(a, **b) for a in out for b in vals]
    return out
skip = ("4096", "slab", "sanitizer", "n_gpus")
for mod in (gp, ge, gf, gd):
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith("test_") or any(s in name for s in skip): continue
        for kw in params_of(fn):
            kw = dict(kw)
            if "tmp_path" in inspect.signature(fn).parameters:
                import tempfile; kw["tmp_path"] = Path(tempfile.mkdtemp())
            try:
                fn(**kw)
            except (AssertionError, pytest.skip.Exception, pytest.fail.Exception):
                pass
            except Exception as e:
                bugs += 1
                tb = traceback.extract_tb(sys.exc_info()[2])[-1]
                print("BUG", mod.__name__, name, list(kw.values())[:3], type(e).__name__, str(e)[:120], "@", tb.filename.split("/")[-1], tb.lineno)
print("done, bugs:", bugs)
