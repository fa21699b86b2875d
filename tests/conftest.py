import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Order of the GPU suite (`-m gpu -x`): the parity tests proper first -- core ideal MHD / two-fluid / modules, then the extended rows, then the drop-in binary --,
# the non-strict first-run tests (written after a round's GPU budget was spent) and the sanitizer runs last, so that whatever happens in a newer test, the record
# of the validated ones is already written.  Tests without the gpu marker keep pytest's own order.
_GPU_FILE_ORDER = ["test_gpu_parity.py", "test_gpu_extended.py", "test_host_binary.py", "test_gpu_fast_instances.py", "test_gpu_device_plan.py", "test_gpu_sanitizer.py"]


def pytest_collection_modifyitems(config, items):
    def key(pair):
        pos, item = pair
        if item.get_closest_marker("gpu") is None:
            return (0, 0, 0, pos)
        name = Path(str(item.fspath)).name
        rank = _GPU_FILE_ORDER.index(name) if name in _GPU_FILE_ORDER else len(_GPU_FILE_ORDER)
        xf = item.get_closest_marker("xfail")
        first_run = 1 if (xf is not None and not xf.kwargs.get("strict", False)) else 0
        return (1, first_run, rank, pos)

    items[:] = [it for _, it in sorted(enumerate(items), key=key)]
