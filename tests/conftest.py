import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# helpers (_refshim, _gloo_worker) are imported as top-level modules: the name `tests` can be
# shadowed by the reference's own `tests` package once /root/reference is on sys.path
TESTS = os.path.join(ROOT, "tests")
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        from mlvectordb_b200 import _capi
        return _capi.lib().mlv_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: no silent fallbacks.
    # Without `-m gpu`, gpu-marked tests are deselected by the driver's `-m "not gpu"`.
    pass


@pytest.fixture(scope="session")
def has_gpu():
    return _has_gpu()
