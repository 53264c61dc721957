import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` through gpurun)")


def _has_gpu():
    try:
        from clive2_b200 import runtime
        return runtime.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: nothing to do here.  Without `-m`, GPU tests are
    # skipped when no device is present so that a plain `pytest tests/` works in the build container.
    if config.getoption("-m"):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container (use gpurun + `-m gpu`)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
