import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    """GPUs the product's own library can use.  A library that is missing or does not load is NOT "no GPU":
    the gpu tests then run and fail loudly (there is no CPU fallback to pass on)."""
    try:
        from simplemoc_b200 import api
        return api.device_count()
    except Exception:
        return -1


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing them."""
    if not any("gpu" in item.keywords for item in items) or _cuda_device_count() != 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Make sure the native libraries exist (cheap when they are up to date)."""
    import __graft_entry__ as entry
    entry.build()
    return True
