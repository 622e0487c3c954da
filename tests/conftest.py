import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100) GPU; run with -m gpu on the GPU box")


def _have_b200():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without an sm_100 GPU skips the gpu-marked tests instead of failing in them."""
    if _have_b200():
        return
    skip = pytest.mark.skip(reason="needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and return the path of libvtamiq_b200.so."""
    from vtamiq_b200.build import build
    return build()
