import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "detectron2"))


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is the product; build it (nvcc cross-compiles without a GPU) if it is missing."""
    from densepose_torchscript_b200 import build
    build.build()


@pytest.fixture(scope="session")
def reference_paths():
    if not have_reference():
        pytest.skip("/root/reference is not present on this machine")
    shims = os.path.join(ROOT, "oracle", "shims")
    for p in (REFERENCE, shims):
        if p not in sys.path:
            sys.path.insert(0, p)
    return REFERENCE
