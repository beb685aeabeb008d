import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    assert pyoracle.selftest() == 0
    return pyoracle


@pytest.fixture(scope="session")
def b200lib():
    """The product library, built in-tree.  CPU tests only check that it loads and exports the ABI."""
    from boundless_b200 import build as _build, lib
    if not os.path.exists(lib.SO_PATH):
        _build.build()
    return lib.load()


@pytest.fixture(scope="session")
def gpu(b200lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    from boundless_b200 import lib
    lib.require_gpu(0)
    return torch
