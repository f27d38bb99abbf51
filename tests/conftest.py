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
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def fs3d():
    """The product package; on a GPU box the CUDA library must be the thing that runs."""
    import fallingsand3d_b200 as pkg
    from fallingsand3d_b200 import build, _lib
    build.build()
    _lib.load()
    return pkg
