import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "data", "")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_library():
    """The C-ABI library, built in-tree if it is not there yet (nvcc cross-compiles without a GPU)."""
    from petite_b200.build import build_library
    return build_library()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return lambda name: np.load(os.path.join(GOLDEN, name + ".npz"))
