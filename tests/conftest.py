import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fixtures():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))


@pytest.fixture(scope="session")
def obj_mesh(fixtures):
    import oracle

    return oracle.load_3ds(fixtures["model/obj.3ds"].tobytes())


@pytest.fixture(scope="session")
def obj2_mesh(fixtures):
    import oracle

    return oracle.load_3ds(fixtures["model/obj2.3ds"].tobytes())
