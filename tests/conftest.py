import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def dome1():
    from rti_b200 import lights
    return lights.dome1_lights()


@pytest.fixture(scope="session")
def M60(oracle, dome1):
    return oracle.pinv(dome1)


def golden_path(name):
    return os.path.join(GOLDEN, name)
