import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built libraries (they are kept out of history): build them once, exactly as
    __graft_entry__.build() does.  This is the build step, not a fallback -- nothing runs without the CUDA library."""
    import subprocess
    lib = os.path.join(ROOT, "rti_b200", "lib", "libptm_b200.so")
    cli = os.path.join(ROOT, "rti_b200", "bin", "ptm-encoder")
    if not (os.path.exists(lib) and os.path.exists(cli)):
        out = subprocess.run(["make", "-C", os.path.join(ROOT, "rti_b200", "csrc"), "all"], capture_output=True, text=True)
        if out.returncode != 0:
            raise pytest.UsageError("building rti_b200/csrc failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])


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
