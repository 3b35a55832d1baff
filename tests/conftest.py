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
    """The CPU oracle (oracle/librtc_oracle.so) bound to the reference-shaped Python API."""
    from tests.oracle_binding import load_oracle

    return load_oracle()


@pytest.fixture()
def rt(oracle):
    """Short alias used by the transcribed reference tests: `rt.Sphere()`, `rt.translation(...)`, ..."""
    return oracle
