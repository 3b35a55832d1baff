import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libraries():
    """A fresh checkout has no built libraries (they are git-ignored): compile them once (nvcc cross-compiles sm_100a
    without a GPU).  Building is not a fallback — with the libraries missing the package refuses to import."""
    pkg = os.path.join(ROOT, "ray_tracer_challenge_b200")
    if not (os.path.exists(os.path.join(pkg, "librtc_b200.so")) and os.path.exists(os.path.join(pkg, "librtc_host.so"))):
        from ray_tracer_challenge_b200 import build as native

        native.build(verbose=True)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/librtc_oracle.so) bound to the reference-shaped Python API."""
    from tests.oracle_binding import load_oracle

    return load_oracle()


@pytest.fixture()
def rt(oracle):
    """Short alias used by the transcribed reference tests: `rt.Sphere()`, `rt.translation(...)`, ..."""
    return oracle
