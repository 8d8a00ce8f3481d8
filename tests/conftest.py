import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import checkers
    return checkers.load_oracle()


@pytest.fixture(scope="session")
def ref_fet():
    import checkers
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return checkers.load_ref_fet()


@pytest.fixture(scope="session")
def ref_css():
    import checkers
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return checkers.load_ref_css()


@pytest.fixture(scope="session")
def fpt():
    """the product API; GPU tests only"""
    import fpt_b200.api as api
    if api.device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return api
