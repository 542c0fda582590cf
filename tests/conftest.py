import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "slow: larger sizes")


@pytest.fixture(scope="module", params=["bls12_381", "bn254"])
def curve_name(request):
    return request.param


@pytest.fixture(scope="module")
def ctx(curve_name):
    """A device context.  Fails (does not skip) when the CUDA library or a B200 is missing."""
    from hekaton_system_b200.backend import Context
    c = Context(curve_name, device=0)
    yield c
    c.close()
