import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "slow: larger CPU-side cases")


@pytest.fixture(scope="session")
def lib():
    import gpu_pattern_matching_b200 as g
    return g.lib()


@pytest.fixture(scope="session")
def device():
    import gpu_pattern_matching_b200 as g
    if g.lib().acm_device_count() <= 0:
        pytest.fail("gpu test selected but no CUDA device is visible")
    d = g.Device(0)
    yield d
    d.close()
