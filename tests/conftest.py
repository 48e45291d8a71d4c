import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: larger CPU case (still part of the default CPU suite)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    oracle.lib()
    return oracle
