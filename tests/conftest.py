import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once per session (nvcc cross-compiles without a GPU)."""
    from libacm_b200 import build
    build.build_all()
    yield


@pytest.fixture(scope="session")
def checker():
    """The reference itself when oracle/_ref is present, else the C restatement."""
    from oracle import bindings
    return bindings.best()


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import bindings
    return bindings.Oracle()
