import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_cuda() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu tier instead of failing it
    (there is no CPU decode path to fall back on)."""
    if _have_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the gpu tier runs on the B200 box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once per session (nvcc cross-compiles without a GPU)."""
    from libacm_b200 import build
    build.build_all()
    yield


@pytest.fixture(scope="session")
def checker(request):
    """The reference itself (oracle/_ref = /root/reference/src/{decode,util}.c compiled unmodified).
    On the GPU box this is the ONLY acceptable checker: the parity tests fail loudly rather than
    compare the CUDA path with our own restatement.  The CPU tier of a box that has neither
    oracle/_ref nor /root/reference falls back to the restatement, which tests/test_oracle.py and
    tests/golden pin to the reference."""
    from oracle import bindings
    chk = bindings.best()
    if _have_cuda():
        assert chk.kind == "reference", (
            "oracle/_ref/libacm_ref.so is missing: the GPU parity tests refuse to run against "
            "the restatement (build it here with `make -C oracle` and let it travel)")
    return chk


@pytest.fixture(scope="session")
def oracle_port():
    from oracle import bindings
    return bindings.Oracle()
