import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import cpzload  # noqa: E402

cpz = cpzload.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def pkg():
    return cpz


@pytest.fixture(scope="session")
def ctx():
    """One cpz context on cuda:0 for the whole GPU test session."""
    from cpz_b200 import engine
    c = engine.Context(0)
    yield c
    c.close()
