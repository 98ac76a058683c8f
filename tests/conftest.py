import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def ilm():
    """The product package (directory `immersedlayers.jl_b200`, imported as ilm_b200)."""
    import ilm_b200
    return ilm_b200


@pytest.fixture(scope="session")
def oracle():
    import ilm_oracle
    return ilm_oracle
