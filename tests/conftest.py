import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_pkg():
    """The package directory is `ring-zk_b200` (hyphen): import it by string."""
    return importlib.import_module("ring-zk_b200")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()
