"""pytest configuration: registers the `gpu` marker and puts the repo root on
sys.path so tests can import `oracle.pyoracle` (checker) and the product package
`sdr-server_b200` (via importlib -- the directory name has a hyphen)."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (sdr-server_b200/)."""
    return importlib.import_module("sdr-server_b200")


@pytest.fixture(scope="session")
def fixtures():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")) as f:
        return json.load(f)
