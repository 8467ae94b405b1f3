import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def data_dir():
    return os.path.join(ROOT, "tests", "data")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ref_fixtures(golden_dir):
    return json.load(open(os.path.join(golden_dir, "ref_fixtures.json")))


@pytest.fixture(scope="session")
def ref_fuzz(golden_dir):
    return json.load(open(os.path.join(golden_dir, "ref_graph_fuzz.json")))["cases"]
