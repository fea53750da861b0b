import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def c1():
    """Data of the reference's LM-solver tests (c++/tests/fixtures.h:9-68)."""
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "c1_fixture.json")))
    return {k: (np.array(v) if isinstance(v, list) else v) for k, v in fx.items()}


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden_ref.json")))
