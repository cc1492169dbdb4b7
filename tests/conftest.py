import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def known_answers():
    with open(os.path.join(GOLDEN, "reference_known_answers.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def sct_golden():
    with open(os.path.join(GOLDEN, "sct_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def librsk():
    """Build (if stale) and load the C-ABI library; works without a GPU."""
    import __graft_entry__ as entry

    entry.build()
    from pyradiosky_b200 import _lib

    return _lib.load()


HERA_LAT = -np.deg2rad(30 + 43 / 60 + 17.5 / 3600)
HERA_LON = np.deg2rad(21 + 25 / 60 + 41.9 / 3600)
HERA_H = 1073.0
JD_2015_03_01 = 2457082.5


@pytest.fixture(scope="session")
def hera():
    return HERA_LON, HERA_LAT, HERA_H, JD_2015_03_01
