import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "first_hw_run: GPU test written after round 1's GPU budget ran out (never executed on a "
                                       "B200 yet); ordering only -- such tests run AFTER the hardware-validated ones so that "
                                       "`pytest -x` reports the validated state first")


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda it: it.get_closest_marker("first_hw_run") is not None)       # stable: file order kept otherwise


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
