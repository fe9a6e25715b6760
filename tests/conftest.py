import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

# CMARL_GOLDEN_DIR: run the fixture-based tests against a freshly regenerated set (tests/golden/compare_golden.py)
GOLDEN = Path(os.environ.get("CMARL_GOLDEN_DIR") or REPO / "tests" / "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(GOLDEN / f"{name}.npz")

    return load
