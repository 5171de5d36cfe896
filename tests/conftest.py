"""pytest configuration: the `gpu` marker and shared fixtures."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
CRC24 = "0x1864CFB"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def g128():
    return dict(np.load(GOLDEN / "scl_p128.npz"))


@pytest.fixture(scope="session")
def gtoy():
    return dict(np.load(GOLDEN / "scl_toy.npz"))


@pytest.fixture(scope="session")
def gnr():
    return dict(np.load(GOLDEN / "nr_p128.npz"))


@pytest.fixture(scope="session")
def published():
    return json.loads((GOLDEN / "published.json").read_text())


@pytest.fixture(scope="session")
def gldpc():
    return dict(np.load(GOLDEN / "ldpc.npz"))
