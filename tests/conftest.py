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


REF_TESTS = ROOT / "tests" / "ref_tests"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.hookimpl(tryfirst=True)
def pytest_collection_modifyitems(config, items):
    """tests/ref_tests/ holds the reference's own 23 tests, vendored UNMODIFIED by scripts/vendor_ref_tests.py; they run
    against the GPU mirror package (`dl_scl_polar` = repo-root alias of polar_code_b200/dl_scl_polar, no CPU fallback),
    so every one of them gets the `gpu` marker here instead of by editing the files."""
    for it in items:
        if REF_TESTS in Path(str(it.fspath)).resolve().parents:
            it.add_marker(pytest.mark.gpu)


def pytest_sessionstart(session):
    """Refuse to run edited copies: the point of tests/ref_tests/ is that the files are the reference's, byte for byte."""
    import hashlib
    man = REF_TESTS / "MANIFEST.json"
    if not man.exists():
        return
    for name, digest in json.loads(man.read_text())["sha256"].items():
        got = hashlib.sha256((REF_TESTS / name).read_bytes()).hexdigest()
        if got != digest:
            raise pytest.UsageError(f"tests/ref_tests/{name} differs from the vendored reference file")


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
