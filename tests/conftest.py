"""pytest configuration: registers the `gpu` marker (tests that need a real B200) and puts the
repository root on sys.path so `latok_b200`, `oracle` and `tests.corpus` import from the tree."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
