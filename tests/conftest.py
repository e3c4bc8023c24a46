import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "refbin: needs the compiled reference under oracle/_ref (skipped when absent)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
