import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUT = os.path.join(GOLDEN, "input")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def input_dir():
    return INPUT


def _has_gpu():
    try:
        from pfemfort_b200 import solver
        return solver.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """GPU tests must run the CUDA library: fail loudly (never skip silently) when it cannot be used."""
    from pfemfort_b200 import solver
    solver.load_library()
    if solver.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need a B200")
    return solver
