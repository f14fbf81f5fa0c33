import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (no-op when they are up to date; GROM_SKIP_BUILD=1 trusts the files that are there)."""
    if not os.environ.get("GROM_SKIP_BUILD"):
        import __graft_entry__ as g
        g.build()
    yield
