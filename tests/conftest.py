import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "mcmc-multispde_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure libmspde.so exists (cross-compiles without a GPU)."""
    lib = os.path.join(ROOT, "mcmc-multispde_b200", "mspde", "libmspde.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()
    yield
