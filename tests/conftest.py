import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_libs():
    """Build the oracle (test infrastructure) and make sure the product library exists."""
    import oracle
    oracle.build()
    import pflare_b200
    if not os.path.exists(pflare_b200.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return True
