import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the oracle are built in-tree; build them if a fresh checkout lacks them."""
    import sparsex_b200
    if not os.path.exists(sparsex_b200.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    from oracle import pyoracle
    pyoracle.build()
    yield


GOLDEN = os.path.join(ROOT, "tests", "golden")
