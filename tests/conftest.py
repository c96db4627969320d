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
    """Build the oracle (CPU) once; libmcx.so is built by __graft_entry__.build() / mcell_b200.build."""
    from oracle import oracle_py
    oracle_py.build()
    from mcell_b200 import build as b
    b.build()
    yield
