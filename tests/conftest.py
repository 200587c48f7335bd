import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import get_oracle
    return get_oracle()


@pytest.fixture(scope="session")
def engine():
    """One CUDA context for the whole GPU session; fails loudly if the extension or GPU is missing."""
    import kmerutils_b200 as kb
    eng = kb.Engine(0)
    yield eng
    eng.close()
