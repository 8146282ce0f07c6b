import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/libnl_oracle.so), built on demand with gcc."""
    from oracle.nl_oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine handle on cuda:0; fails (never falls back) if the library or a GPU is missing."""
    import nonlin_b200 as nb

    return nb.default_engine(0)
