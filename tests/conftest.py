import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pnlib():
    import probnumdiffeq_jl_b200  # noqa: F401
    from probnumdiffeq_jl_b200 import lib

    lib.load()
    return lib


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.lib()
    return O
