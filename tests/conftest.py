import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): C++ restatement of the reference on OpenBLAS."""
    from oracle import oracle as O
    O.lib()
    O.set_threads(min(8, os.cpu_count() or 1))
    return O


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library, bound to cuda:0.  Fails (not skips) when it is missing."""
    import diaglib_b200 as D
    D.init(0)
    return D
