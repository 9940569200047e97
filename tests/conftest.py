import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("rcx"))


@pytest.fixture(scope="module")
def gpu():
    """One CUDA context per test module through the C ABI.  A missing librcx.so is an error (the product has no CPU
    fallback and the GPU suite must not pass without it); a box without a CUDA device skips."""
    from rcorrector_b200 import Corrector, RcxError

    try:
        c = Corrector(0)
    except RcxError as e:
        if "rcx error 1:" in str(e):  # RCX_ERR_CUDA from rcx_create: no device visible
            pytest.skip("no CUDA device: %s" % e)
        raise
    yield c
    c.close()
