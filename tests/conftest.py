import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle import RefLib, REF_SO
    if not os.path.exists(REF_SO) and not os.path.exists("/root/reference/src/lib/otezip.c"):
        pytest.skip("compiled reference not available")
    return RefLib()


@pytest.fixture(scope="session")
def ctx():
    from otezip_b200 import Ctx
    c = Ctx(0)
    yield c
    c.close()
