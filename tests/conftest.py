import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on CPU")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement (oracle/libccd_oracle.so); built on demand — gcc only."""
    from oracle import bind
    if not bind.have_port():
        bind.build(ref=False, port=True)
    return bind.Port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled into oracle/_ref (absent when /root/reference never existed)."""
    from oracle import bind
    if not bind.have_ref():
        if os.path.exists("/root/reference/src/CTCD.cpp"):
            bind.build(ref=True, port=False)
        else:
            pytest.skip("oracle/_ref/libccdref.so not built and /root/reference absent")
    return bind.Ref()


@pytest.fixture(scope="session")
def ctx():
    """CUDA context through the C ABI.  No fallback: a missing library or device is an error, not a skip."""
    from collisiondetection_b200 import api
    c = api.Context(0)
    yield c
    c.close()
