import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle bindings (test infrastructure); builds liboracle.so (and oracle/_ref when /root/reference exists)."""
    from oracle import pyoracle
    pyoracle.build(ref=os.path.isdir("/root/reference"))
    return pyoracle


@pytest.fixture(scope="session")
def ref_nofma(oracle):
    L = oracle.ref_lib("nofma")
    if L is None:
        pytest.skip("oracle/_ref not built (no /root/reference and no prebuilt .so)")
    return L


@pytest.fixture(scope="session")
def gpu_ctx():
    from line3dpp_b200 import capi
    ctx = capi.Context(0)   # raises without a GPU: the product has no CPU fallback
    yield ctx
    ctx.close()
