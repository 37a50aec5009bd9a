import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def dsb():
    import dsstne_b200
    return dsstne_b200


@pytest.fixture(scope="session")
def ctx(dsb):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    dsb.lib()          # raises loudly when the CUDA extension was not built
    c = dsb.Context(0)
    yield c
    c.close()
