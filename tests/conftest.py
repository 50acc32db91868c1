import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from common import oracle
    return oracle()


@pytest.fixture(scope="session")
def orc64():
    from common import oracle
    return oracle(f64=True)


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nvdiffrecmc_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("libmcshade.so is missing on a GPU box: the product has no CPU fallback (run __graft_entry__.build())")
    return torch.device("cuda:0")
