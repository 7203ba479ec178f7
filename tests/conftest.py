import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def sq():
    import sqeazy_b200

    if not os.path.exists(sqeazy_b200.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    sqeazy_b200.lib()
    return sqeazy_b200


@pytest.fixture(scope="session")
def port():
    from oracle import oracle

    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle

    r = oracle.ref()
    if not r.available:
        pytest.skip("oracle/_ref (compiled reference stages) not built")
    return r


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def cuda(sq):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    sq.set_device(0)
    return torch
