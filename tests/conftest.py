import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def handle():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a GPU; the engine has no CPU fallback"
    import cusp_autotuned_b200 as cusp
    return cusp.default_handle()


@pytest.fixture(scope="session")
def dev():
    import torch
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_spmv.npz"))
