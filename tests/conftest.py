import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The CUDA library, loaded — GPU tests must run our kernels, never a fallback."""
    import torch
    from mla_b200 import _lib
    assert torch.cuda.is_available(), "GPU test selected without a GPU"
    lib = _lib.lib()
    assert lib.mla_device_check() == 0, lib.mla_last_error().decode()
    return lib


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
