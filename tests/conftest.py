import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda_backend():
    """The CUDA backend under test.  Fails loudly (never falls back) when the library or GPU is missing."""
    from zgml_b200 import CudaBackend
    be = CudaBackend(0)
    yield be
    be.close()
