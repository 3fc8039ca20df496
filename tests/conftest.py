import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    orc.set_mode(orc.MODE_PAS)
    return orc


@pytest.fixture(scope="session")
def cuda_backend():
    """The global `cuda` instance (ntensors.pas:1053-1056).  Fails loudly when the .so or the GPU is missing."""
    from tensorium_b200 import tensor
    return tensor.initCUDART(0)
