import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle_c():
    """The C oracle (oracle/vt_oracle.c), built on demand with gcc."""
    from oracle import coracle
    coracle.lib()
    return coracle


@pytest.fixture(scope="session")
def vtlib():
    """libvtseg.so.  GPU tests must fail, not skip, when it is missing: there is no fallback path."""
    from video_transformer_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test started without a CUDA device"
    torch.zeros(1, device="cuda")
    return torch.device("cuda:0")
