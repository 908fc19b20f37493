import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a GPU test that is selected on a box without CUDA must fail loudly, not skip silently,
    # unless it was deselected with -m "not gpu"
    pass


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ecad_b200 import _lib

    _lib.check(_lib.load().ecadk_device_check(0), "device_check")
    return torch.device("cuda:0")
