import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


os.environ.setdefault("CA_POISON_WS", "1")  # NaN-filled workspaces: reads of never-written slots cannot hide


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a B200 (sm_100) GPU; run with -m gpu")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a gpu-marked test was selected but no CUDA device is visible")
    from cognitive_aim_depth_estimation_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.ca_device_check(0), "ca_device_check")
    return torch.device("cuda:0")
