import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: test needs >= 2 CUDA devices")


def pytest_collection_modifyitems(config, items):
    """``gpu`` tests are skipped (not errored) on a machine without CUDA, ``multigpu`` tests with fewer than two devices,
    so a plain ``pytest tests`` is green everywhere; ``-m gpu`` on the B200 box runs them all."""
    try:
        import torch

        n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        n = 0
    skip_gpu = pytest.mark.skip(reason="needs a CUDA device")
    skip_multi = pytest.mark.skip(reason="needs >= 2 CUDA devices")
    for item in items:
        if n == 0 and "gpu" in item.keywords:
            item.add_marker(skip_gpu)
        elif n < 2 and "multigpu" in item.keywords:
            item.add_marker(skip_multi)


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.fixture
def port():
    return free_port()
