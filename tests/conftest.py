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
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session", autouse=True)
def _oracle_warm_up():
    """The CPU checker evaluates every one of its paths once before any test compares against it: on some boxes of the GPU pool the
    first evaluation of a torch CPU expression in a process came back ~1e-4 off (``oracle.loss_port.warm_up`` has the evidence)."""
    from oracle import loss_port
    loss_port.warm_up()
    yield

