import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _poison_recycled_device_memory(request):
    """Before every GPU test, fill a large block with NaN and hand it back to the caching allocator: the `torch.empty`
    outputs and scratch buffers the test then allocates start as NaN, so a kernel that leaves part of an output unwritten
    (or reads scratch it never wrote) fails deterministically instead of depending on what the last process left in HBM."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import torch
    if torch.cuda.is_available():
        junk = torch.full((128 * 1024 * 1024,), float("nan"), device="cuda:0")
        del junk
    yield
