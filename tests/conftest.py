import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def nla():
    """The package in `nextla.jl_b200/` (built on demand; loading it needs no GPU)."""
    import __graft_entry__ as ge

    ge._load_build_module().build()
    return ge.load_package()


@pytest.fixture(scope="session")
def gpu(nla):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return nla.default_handle(0)
