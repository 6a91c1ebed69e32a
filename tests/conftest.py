import os
import subprocess
import sys

import pytest


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_PATH = os.path.join(ROOT, "oracle", "librpbmd_oracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure).  Built on demand from oracle/Makefile."""
    from reactive_pb_nn_md_b200._binding import Library
    if not os.path.exists(ORACLE_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return Library(ORACLE_PATH)


@pytest.fixture(scope="session")
def cuda_lib():
    from reactive_pb_nn_md_b200._binding import load_cuda
    return load_cuda()
