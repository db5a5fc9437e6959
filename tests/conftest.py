import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have or os.environ.get("FBGPU_TEST_ON_EMULATOR"):      # tests/test_emu_kernels.py re-runs gpu-marked bodies on the CPU kernel interpreter
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_builds():
    """make sure the in-tree native pieces exist (no-op when they are up to date): libfbgpu.so (nvcc cross-compiles
    without a GPU), the datagen helper and the CPU oracle"""
    import __graft_entry__
    __graft_entry__.build()
