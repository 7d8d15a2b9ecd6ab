import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The C-ABI library is a build artefact (git-ignored): compile it once if the tree has none.
    nvcc cross-compiles sm_100a without a GPU; the product has no fallback if this fails."""
    lib = os.path.join(ROOT, "xkv_b200", "libxkv_b200.so")
    if not os.path.exists(lib):
        import subprocess

        subprocess.check_call(["make", "-C", os.path.join(ROOT, "xkv_b200", "csrc"), "-j8"])


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass without a device
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container; run with gpurun")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
