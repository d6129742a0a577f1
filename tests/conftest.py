import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: without one they are skipped.  With a device present nothing is skipped -- a
    missing CUDA extension then fails loudly in the product code itself (hive_b200.lib() raises)."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:      # noqa
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="gpu test: no CUDA device")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "env_rollouts.npz"))
