"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs in the build container (no GPU): oracle vs golden fixtures, host logic, C-ABI
symbol checks, gloo world_size-2 tests.  `-m gpu` runs on a B200 and calls the CUDA path through
the C-ABI; it never touches /root/reference.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_path():
    return dict(np.load(os.path.join(GOLDEN, "path_small.npz")))


@pytest.fixture(scope="session")
def golden_model():
    return dict(np.load(os.path.join(GOLDEN, "model_slice.npz")))


@pytest.fixture(scope="session")
def golden_tiles():
    return dict(np.load(os.path.join(GOLDEN, "tiles.npz")))


@pytest.fixture(scope="session")
def golden_reductions():
    return dict(np.load(os.path.join(GOLDEN, "reductions.npz")))


@pytest.fixture(scope="session")
def golden_grads():
    return dict(np.load(os.path.join(GOLDEN, "grads.npz")))
