import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available():
    try:
        from adrt_b200 import _lib

        return _lib.load().adrt_b200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "small.npz"))


@pytest.fixture(scope="session")
def golden_hashes():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "hashes.json")) as f:
        return json.load(f)


def pytest_collection_modifyitems(config, items):
    # GPU tests must fail loudly (not skip) on a GPU box whose library is broken,
    # but on a CPU-only machine `-m gpu` simply has nothing it can run.
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
