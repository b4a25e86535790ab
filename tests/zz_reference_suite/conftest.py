"""The reference's own test-suite (staged unmodified by ``stage.py``) run against this
engine under the reference's import name: every test collected below this directory
imports ``adrt`` = ``/root/repo/adrt`` (an alias tree of ``adrt_b200``), needs the GPU
(there is no CPU fallback) and is therefore marked ``gpu``."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "shims")
if SHIMS not in sys.path:
    sys.path.insert(0, SHIMS)  # more_itertools stand-in (not installed in this image)


def pytest_collection_modifyitems(config, items):
    for item in items:
        if os.path.join("zz_reference_suite", "_staged") in str(item.fspath):
            item.add_marker(pytest.mark.gpu)
