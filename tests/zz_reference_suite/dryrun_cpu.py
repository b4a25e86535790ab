"""DEVELOPMENT TOOL (test infrastructure only, never imported by the product).

Runs the staged reference test-suite on a machine WITHOUT a GPU by standing the CPU
oracle in for ``libadrt_b200.so``'s host (NumPy) entry points, so that the Python
layer of the drop-in -- argument validation, error messages, arity checks, the ``adrt``
alias tree, generators -- can be exercised before spending GPU time.  Tests that need
device tensors (the multigrid drivers) cannot run this way and are reported as
failures/errors here; the real acceptance run is ``pytest tests -m gpu`` on the B200.

    python tests/zz_reference_suite/dryrun_cpu.py [pytest args]
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "shims"))

from adrt_b200 import _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402

_CT = {0: (np.float32, ctypes.c_float), 1: (np.float64, ctypes.c_double)}


def _view(ptr, shape, code):
    npt, ct = _CT[code]
    count = int(np.prod(shape))
    buf = (ct * count).from_address(ptr)
    return np.frombuffer(buf, dtype=npt).reshape(shape)


class FakeLib:
    """adrt_b200_host_* on the oracle; everything else raises."""

    def adrt_b200_device_count(self):
        return 1

    def adrt_b200_last_error(self):
        return b"fake"

    def adrt_b200_launch_count(self):
        return 0

    def _sino(self, fn, i, o, B, n, code, *_):
        _view(o, (B, 4, 2 * n - 1, n), code)[...] = fn(_view(i, (B, 4, 2 * n - 1, n), code))
        return 0

    def adrt_b200_host_adrt(self, i, o, B, n, code, dev):
        _view(o, (B, 4, 2 * n - 1, n), code)[...] = O.adrt(_view(i, (B, n, n), code))
        return 0

    def adrt_b200_host_adrt_init(self, i, o, B, n, code, dev):
        _view(o, (B, 4, 2 * n - 1, n), code)[...] = O.adrt_init(_view(i, (B, n, n), code))
        return 0

    def adrt_b200_host_bdrt(self, i, o, B, n, code, dev):
        return self._sino(O.bdrt, i, o, B, n, code)

    def adrt_b200_host_iadrt(self, i, o, B, n, code, dev):
        return self._sino(O.iadrt, i, o, B, n, code)

    def adrt_b200_host_adrt_step(self, i, o, B, n, step, code, dev):
        return self._sino(lambda a: O.adrt_step(a, step), i, o, B, n, code)

    def adrt_b200_host_bdrt_step(self, i, o, B, n, step, code, dev):
        return self._sino(lambda a: O.bdrt_step(a, step), i, o, B, n, code)

    def adrt_b200_host_interp_to_cart(self, i, o, B, n, code, dev):
        _view(o, (B, n, 4 * n), code)[...] = O.interp_to_cart(_view(i, (B, 4, 2 * n - 1, n), code))
        return 0

    def adrt_b200_host_fmg_restriction(self, i, o, B, n, code, dev):
        _view(o, (B, 4, n - 1, n // 2), code)[...] = O.press_fmg_restriction(_view(i, (B, 4, 2 * n - 1, n), code))
        return 0

    def adrt_b200_host_fmg_prolongation(self, i, o, B, h, w, code, dev):
        _view(o, (B, 2 * h, 2 * w), code)[...] = O.press_fmg_prolongation(_view(i, (B, h, w), code))
        return 0

    def adrt_b200_host_fmg_highpass(self, i, o, B, h, w, code, dev):
        _view(o, (B, h, w), code)[...] = O.press_fmg_highpass(_view(i, (B, h, w), code))
        return 0


if __name__ == "__main__":
    import pytest

    fake = FakeLib()
    _lib.load = lambda: fake
    _lib.require_device = lambda: None
    sys.exit(pytest.main([os.path.join(HERE, "_staged"), "-q", "--no-header", "-p", "no:cacheprovider",
                          "--rootdir", os.path.join(ROOT, "tests"), *sys.argv[1:]]))
