"""adrt_b200 -- B200-native Approximate Discrete Radon Transform.

A drop-in for the Python API of karlotness/adrt (``adrt``, ``bdrt``, ``iadrt``,
``iadrt_fmg``, ``core``, ``utils``) whose numerical core is hand-written
sm_100a CUDA behind the C ABI in ``include/adrt_b200.h``.  NumPy arrays and
CUDA tensors are accepted; there is no CPU fallback.

    import adrt_b200 as adrt
    y = adrt.adrt(x)            # (B?, n, n) -> (B?, 4, 2n-1, n)
    z = adrt.bdrt(y)
"""
from __future__ import annotations

import itertools

import numpy as np

from . import core, utils
from ._adrt_cdefs import get_device, set_device
from ._wrappers import adrt, bdrt, iadrt

__all__ = ["adrt", "iadrt", "bdrt", "iadrt_fmg", "utils", "core", "set_device", "get_device"]
__version__ = "1.2.1.dev"  # API level of the reference this mirrors (adrt/__init__.py:63)


def iadrt_fmg(a, /, *, max_iters=None):
    """Approximate inverse by repeated full-multigrid refinement.

    Iterates :func:`adrt_b200.core.iadrt_fmg_iter` and returns the iterate
    after which the residual norm ``||adrt(x) - a||`` stopped decreasing (or
    the ``max_iters``-th).  No batch dimension; returns a writable ``(n, n)``
    array (adrt/__init__.py:122-148).
    """
    if a.ndim > 3:
        raise ValueError(f"batch dimension not supported for iadrt_fmg, got {a.ndim} dimensions")
    if max_iters is not None and max_iters < 1:
        raise ValueError(f"must allow at least one iteration, but specified {max_iters}")
    from . import _adrt_cdefs as cd
    from ._wrappers import _normalize_array

    a = _normalize_array(a)
    as_numpy = isinstance(a, np.ndarray)
    if as_numpy:
        core._check_fmg_input(a)
        dev = core._to_device(a)
    else:
        dev = a

    def residual(x):
        # the reference's float(np.linalg.norm(adrt(x) - a)) (adrt/__init__.py:135).  The
        # subtraction runs on the device (elementwise IEEE, same bits as NumPy's); for NumPy
        # callers the norm itself is NumPy's, so that the stopping rule `res2 < res1` sees the
        # very numbers the reference would and selects the same iterate.  CUDA-tensor callers
        # (an extension) get a device reduction: same value up to summation order.
        r = cd.sub(cd.adrt(x), dev)
        if as_numpy:
            return float(np.linalg.norm(r.cpu().numpy()))
        return float(r.reshape(-1).norm())

    best = None
    pairs = ((x, residual(x)) for x in itertools.islice(core.iadrt_fmg_iter(dev, copy=False), max_iters))
    for (best, res1), (_, res2) in itertools.pairwise(itertools.chain(pairs, [(None, np.inf)])):
        if not res2 < res1:
            break
    return best.cpu().numpy().copy() if as_numpy else best.clone()
