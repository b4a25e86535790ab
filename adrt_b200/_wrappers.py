"""Array normalisation + thin public wrappers over ``adrt_b200._adrt_cdefs``.

Mirrors the role of the reference's ``adrt/_wrappers.py`` (normalise, then one
call into the native module; :96-112 and :115-390 there).  CUDA tensors pass
through untouched apart from ``.contiguous()``.
"""
from __future__ import annotations

import operator

import numpy as np

from . import _adrt_cdefs

__all__: list[str] = []


def _format_object_type(obj, /) -> str:
    t = type(obj)
    return t.__qualname__ if t.__module__ == "builtins" else f"{t.__module__}.{t.__qualname__}"


def _normalize_array(a, /):
    """Return `a` in a layout the native boundary accepts.

    NumPy: native byte order, C order, aligned (copying only when needed), as
    in the reference (_wrappers.py:96-112).  CUDA tensors: made contiguous.
    Anything else: ``TypeError``.
    """
    if isinstance(a, np.ndarray):
        a = np.asarray(a, a.dtype.newbyteorder("="), "C")
        return a if a.flags.aligned else a.copy("C")
    if (not _adrt_cdefs._is_torch_tensor(a) and hasattr(a, "__dlpack__") and hasattr(a, "__dlpack_device__")
            and int(a.__dlpack_device__()[0]) == 2):  # kDLCUDA producers (CuPy, ...): zero-copy import
        import torch

        a = torch.from_dlpack(a)
    if _adrt_cdefs._is_torch_tensor(a) and a.is_cuda:
        return a.contiguous()
    raise TypeError(f"array must be numpy.ndarray, but got {_format_object_type(a)}")


def _public(module):
    def deco(fn):
        fn.__module__ = module
        return fn
    return deco


@_public("adrt_b200")
def adrt(a, /, *, out=None):
    """Approximate discrete Radon transform of square power-of-two image(s).

    ``(B?, n, n)`` float32/float64 -> ``(B?, 4, 2n-1, n)``; bit-identical to
    the reference's ``adrt.adrt``.
    """
    return _adrt_cdefs.adrt(_normalize_array(a), out=out)


@_public("adrt_b200.core")
def adrt_step(a, /, step, *, out=None):
    """One butterfly stage ``step`` (``0 <= step < num_iters(n)``) of the ADRT
    on an array produced by :func:`adrt_b200.core.adrt_init`."""
    return _adrt_cdefs.adrt_step(_normalize_array(a), operator.index(step), out=out)


@_public("adrt_b200")
def iadrt(a, /, *, out=None):
    """Exact (ill-conditioned) single-quadrant inverse; combine the quadrants
    of the result with :func:`adrt_b200.utils.truncate` and a mean."""
    return _adrt_cdefs.iadrt(_normalize_array(a), out=out)


@_public("adrt_b200")
def bdrt(a, /, *, out=None):
    """Back-projection (per-quadrant transpose of :func:`adrt`); same shape in
    and out.  ``truncate(bdrt(y)).sum(-3)`` is exactly ``adrt``-transpose."""
    return _adrt_cdefs.bdrt(_normalize_array(a), out=out)


@_public("adrt_b200.core")
def bdrt_step(a, /, step, *, out=None):
    """One stage ``step`` of the back-projection."""
    return _adrt_cdefs.bdrt_step(_normalize_array(a), operator.index(step), out=out)


@_public("adrt_b200.utils")
def interp_to_cart(a, /, *, out=None):
    """Nearest-neighbour resampling of an ADRT output onto a regular
    ``(t, theta)`` grid: ``(B?, 4, 2n-1, n)`` -> ``(B?, n, 4n)``."""
    return _adrt_cdefs.interp_to_cart(_normalize_array(a), out=out)


@_public("adrt_b200.core")
def threading_enabled() -> bool:
    """Whether the core runs multithreaded (always true on the GPU engine)."""
    return _adrt_cdefs.OPENMP_ENABLED


def _press_fmg_restriction(a, /):
    return _adrt_cdefs.press_fmg_restriction(_normalize_array(a))


def _press_fmg_prolongation(a, /):
    return _adrt_cdefs.press_fmg_prolongation(_normalize_array(a))


def _press_fmg_highpass(a, /):
    return _adrt_cdefs.press_fmg_highpass(_normalize_array(a))
