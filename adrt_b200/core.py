"""Step-level routines and generators (drop-in for the reference's ``adrt.core``).

``num_iters``, ``adrt_init``, ``adrt_step``/``adrt_iter``,
``bdrt_step``/``bdrt_iter``, ``iadrt_fmg_step``/``iadrt_fmg_iter`` and
``threading_enabled`` have the reference's signatures and semantics
(/root/reference/src/adrt/core.py:94-381).  The multigrid routines keep every
intermediate on the GPU between native calls; the reference bounces through
NumPy at each level (core.py:318-331).
"""
from __future__ import annotations

import operator

import numpy as np

from . import _adrt_cdefs
from ._wrappers import (
    _format_object_type,
    _normalize_array,
    adrt_step,
    bdrt_step,
    threading_enabled,
)

__all__ = [
    "num_iters",
    "adrt_step",
    "adrt_init",
    "adrt_iter",
    "bdrt_step",
    "bdrt_iter",
    "threading_enabled",
    "iadrt_fmg_step",
    "iadrt_fmg_iter",
]


def num_iters(n, /) -> int:
    """``ceil(log2(n))`` butterfly stages for an ``n x n`` image; ``num_iters(0) == 0``
    (core.py:115-120, adrt_cdefs_common.cpp:140-142)."""
    n = operator.index(n)
    if n < 0:
        raise ValueError(f"non-negative value required for iteration count, but got {n}")
    return n.bit_length() - (n.bit_count() == 1)


def _is_tensor(a) -> bool:
    return _adrt_cdefs._is_torch_tensor(a)


def adrt_init(a, /):
    """Stack the four oriented copies of the image(s) into an ADRT-shaped array.

    ``(B?, n, n)`` of any dtype -> ``(B?, 4, 2n-1, n)``, rows ``n..2n-2`` zero
    (core.py:123-176).  Float CUDA tensors are initialised by a device kernel.
    """
    is_t = _is_tensor(a)
    if not (isinstance(a, np.ndarray) or is_t):
        raise TypeError(f"array must be numpy.ndarray, but got {_format_object_type(a)}")
    if a.ndim > 3 or a.ndim < 2:
        raise ValueError(f"array must have between 2 and 3 dimensions, but had {a.ndim}")
    shape = tuple(a.shape)
    if shape[-1] != shape[-2] or operator.index(shape[-1]).bit_count() != 1:
        raise ValueError("array must be square with a power of two shape")
    if not all(shape):
        raise ValueError(
            f"all array dimensions must be nonzero, but found zero in dimension {shape.index(0)}"
        )
    n = shape[-1]
    if is_t:
        import torch

        if a.is_cuda and a.dtype in (torch.float32, torch.float64):
            return _adrt_cdefs.adrt_init(a.contiguous())
        ret = a.new_zeros((*shape[:-2], 4, 2 * n - 1, n))
        ret[..., 0, :n, :] = a.flip(-1).transpose(-1, -2)
        ret[..., 1, :n, :] = a.flip(-2)
        ret[..., 2, :n, :] = a
        ret[..., 3, :n, :] = a.flip((-1, -2)).transpose(-1, -2)
        return ret
    ret = np.zeros_like(a, shape=(*shape[:-2], 4, 2 * n - 1, n))
    ret[..., 0, :n, :] = np.flip(a, axis=-1).swapaxes(-1, -2)
    ret[..., 1, :n, :] = np.flip(a, axis=-2)
    ret[..., 2, :n, :] = a
    ret[..., 3, :n, :] = np.flip(a, axis=(-1, -2)).swapaxes(-1, -2)
    return ret


def _snapshot(a, copy):
    if _is_tensor(a):
        return a.clone() if copy else a
    a.setflags(write=False)
    return a.copy() if copy else a.view()


def adrt_iter(a, /, *, copy=True):
    """Yield ``adrt_init(a)`` and then the array after every ADRT stage
    (``num_iters(n) + 1`` items; core.py:217-223)."""
    a = adrt_init(a)
    yield _snapshot(a, copy)
    for i in range(num_iters(a.shape[-1])):
        a = adrt_step(a, i)
        yield _snapshot(a, copy)


def bdrt_iter(a, /, *, copy=True):
    """Yield the array after every back-projection stage (``num_iters(n)``
    items; core.py:258-262)."""
    for i in range(num_iters(a.shape[-1])):
        a = bdrt_step(a, i)
        yield _snapshot(a, copy)


# ---------------------------------------------------------------------------
# Press full-multigrid inverse, device resident
# ---------------------------------------------------------------------------

def _to_device(a):
    """NumPy -> CUDA tensor on the package's device (one H2D copy)."""
    import torch

    from . import _lib

    _lib.require_device()
    return torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{_adrt_cdefs.get_device()}")


def _fmg_step_device(a):
    """One FMG pass on a CUDA tensor ``(B?, 4, 2n-1, n)`` -> ``(B?, n, n)``.

    Same operator sequence as core.py:318-331, run level by level inside one
    native call (``adrt_b200_fmg_step``): ``np.mean(x / (m-1), axis=-3)`` is the
    ``truncate_mean`` kernel, which divides each quadrant first and then sums
    ``((q0+q1)+q2)+q3`` before the ``/4`` exactly as NumPy does, and the
    back-projection only computes the offsets ``truncate`` keeps.
    """
    return _adrt_cdefs.fmg_step(a)


def iadrt_fmg_step(a, /):
    """Estimated inverse of ``adrt`` by one full-multigrid pass (Press 2006).

    ``(B?, 4, 2n-1, n)`` float -> ``(B?, n, n)`` (core.py:265-331).  NumPy in,
    NumPy out; CUDA tensor in, CUDA tensor out.
    """
    a = _normalize_array(a)
    if _is_tensor(a):
        return _fmg_step_device(a)
    _check_fmg_input(a)
    return _fmg_step_device(_to_device(a)).cpu().numpy()


def _check_fmg_input(a):
    # run the native boundary's validation (shape / dtype messages) up front
    arr = _adrt_cdefs._extract_array(a)
    shape = _adrt_cdefs._array_shape(arr, 3, 4)
    if not _adrt_cdefs._is_adrt_output_shape(shape):
        raise ValueError("array must have a valid ADRT output shape")
    _adrt_cdefs._dtype_code(arr)


def iadrt_fmg_iter(a, /, *, copy=True):
    """Infinite generator of successively refined FMG inverses:
    ``x0 = step(a)``, ``x_{k+1} = x_k + step(a - adrt(x_k))`` (core.py:375-381)."""
    a = _normalize_array(a)
    as_numpy = not _is_tensor(a)
    if as_numpy:
        _check_fmg_input(a)
        a = _to_device(a)
    cd = _adrt_cdefs

    def emit(x):
        if as_numpy:
            h = x.cpu().numpy()
            h.setflags(write=False)
            return h.copy() if copy else h.view()
        return x.clone() if copy else x

    # the public iadrt_fmg_step is looked up at call time, as in the reference (core.py:375-381):
    # its tests count the calls by patching adrt.core.iadrt_fmg_step (tests/test_iadrt_fmg.py:42-55)
    inv = iadrt_fmg_step(a)
    yield emit(inv)
    while True:
        inv = cd.add(inv, iadrt_fmg_step(cd.sub(a, cd.adrt(inv))))
        yield emit(inv)
