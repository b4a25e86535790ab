"""Layout utilities (drop-in for the reference's ``adrt.utils``).

``stitch_adrt``, ``unstitch_adrt``, ``truncate``, ``coord_adrt``,
``coord_cart_to_adrt`` follow /root/reference/src/adrt/utils.py:65-435 and
work on NumPy arrays with any number of leading dims.  ``stitch_adrt``,
``unstitch_adrt`` and ``truncate`` also take CUDA tensors: float32/float64
tensors go through the gather kernels ``adrt_b200_stitch`` /
``adrt_b200_unstitch`` / ``adrt_b200_truncate`` (one pass over the data,
SURVEY 8f rank 3), other dtypes through torch indexing, so nothing leaves
the GPU.  ``coord_adrt`` tables are computed once per ``n`` (and once per
device with ``device=``) and handed out as copies.
``interp_to_cart`` is the native gather (adrt_b200_interp_to_cart).
"""
from __future__ import annotations

import functools
import operator
import typing

import numpy as np

from . import _adrt_cdefs
from ._wrappers import interp_to_cart

__all__ = [
    "stitch_adrt",
    "unstitch_adrt",
    "truncate",
    "coord_adrt",
    "coord_cart_to_adrt",
    "interp_to_cart",
]


def _is_tensor(a) -> bool:
    return _adrt_cdefs._is_torch_tensor(a)


def _flip(a, axes):
    return a.flip(axes) if _is_tensor(a) else np.flip(a, axis=axes)


def _swap(a):
    return a.transpose(-1, -2) if _is_tensor(a) else a.swapaxes(-1, -2)


def _device_float(a) -> bool:
    """CUDA float32 / float64 tensor: the dtypes the gather kernels are instantiated for."""
    if not (_is_tensor(a) and a.is_cuda):
        return False
    import torch

    return a.dtype in (torch.float32, torch.float64)


def _device_gather(name, a, lead, in_tail, out_tail, n, flag):
    """Run adrt_b200_<name>(in, out, B, n, flag, dtype, stream) on a CUDA tensor whose last
    dims are `in_tail`; leading dims are flattened into the batch."""
    import torch

    from . import _lib

    batch = 1
    for s in lead:
        batch *= int(s)
    out = torch.empty((*lead, *out_tail), dtype=a.dtype, device=a.device)
    if batch == 0 or out.numel() == 0:
        return out
    src = a.contiguous()
    lib = _lib.load()
    code = _lib.F32 if a.dtype == torch.float32 else _lib.F64
    with torch.cuda.device(a.device):
        rc = getattr(lib, f"adrt_b200_{name}")(src.data_ptr(), out.data_ptr(), batch, n, int(flag), code,
                                                torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, name)
    return out


def _stack(parts, axis):
    if _is_tensor(parts[0]):
        import torch

        return torch.stack(parts, dim=axis)
    return np.stack(parts, axis=axis)


def stitch_adrt(a, /, *, remove_repeated=False):
    """Lay the four quadrants side by side as one ``(..., 3n-2, 4n)`` image
    (``4n-4`` wide with ``remove_repeated``): quadrant ``i`` fills column band
    ``i``; bands 0,1 start at row 0, bands 2,3 end at the last row; odd
    quadrants are flipped on both axes (utils.py:111-134)."""
    n = a.shape[-1]
    if tuple(a.shape[-3:]) != (4, 2 * n - 1, n):
        raise ValueError(f"unsuitable shape for ADRT output processing {tuple(a.shape)}")
    rows_in, rows_out = 2 * n - 1, 3 * n - 2
    width = n - (1 if remove_repeated else 0)
    lead = tuple(a.shape[:-3])
    if _device_float(a) and n <= 16384:
        return _device_gather("stitch", a, lead, (4, rows_in, n), (rows_out, 4 * width), n, remove_repeated)
    if _is_tensor(a):
        canvas = a.new_zeros((*lead, rows_out, 4, width))
    else:
        canvas = np.zeros_like(a, shape=(*lead, rows_out, 4, width), order="C")
    for i in range(4):
        quad = a[..., i, :, :]
        if i % 2:
            quad = _flip(quad, (-1, -2))
        if remove_repeated:
            quad = quad[..., :-1]
        if i < 2:
            canvas[..., :rows_in, i, :] = quad
        else:
            canvas[..., rows_out - rows_in:, i, :] = quad
    return canvas.reshape((*lead, rows_out, 4 * width))


def unstitch_adrt(a, /):
    """Inverse of :func:`stitch_adrt` for either width (utils.py:162-188)."""
    n = (a.shape[-2] + 2) // 3
    if a.shape[-2] != 3 * n - 2 or (a.shape[-1] != 4 * n and a.shape[-1] != 4 * n - 4):
        raise ValueError(f"unsuitable shape for ADRT unstitching {tuple(a.shape)}")
    trimmed = a.shape[-1] == 4 * n - 4
    rows = 2 * n - 1
    if _device_float(a) and n <= 16384 and a.shape[-1] > 0:
        return _device_gather("unstitch", a, tuple(a.shape[:-2]), tuple(a.shape[-2:]), (4, rows, n), n, trimmed)
    a = a.reshape((*a.shape[:-1], 4, n - (1 if trimmed else 0)))
    quads = []
    for q in range(4):
        quad = a[..., :, q, :]
        if trimmed:
            # the dropped column equals the first column of the next band
            nb = a[..., :, (q + 1) % 4, 0:1]
            if q == 3:
                nb = _flip(nb, (-2,))
            if _is_tensor(a):
                import torch

                quad = torch.cat([quad, nb], dim=-1)
            else:
                quad = np.concatenate([quad, nb], axis=-1)
        quad = quad[..., :rows, :] if q < 2 else quad[..., quad.shape[-2] - rows:, :]
        if q % 2:
            quad = _flip(quad, (-1, -2))
        quads.append(quad)
    return _stack(quads, -3)


def truncate(a, /):
    """Cut each quadrant to its top ``n x n`` square and undo ``adrt_init``'s
    orientation: ``(..., 4, 2n-1, n)`` -> ``(..., 4, n, n)`` (utils.py:231-242)."""
    n = a.shape[-1]
    if tuple(a.shape[-3:]) != (4, 2 * n - 1, n):
        raise ValueError(f"unsuitable shape for ADRT output processing {tuple(a.shape)}")
    if _device_float(a) and n <= 16384:
        return _adrt_cdefs.truncate(a.reshape((-1, 4, 2 * n - 1, n))).reshape((*a.shape[:-3], 4, n, n))
    return _stack(
        [
            _swap(_flip(a[..., 0, :n, :n], (-2,))),
            _flip(a[..., 1, :n, :n], (-2,)),
            a[..., 2, :n, :n],
            _swap(_flip(a[..., 3, :n, :n], (-1, -2))),
        ],
        -3,
    )


class ADRTCoord(typing.NamedTuple):
    offset: np.ndarray
    angle: np.ndarray


def _check_domain_size(n) -> int:
    n = operator.index(n)
    if n < 2:
        raise ValueError(f"invalid Radon domain size {n}, must be at least 2")
    if n.bit_count() != 1:
        raise ValueError(f"invalid Radon domain size {n}, must be a power of two")
    return n


def coord_adrt(n, /, *, device=None) -> ADRTCoord:
    """Radon-domain coordinates of every ADRT entry: ``offset`` ``(4, 2n-1, n)``
    and ``angle`` ``(4, 1, n)`` in float64 (utils.py:304-325).

    The tables depend on ``n`` only: they are computed once per ``n`` and every call gets
    its own writable copy (the reference's contract).  ``device=`` (an extension) returns
    CUDA tensors instead, uploaded once per ``(n, device)`` and shared read-only."""
    n = _check_domain_size(n)
    if device is not None:
        return _coord_adrt_device(n, str(device))
    base = _coord_adrt_table(n)
    return ADRTCoord(base.offset.copy(), base.angle.copy())


@functools.lru_cache(maxsize=8)
def _coord_adrt_device(n: int, device: str) -> ADRTCoord:
    import torch

    base = _coord_adrt_table(n)
    return ADRTCoord(torch.from_numpy(base.offset).to(device), torch.from_numpy(base.angle).to(device))


@functools.lru_cache(maxsize=8)
def _coord_adrt_table(n: int) -> ADRTCoord:
    heights, step = np.linspace(1, (1 - n) / n, num=2 * n - 1, endpoint=False, retstep=True, dtype=np.float64)
    heights += step / 2
    slope = np.linspace(0, 1, num=n, endpoint=True, dtype=np.float64)
    theta = np.arctan(slope)
    theta_off = theta - (np.pi / 2)
    h0 = ((np.add.outer(heights, ((2 * n - 1) / (2 * n)) * slope) / (1 + slope)) - 0.5) * (
        np.cos(theta) + np.sin(theta)
    )
    offsets = np.tile(np.stack([h0, -h0], axis=0), (2, 1, 1))
    angles = np.expand_dims(np.stack([theta_off, -theta, theta, -theta_off], axis=0), axis=1)
    offsets.setflags(write=False)
    angles.setflags(write=False)
    return ADRTCoord(offsets, angles)


class ADRTIndex(typing.NamedTuple):
    quadrant: np.ndarray
    height: np.ndarray
    slope: np.ndarray
    factor: np.ndarray


def coord_cart_to_adrt(theta, t, n) -> ADRTIndex:
    """Nearest ADRT index ``(quadrant, height, slope)`` and scale ``factor`` for
    continuous Radon points ``(theta, t)`` (utils.py:407-435)."""
    n = _check_domain_size(n)
    if theta.shape != t.shape:
        raise ValueError(f"mismatched shapes for theta and t {theta.shape} vs. {t.shape}")
    half_pi = np.pi / 2
    theta = np.where(np.abs(theta) <= half_pi, theta, np.remainder(theta + half_pi, np.pi) - half_pi)
    q = np.floor(np.clip(theta / (np.pi / 4), -2, 1)).astype(np.int8) + 2
    th0 = np.pi / 4 - np.abs(np.abs(theta) - np.pi / 4)
    si = np.around(np.tan(th0) * (n - 1)).astype(np.uint64)
    factor = np.sqrt(1 + (si / (n - 1)) ** 2)
    sgn = 2 * (q % 2) - 1
    h = (0.5 * (1 + np.tan(th0)) + (sgn * t) / np.cos(th0)) * n
    hi = (np.round(2 * h).astype(np.int64) - 1) // 2
    return ADRTIndex(q.astype(np.uint8), hi, si, factor)
