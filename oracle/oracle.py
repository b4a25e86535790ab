"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

ctypes bindings to ``oracle/_build/libadrt_oracle.so`` (plain-C restatement,
``oracle/adrt_oracle.c``) plus NumPy restatements of the reference's
Python-level glue (``truncate``, ``stitch_adrt``, ``iadrt_fmg_step`` ...).
Each function cites the reference file:line it follows (paths relative to
/root/reference/src/adrt/).

Parity status: PINNED -- see tests/test_oracle.py (golden vectors generated
from the unmodified reference + live comparison against oracle/_ref).

Only tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libadrt_oracle.so")
_lib = None


def build() -> None:
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _suffix(a: np.ndarray) -> str:
    if a.dtype == np.float32:
        return "f32"
    if a.dtype == np.float64:
        return "f64"
    raise TypeError(f"unsupported array dtype {a.dtype}")


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _call(name, a, *args):
    fn = getattr(_load(), f"{name}_{_suffix(a)}")
    fn.restype = None
    fn(*args)


def num_iters(n: int) -> int:
    """adrt_cdefs_common.cpp:140-142 / core.py:115-120."""
    return n.bit_length() - (bin(n).count("1") == 1)


def _batched(a, nd):
    a = np.ascontiguousarray(a)
    squeeze = a.ndim == nd - 1
    if squeeze:
        a = a[np.newaxis]
    assert a.ndim == nd
    return a, squeeze


def adrt_init(x):
    """core.py:123-176."""
    x, sq = _batched(x, 3)
    B, n, _ = x.shape
    out = np.empty((B, 4, 2 * n - 1, n), dtype=x.dtype)
    _call("oracle_adrt_init", x, _ptr(x), ctypes.c_long(B), ctypes.c_long(n), _ptr(out))
    return out[0] if sq else out


def adrt_step(a, step):
    """adrt_cdefs_adrt.hpp:215-258."""
    a, sq = _batched(a, 4)
    B, _, _, n = a.shape
    out = np.empty_like(a)
    _call("oracle_adrt_step", a, _ptr(a), ctypes.c_long(B), ctypes.c_long(n), ctypes.c_int(step), _ptr(out))
    return out[0] if sq else out


def bdrt_step(a, step):
    """adrt_cdefs_bdrt.hpp:190-244."""
    a, sq = _batched(a, 4)
    B, _, _, n = a.shape
    out = np.empty_like(a)
    _call("oracle_bdrt_step", a, _ptr(a), ctypes.c_long(B), ctypes.c_long(n), ctypes.c_int(step),
          ctypes.c_int(num_iters(n)), _ptr(out))
    return out[0] if sq else out


def adrt(x):
    """adrt_cdefs_adrt.hpp:101-212."""
    x, sq = _batched(x, 3)
    B, n, _ = x.shape
    out = np.empty((B, 4, 2 * n - 1, n), dtype=x.dtype)
    tmp = np.empty_like(out)
    _call("oracle_adrt", x, _ptr(x), ctypes.c_long(B), ctypes.c_long(n), ctypes.c_int(num_iters(n)), _ptr(tmp), _ptr(out))
    return out[0] if sq else out


def _sino_op(name, a):
    a, sq = _batched(a, 4)
    B, _, _, n = a.shape
    out = np.empty_like(a)
    tmp = np.empty_like(a)
    _call(name, a, _ptr(a), ctypes.c_long(B), ctypes.c_long(n), ctypes.c_int(num_iters(n)), _ptr(tmp), _ptr(out))
    return out[0] if sq else out


def bdrt(a):
    """adrt_cdefs_bdrt.hpp:121-187."""
    return _sino_op("oracle_bdrt", a)


def iadrt(a):
    """adrt_cdefs_iadrt.hpp:110-176."""
    return _sino_op("oracle_iadrt", a)


def press_fmg_restriction(a):
    """adrt_cdefs_fmg.hpp:53-73."""
    a, sq = _batched(a, 4)
    B, _, _, n = a.shape
    out = np.empty((B, 4, n - 1, n // 2), dtype=a.dtype)
    _call("oracle_fmg_restriction", a, _ptr(a), ctypes.c_long(B), ctypes.c_long(n), _ptr(out))
    return out[0] if sq else out


def press_fmg_prolongation(a):
    """adrt_cdefs_fmg.hpp:75-95."""
    a, sq = _batched(a, 3)
    B, h, w = a.shape
    out = np.empty((B, 2 * h, 2 * w), dtype=a.dtype)
    _call("oracle_fmg_prolongation", a, _ptr(a), ctypes.c_long(B), ctypes.c_long(h), ctypes.c_long(w), _ptr(out))
    return out[0] if sq else out


def press_fmg_highpass(a):
    """adrt_cdefs_fmg.hpp:97-171."""
    a, sq = _batched(a, 3)
    B, h, w = a.shape
    out = np.empty_like(a)
    _call("oracle_fmg_highpass", a, _ptr(a), ctypes.c_long(B), ctypes.c_long(h), ctypes.c_long(w), _ptr(out))
    return out[0] if sq else out


def interp_to_cart(a):
    """adrt_cdefs_interp_adrtcart.hpp:61-114."""
    a, sq = _batched(a, 4)
    B, _, _, n = a.shape
    out = np.empty((B, n, 4 * n), dtype=a.dtype)
    _call("oracle_interp_to_cart", a, _ptr(a), ctypes.c_long(B), ctypes.c_long(n), _ptr(out))
    return out[0] if sq else out


# ----------------------------------------------------------------------------
# NumPy restatements of the reference's Python-level glue
# ----------------------------------------------------------------------------

def truncate(a):
    """utils.py:231-242."""
    n = a.shape[-1]
    return np.stack(
        [
            np.flip(a[..., 0, :n, :n], axis=-2).swapaxes(-1, -2),
            np.flip(a[..., 1, :n, :n], axis=-2),
            a[..., 2, :n, :n],
            np.flip(a[..., 3, :n, :n], axis=(-1, -2)).swapaxes(-1, -2),
        ],
        axis=-3,
    )


def stitch_adrt(a, remove_repeated=False):
    """utils.py:111-134."""
    n = a.shape[-1]
    in_rows, out_rows = 2 * n - 1, 3 * n - 2
    vc = n - (1 if remove_repeated else 0)
    ret = np.zeros((*a.shape[:-3], out_rows, 4, vc), dtype=a.dtype)
    for i in range(4):
        q = a[..., i, :, :]
        if i % 2:
            q = np.flip(q, axis=(-1, -2))
        if remove_repeated:
            q = q[..., :-1]
        if i < 2:
            ret[..., :in_rows, i, :] = q
        else:
            ret[..., -in_rows:, i, :] = q
    return ret.reshape((*a.shape[:-3], out_rows, 4 * vc))


def iadrt_fmg_step(a):
    """core.py:318-331, with np.mean(axis=-3) spelled out as the sequential
    ((q0+q1)+q2)+q3 then /4 that NumPy performs."""
    stack = []
    for _ in range(num_iters(a.shape[-1])):
        stack.append(a)
        a = press_fmg_restriction(a)
    ret = np.ascontiguousarray(a[..., 0, :, :])
    n = 1
    while stack:
        n *= 2
        ret = press_fmg_prolongation(ret)
        r = adrt(ret) - stack.pop()
        t = truncate(bdrt(r)) / a.dtype.type(n - 1)
        m = ((t[..., 0, :, :] + t[..., 1, :, :]) + t[..., 2, :, :]) + t[..., 3, :, :]
        m = m / a.dtype.type(4)
        ret = ret - press_fmg_highpass(np.ascontiguousarray(m))
    return ret


def iadrt_fmg_iter(a, count):
    """core.py:375-381, first ``count`` iterates."""
    out = []
    inv = iadrt_fmg_step(a)
    out.append(inv)
    for _ in range(count - 1):
        inv = inv + iadrt_fmg_step(a - adrt(inv))
        out.append(inv)
    return out
