"""Pinned result arrays for the NumPy path.

The reference returns a fresh ndarray from every call (adrt_cdefs_py.cpp:177).  Done with
``np.empty`` that costs this engine twice: the device-to-host copy has to go through a
pinned bounce buffer plus a host memcpy, and the fresh pages are faulted in one by one --
the measured difference between 0.25 and 0.77 Gpixel/s end to end on 64 x 2048^2 images
(bench.py ``e2e_default`` vs ``e2e``).  So large results are still fresh, writable,
C-contiguous arrays that belong to the caller, but their memory comes from a small pool
of page-locked blocks (``cudaHostAlloc`` through ``adrt_b200_host_alloc_pinned``): the GPU
writes them directly, and a result that is fed back into the next call (``bdrt(adrt(x))``)
is read directly too.  A block returns to the pool when the last view of its array dies
(``weakref.finalize``); the pool keeps at most ``ADRT_B200_PINNED_POOL_MB`` (default: a
quarter of the host's memory divided by $LOCAL_WORLD_SIZE, at most 64 GiB) and frees the rest.
``ADRT_B200_PINNED_RESULTS=0`` turns the feature off (plain ``np.empty``, ``owndata`` true).
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref

import numpy as np

from . import _lib

_THRESHOLD = 32 << 20          # results smaller than this stay ordinary np.empty arrays
_ROUND = 2 << 20               # block sizes are multiples of 2 MiB so that similar shapes share blocks

_lock = threading.Lock()
_free: dict[int, list[int]] = {}
_cached = 0
_cap = None


def enabled() -> bool:
    return os.environ.get("ADRT_B200_PINNED_RESULTS", "1") != "0"


def _pool_cap() -> int:
    global _cap
    if _cap is None:
        env = os.environ.get("ADRT_B200_PINNED_POOL_MB")
        if env is not None:
            _cap = max(0, int(env)) << 20
        else:
            try:
                total = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
            except (ValueError, OSError):
                total = 16 << 30
            # one process per GPU (torchrun): the ranks of a node share the quarter
            ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
            _cap = min(total // 4 // ranks, 64 << 30)
    return _cap


def _release(ptr: int, size: int) -> None:
    global _cached
    with _lock:
        if _cached + size <= _pool_cap():
            _free.setdefault(size, []).append(ptr)
            _cached += size
            return
    try:
        _lib.load().adrt_b200_host_free_pinned(ctypes.c_void_p(ptr))
    except Exception:  # interpreter shutdown: the OS reclaims the block
        pass


def empty(shape, dtype):
    """A fresh writable C-contiguous ndarray in page-locked memory, or None when the
    result is small, the feature is off or the allocation fails (callers then use
    ``np.empty``)."""
    global _cached
    dtype = np.dtype(dtype)
    count = 1
    for s in shape:
        count *= int(s)
    nbytes = count * dtype.itemsize
    if nbytes < _THRESHOLD or not enabled():
        return None
    size = (nbytes + _ROUND - 1) // _ROUND * _ROUND
    ptr = None
    with _lock:
        bucket = _free.get(size)
        if bucket:
            ptr = bucket.pop()
            _cached -= size
    if ptr is None:
        ptr = _lib.load().adrt_b200_host_alloc_pinned(size)
        if not ptr:
            return None
    buf = (ctypes.c_char * size).from_address(ptr)
    weakref.finalize(buf, _release, ptr, size)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def trim() -> None:
    """Free every cached block (e.g. before handing the host memory to something else)."""
    global _cached
    with _lock:
        blocks = [(p, s) for s, ps in _free.items() for p in ps]
        _free.clear()
        _cached = 0
    lib = _lib.load()
    for p, _ in blocks:
        lib.adrt_b200_host_free_pinned(ctypes.c_void_p(p))
