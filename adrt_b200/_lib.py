"""ctypes binding of ``libadrt_b200.so`` (the C ABI in ``include/adrt_b200.h``).

The library is built in-tree by ``make -C adrt_b200/csrc`` (or
``__graft_entry__.build()``).  There is deliberately no fallback: if the
shared object is missing, or there is no CUDA device when a compute function
is called, this module raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# ADRT_B200_LIB names another build of the same library (A/B tuning of compile-time options)
LIB_PATH = os.environ.get("ADRT_B200_LIB") or os.path.join(_HERE, "libadrt_b200.so")

F32, F64 = 0, 1

_c_i64 = ctypes.c_int64
_c_vp = ctypes.c_void_p
_c_int = ctypes.c_int
_c_sz = ctypes.c_size_t

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks
# against include/adrt_b200.h.
SIGNATURES = {
    "adrt_b200_version": (_c_int, []),
    "adrt_b200_last_error": (ctypes.c_char_p, []),
    "adrt_b200_device_count": (_c_int, []),
    "adrt_b200_set_mode": (None, [_c_int]),
    "adrt_b200_get_mode": (_c_int, []),
    "adrt_b200_launch_count": (_c_i64, []),
    "adrt_b200_num_iters": (_c_int, [_c_i64]),
    "adrt_b200_adrt_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int]),
    "adrt_b200_adrt": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_bdrt_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int]),
    "adrt_b200_bdrt": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_adrt_quadrants_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_adrt_quadrants": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_bdrt_planes_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int]),
    "adrt_b200_bdrt_planes": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_bdrt_rows": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_normal_operator_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int]),
    "adrt_b200_normal_operator": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, ctypes.c_double, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_adrt_bdrt_rows_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_adrt_bdrt_rows": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_part_exchange_pitch": (_c_sz, [_c_i64, _c_int, _c_int, _c_int]),
    "adrt_b200_part_exchange_cols": (_c_sz, [_c_i64, _c_int, _c_int, _c_i64]),
    "adrt_b200_part_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_adrt_part": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_bdrt_part": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_adrt_step": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_vp]),
    "adrt_b200_bdrt_step": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_vp]),
    "adrt_b200_adrt_init": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_vp]),
    "adrt_b200_iadrt_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int]),
    "adrt_b200_iadrt": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_fmg_restriction": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_vp]),
    "adrt_b200_fmg_prolongation": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_int, _c_vp]),
    "adrt_b200_fmg_highpass": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_int, _c_vp]),
    "adrt_b200_fmg_step_workspace_bytes": (_c_sz, [_c_i64, _c_i64, _c_int]),
    "adrt_b200_fmg_step": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_interp_to_cart": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_vp]),
    "adrt_b200_truncate": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_vp]),
    "adrt_b200_stitch": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_vp]),
    "adrt_b200_unstitch": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_vp]),
    "adrt_b200_truncate_mean": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, ctypes.c_double, _c_int, _c_vp]),
    "adrt_b200_truncate_mean_shares": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, ctypes.c_double, _c_int, _c_vp]),
    "adrt_b200_sub": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_vp]),
    "adrt_b200_add": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_vp]),
    "adrt_b200_cg_workspace_bytes": (_c_sz, []),
    "adrt_b200_cg_dot": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_cg_update": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_vp, _c_sz, _c_vp]),
    "adrt_b200_cg_direction": (_c_int, [_c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_vp]),
    "adrt_b200_host_adrt": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_host_bdrt": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_host_iadrt": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_host_adrt_step": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_int]),
    "adrt_b200_host_bdrt_step": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int, _c_int]),
    "adrt_b200_host_adrt_init": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_host_fmg_restriction": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_host_fmg_prolongation": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_host_fmg_highpass": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_host_interp_to_cart": (_c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_int, _c_int]),
    "adrt_b200_host_alloc_pinned": (_c_vp, [_c_sz]),
    "adrt_b200_host_free_pinned": (None, [_c_vp]),
    "adrt_b200_host_is_pinned": (_c_int, [_c_vp]),
}

_lib = None
_lock = threading.Lock()


class ADRTB200Error(RuntimeError):
    """A call into libadrt_b200.so failed (CUDA error, allocation failure...)."""


def load() -> ctypes.CDLL:
    """Load the shared library (once) and attach signatures."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} not found: build it with `make -C adrt_b200/csrc` "
                        "(adrt_b200 has no CPU fallback)"
                    )
                lib = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(lib, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = lib
    return _lib


def last_error() -> str:
    return load().adrt_b200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = last_error()
        if rc == 4:
            raise MemoryError(f"{what}: {msg}")
        raise ADRTB200Error(f"{what} failed (status {rc}): {msg}")


_device_checked = False


def require_device() -> None:
    """Fail loudly when there is no GPU: the product path never runs on CPU."""
    global _device_checked
    if not _device_checked:
        n = load().adrt_b200_device_count()
        if n <= 0:
            raise ADRTB200Error(
                "adrt_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback"
            )
        _device_checked = True
