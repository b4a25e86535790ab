"""CPU: the C-ABI library loads and exports every symbol include/adrt_b200.h
declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

from adrt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "adrt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adrt_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert len(syms) >= 35
    for must in ("adrt_b200_adrt", "adrt_b200_bdrt", "adrt_b200_host_adrt", "adrt_b200_interp_to_cart"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_signature_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_pure_helpers():
    lib = _lib.load()
    assert lib.adrt_b200_version() >= 100
    for n, k in ((0, 0), (1, 0), (2, 1), (3, 2), (4, 2), (1024, 10), (2048, 11), (2049, 12)):
        assert lib.adrt_b200_num_iters(n) == k
    assert lib.adrt_b200_launch_count() >= 0
    assert isinstance(_lib.last_error(), str)


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.adrt_b200_adrt(None, None, 1, 8, 0, None, 0, None) == 1
    buf = ctypes.create_string_buffer(128)
    p = ctypes.addressof(buf)
    o = p + 64
    assert lib.adrt_b200_adrt(p, o, 1, 12, 0, None, 0, None) == 1
    assert "power of two" in _lib.last_error()
    assert lib.adrt_b200_adrt(p, o, 0, 8, 0, None, 0, None) == 1
    assert lib.adrt_b200_adrt(p, o, 1, 8, 7, None, 0, None) == 1
    assert lib.adrt_b200_adrt_step(p, o, 1, 8, 3, 0, None) == 1
    assert lib.adrt_b200_fmg_highpass(p, o, 1, 1, 8, 0, None) == 1
    # no transform runs in place: in == out is an argument error for every one of them
    for fn in (lib.adrt_b200_adrt, lib.adrt_b200_bdrt, lib.adrt_b200_iadrt, lib.adrt_b200_bdrt_planes):
        assert fn(p, p, 1, 8, 0, None, 0, None) == 1
        assert "alias" in _lib.last_error()
    assert lib.adrt_b200_bdrt_rows(p, p, 1, 8, 8, 0, None, 0, None) == 1
    assert lib.adrt_b200_adrt_step(p, p, 1, 8, 0, 0, None) == 1
    assert lib.adrt_b200_bdrt_step(p, p, 1, 8, 0, 0, None) == 1


def test_no_gpu_fails_loudly():
    import numpy as np

    import adrt_b200

    if _lib.load().adrt_b200_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.ADRTB200Error):
        adrt_b200.adrt(np.zeros((4, 4), dtype=np.float32))
