"""TEST INFRASTRUCTURE ONLY -- loaders for the unmodified reference build.

``oracle/_ref/_adrt_cdefs.abi3.so`` is the reference's own C++/OpenMP core
(/root/reference/src/adrt/adrt_cdefs_py.cpp + adrt_cdefs_common.cpp) compiled
by ``oracle/Makefile`` from the sources where they lie.  It is git-ignored and
travels to the GPU box as a prebuilt file.  Only ``tests/``, ``bench.py``'s
reference / cpu_baseline legs and ``__graft_entry__.smoke()`` may import this
module; the product package ``adrt_b200`` never does.

``load_ref_cdefs()``   -> the native module (works anywhere the .so exists).
``load_ref_package()`` -> the full reference Python package, importable only
                          where /root/reference is mounted (this container);
                          used by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "_adrt_cdefs.abi3.so")
REF_SRC = "/root/reference/src/adrt"

_cdefs = None
_pkg = None


def have_ref_cdefs() -> bool:
    return os.path.exists(REF_SO)


def load_ref_cdefs():
    """Return the reference's native module ``_adrt_cdefs`` (9 functions)."""
    global _cdefs
    if _cdefs is None:
        if not have_ref_cdefs():
            raise FileNotFoundError(
                f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists"
            )
        # The init symbol is PyInit__adrt_cdefs: it depends only on the last
        # component of the module name.
        loader = importlib.machinery.ExtensionFileLoader("adrt_ref._adrt_cdefs", REF_SO)
        spec = importlib.util.spec_from_file_location(
            "adrt_ref._adrt_cdefs", REF_SO, loader=loader
        )
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _cdefs = mod
    return _cdefs


def have_ref_package() -> bool:
    return have_ref_cdefs() and os.path.isdir(REF_SRC)


def load_ref_package():
    """Import the reference Python package as ``adrt_ref`` without copying it.

    The package's ``__path__`` points at the read-only reference sources and
    the native submodule is pre-seeded from ``oracle/_ref``.
    """
    global _pkg
    if _pkg is None:
        if not have_ref_package():
            raise FileNotFoundError("reference sources not mounted at " + REF_SRC)
        cdefs = load_ref_cdefs()
        sys.modules["adrt_ref._adrt_cdefs"] = cdefs
        spec = importlib.util.spec_from_file_location(
            "adrt_ref",
            os.path.join(REF_SRC, "__init__.py"),
            submodule_search_locations=[REF_SRC],
        )
        pkg = importlib.util.module_from_spec(spec)
        sys.modules["adrt_ref"] = pkg
        pkg._adrt_cdefs = cdefs
        spec.loader.exec_module(pkg)
        _pkg = pkg
    return _pkg
