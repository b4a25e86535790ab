"""``adrt`` -- the reference's import name, served by the B200 engine.

``import adrt`` on a machine with this repository on ``sys.path`` gives the
package tree the reference ships (/root/reference/src/adrt/__init__.py:52-63):
``adrt.adrt / iadrt / bdrt / iadrt_fmg``, ``adrt.core``, ``adrt.utils`` and the
private modules its own test-suite reaches into, ``adrt._adrt_cdefs`` (the native
module, tests/test_adrt.py:85-166) and ``adrt._wrappers`` (``_press_fmg_*``,
tests/test_press_fmg_highpass.py:56).  Every one of them *is* the corresponding
``adrt_b200`` module object -- no second copy of any function exists, so
``monkeypatch.setattr(adrt.core, "iadrt_fmg_step", ...)``
(tests/test_iadrt_fmg.py:42-55) patches the module the engine itself calls
through.  There is no CPU implementation behind this name either.
"""
import sys as _sys

import adrt_b200 as _impl
from adrt_b200 import _adrt_cdefs, _wrappers, core, utils
from adrt_b200 import adrt, bdrt, iadrt, iadrt_fmg

__all__ = ["adrt", "iadrt", "bdrt", "iadrt_fmg", "utils", "core"]
__version__ = _impl.__version__

# `import adrt.core`, `from adrt.utils import ...`, `import adrt._adrt_cdefs` resolve to
# the same module objects
for _name, _mod in (("core", core), ("utils", utils), ("_adrt_cdefs", _adrt_cdefs), ("_wrappers", _wrappers)):
    _sys.modules[__name__ + "." + _name] = _mod
del _name, _mod, _sys
