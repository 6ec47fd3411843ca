"""Shim: the reference's `loss` package resolves to aline_b200.loss (see ../README.md)."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _ROOT not in _sys.path:          # make the aline_b200 package importable when only compat/ was put on the path
    _sys.path.append(_ROOT)
from aline_b200.loss import *  # noqa: F401,F403,E402
