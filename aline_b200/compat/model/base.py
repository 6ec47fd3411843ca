"""Shim: `model.base` of the reference -> aline_b200.model.base."""
from aline_b200.model.base import *  # noqa: F401,F403
from aline_b200.model.base import __dict__ as _d  # noqa: F401
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
