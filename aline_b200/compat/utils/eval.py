"""Shim: `utils.eval` of the reference -> aline_b200.utils.eval."""
from aline_b200.utils.eval import *  # noqa: F401,F403
from aline_b200.utils.eval import __dict__ as _d  # noqa: F401
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
