"""Shim: `utils.target_mask` of the reference -> aline_b200.utils.target_mask."""
from aline_b200.utils.target_mask import *  # noqa: F401,F403
from aline_b200.utils.target_mask import __dict__ as _d  # noqa: F401
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
