"""Shim: `utils.misc` of the reference -> aline_b200.utils.misc."""
from aline_b200.utils.misc import *  # noqa: F401,F403
from aline_b200.utils.misc import __dict__ as _d  # noqa: F401
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
