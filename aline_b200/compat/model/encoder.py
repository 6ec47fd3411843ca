"""Shim: `model.encoder` of the reference -> aline_b200.model.encoder."""
from aline_b200.model.encoder import *  # noqa: F401,F403
from aline_b200.model.encoder import __dict__ as _d  # noqa: F401
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
