"""Shim: `tasks.psychometric` of the reference -> aline_b200.tasks.psychometric."""
from aline_b200.tasks.psychometric import *  # noqa: F401,F403
from aline_b200.tasks.psychometric import __dict__ as _d  # noqa: F401
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
