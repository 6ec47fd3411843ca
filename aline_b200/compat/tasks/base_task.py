"""Shim: `tasks.base_task` of the reference -> aline_b200.tasks.base_task."""
from aline_b200.tasks.base_task import *  # noqa: F401,F403
from aline_b200.tasks.base_task import __dict__ as _d  # noqa: F401
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
