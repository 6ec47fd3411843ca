"""Shim: `tasks.location_finding` of the reference -> aline_b200.tasks.location_finding."""
from aline_b200.tasks.location_finding import *  # noqa: F401,F403
from aline_b200.tasks.location_finding import __dict__ as _d  # noqa: F401
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
