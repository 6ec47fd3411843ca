"""Shim: `distributions.censored_sigmoid_normal` of the reference -> aline_b200.distributions.censored_sigmoid_normal."""
from aline_b200.distributions.censored_sigmoid_normal import *  # noqa: F401,F403
from aline_b200.distributions.censored_sigmoid_normal import __dict__ as _d  # noqa: F401
globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
